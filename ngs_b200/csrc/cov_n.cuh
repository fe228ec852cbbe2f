// Coverage under `-n` — the second pass's record counter (reference: src/qc/command.rs:350-397 with
// RecordCounter, src/utils/display.rs:43-65).
//
// Pass 2 keeps ONE counter across all reference sequences: it is incremented for every record a contig's query
// yields (whether or not a facet supports the contig), and `time_to_break` only leaves the CURRENT contig's loop.
// So with `-n N` the reference processes
//     the first N query-yielded records in file order, and — because the counter stays >= N afterwards —
//     exactly the first yielded record of every later contig.
// In file order (a coordinate-sorted BAM lists its contigs in header order, which the BAI requires anyway) a
// yielded record with global yield rank Y is therefore processed iff  Y < N  or  it is the first yielded record
// of its contig.  Three small kernels per wave, used only when max_records != 0 (the facet kernel scatters
// coverage itself otherwise):
//   cov_n_mark   per record: reference id + 1 if the query would yield it, else 0
//   cov_n_rank   one CTA: running yield count + the previous yielded record's contig (state crosses waves in RunState)
//   cov_n_apply  per processed record: the same scatter as facets.cuh
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "facets.cuh"

namespace ngsq {

struct CovNParams {
  const uint8_t* d;
  const uint64_t* rec;
  RunState* st;
  uint64_t max_records;
  int32_t n_ref;
  const uint32_t* ref_len;
  const uint8_t* cov_enabled;
  const uint64_t* diff_base;
  int32_t* diff;
  const uint32_t* cov_slot;
  const uint32_t* tile_off;
  int32_t* tile_sum;
  uint64_t* res;
  uint32_t* mark;  // one word per record of the wave
};

// Interval of a record as the per-contig query sees it: yielded iff the record names a
// reference, has a position, and [start, end] meets [1, L].  Malformed records are left to the facet kernel's
// verdict (never yielded here).
__device__ __forceinline__ bool cov_n_interval(const uint8_t* p, int32_t n_ref, const uint32_t* ref_len, int32_t* ref_out, int64_t* start_out,
                                               int64_t* end_out, uint32_t* span_out) {
  const uint32_t bs = ld_u32_unaligned(p);
  const int32_t ref = (int32_t)ld_u32_unaligned(p + 4), pos = (int32_t)ld_u32_unaligned(p + 8);
  const uint32_t w3 = ld_u32_unaligned(p + 12), w4 = ld_u32_unaligned(p + 16), lseq = ld_u32_unaligned(p + 20);
  const uint32_t lname = w3 & 255, ncig = w4 & 0xFFFF;
  if (32ull + lname + 4ull * ncig + (lseq + 1ull) / 2 + lseq > bs) return false;
  if (ref < 0 || ref >= n_ref || pos < 0) return false;
  const uint8_t* cig = p + 36 + lname;
  uint32_t span = 0;
  for (uint32_t i = 0; i < ncig; ++i) {
    const uint32_t op = ld_u32_unaligned(cig + 4 * i), k = op & 15;
    if (k > 8) return false;
    if ((0x18D >> k) & 1) span += op >> 4;  // M D N = X (utils/cigar.rs:6-11)
  }
  const int64_t L = ref_len[ref];
  const int64_t start = (int64_t)pos + 1, end = start + (int64_t)span - 1;
  if (!(start <= L && end >= 1)) return false;
  *ref_out = ref; *start_out = start; *end_out = end; *span_out = span;
  return true;
}

__global__ void __launch_bounds__(256) cov_n_mark_kernel(CovNParams P) {
  const uint64_t n_rec = P.st->fatal ? 0 : P.st->wave_rec;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (uint64_t)gridDim.x * blockDim.x) {
    int32_t ref; int64_t a, b; uint32_t span;
    P.mark[r] = cov_n_interval(P.d + (P.rec[r] & kRecOffMask), P.n_ref, P.ref_len, &ref, &a, &b, &span) ? (uint32_t)ref + 1u : 0u;
  }
}

// One CTA walks the wave's marks in file order, 1024 at a time: scan of (yield count, last yielded contig).
__global__ void __launch_bounds__(1024) cov_n_rank_kernel(CovNParams P) {
  __shared__ uint32_t w_cnt[32], w_last[32];
  __shared__ uint64_t run_cnt;
  __shared__ uint32_t run_last;
  RunState* st = P.st;
  const uint64_t n_rec = st->fatal ? 0 : st->wave_rec;
  if (threadIdx.x == 0) { run_cnt = st->cov_yielded; run_last = st->cov_last_ref; }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint64_t start = 0; start < n_rec; start += blockDim.x) {
    const uint64_t r = start + threadIdx.x;
    const uint32_t m = r < n_rec ? P.mark[r] : 0u;
    // inclusive scan of (count, last non-zero mark) over the warp
    uint32_t cnt = m ? 1u : 0u, last = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t c2 = __shfl_up_sync(0xFFFFFFFFu, cnt, o), l2 = __shfl_up_sync(0xFFFFFFFFu, last, o);
      if ((int)lane >= o) { cnt += c2; if (!last) last = l2; }
    }
    if (lane == 31) { w_cnt[wid] = cnt; w_last[wid] = last; }
    __syncthreads();
    if (wid == 0) {
      uint32_t c = w_cnt[lane], l = w_last[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t c2 = __shfl_up_sync(0xFFFFFFFFu, c, o), l2 = __shfl_up_sync(0xFFFFFFFFu, l, o);
        if ((int)lane >= o) { c += c2; if (!l) l = l2; }
      }
      w_cnt[lane] = c; w_last[lane] = l;
    }
    __syncthreads();
    // exclusive prefix of this thread: earlier lanes of the warp, earlier warps of the chunk, earlier chunks / waves
    uint32_t ex_cnt = __shfl_up_sync(0xFFFFFFFFu, cnt, 1), ex_last = __shfl_up_sync(0xFFFFFFFFu, last, 1);
    if (lane == 0) { ex_cnt = 0; ex_last = 0; }
    const uint32_t wb_cnt = wid ? w_cnt[wid - 1] : 0u, wb_last = wid ? w_last[wid - 1] : 0u;
    const uint64_t Y = run_cnt + wb_cnt + ex_cnt;
    const uint32_t prev = ex_last ? ex_last : (wb_last ? wb_last : run_last);
    if (r < n_rec) P.mark[r] = (m && (Y < P.max_records || prev != m)) ? 1u : 0u;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
      run_cnt += w_cnt[31];
      if (w_last[31]) run_last = w_last[31];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { st->cov_yielded = run_cnt; st->cov_last_ref = run_last; }
}

__global__ void __launch_bounds__(256) cov_n_apply_kernel(CovNParams P) {
  const uint64_t n_rec = P.st->fatal ? 0 : P.st->wave_rec;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (uint64_t)gridDim.x * blockDim.x) {
    if (!P.mark[r]) continue;
    int32_t ref; int64_t start, end; uint32_t span;
    if (!cov_n_interval(P.d + (P.rec[r] & kRecOffMask), P.n_ref, P.ref_len, &ref, &start, &end, &span)) continue;
    if (!P.cov_enabled[ref]) continue;  // counted, but no facet supports the contig (coverage.rs:133-138)
    const int64_t L = P.ref_len[ref];
    P.res[P.cov_slot[ref] + COV_TOUCHED] = 1;
    if (span) {
      int32_t* df = P.diff + P.diff_base[ref];
      atomicAdd(df + start, 1);
      const int64_t e = end < L ? end : L;
      atomicAdd(df + e + 1, -1);
      if (end > L) atomicAdd((unsigned long long*)&P.res[R_NONSENSICAL], (unsigned long long)(end - L));
      atomicAdd(P.tile_sum + P.tile_off[ref] + (uint32_t)(start >> 12), 1);
      if (e + 1 <= L) atomicAdd(P.tile_sum + P.tile_off[ref] + (uint32_t)((e + 1) >> 12), -1);
    }
  }
}

}  // namespace ngsq
