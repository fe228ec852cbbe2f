// K3 — record-boundary scan over the inflated byte stream (reference: the record framing
// that bam::Reader::records / query perform one record at a time, src/qc/command.rs:305,
// :369-377).  BAM records form a linked list (next = off + 4 + block_size), which is serial
// over the whole stream.  We break the chain per BGZF block:
//   1. find_first  — one warp per BGZF block tests its first bytes, 32 candidate offsets at a
//                    time, for a plausible record header confirmed by a 3-record chain;
//   2. walk_count  — one thread per block walks from its first record to the block end, counts
//                    records and CHECKS CLOSURE: the walk must land exactly on the first record
//                    found for the block it lands in, and blocks it jumps over must have none;
//   3. exclusive scan of the counts;
//   4. walk_emit   — same walk, writes rec[i] = (block << 16) | offset_in_block.
// A closure failure sets an error word; the engine then redoes the shard with the serial
// fallback (chain_serial) so a false positive can never change results.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ngsq {

constexpr uint32_t kNoFirst = 0xFFFFFFFFu;
constexpr uint32_t kMaxRecordBytes = 1u << 28;

struct ScanErr {
  uint32_t chain;       // closure failures
  uint32_t bad_record;  // implausible record met on a verified chain
  uint32_t truncated;   // chain runs past the end of the data
  uint32_t max_lseq;
};

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = w[0];
  if (sh == 0) return lo;
  uint32_t hi = w[1];
  return __funnelshift_r(lo, hi, sh);
}

// Header plausibility of a record starting at `off` (absolute inflated offset).
// Returns the offset of the next record, or 0 when implausible.
__device__ __forceinline__ uint64_t plausible(const uint8_t* d, uint64_t off, uint64_t d_end, int32_t n_ref) {
  if (off + 36 > d_end) return 0;
  const uint8_t* p = d + off;
  uint32_t bs = ld_u32_unaligned(p);
  if (bs < 34 || bs > kMaxRecordBytes || off + 4 + bs > d_end) return 0;
  int32_t ref = (int32_t)ld_u32_unaligned(p + 4);
  int32_t pos = (int32_t)ld_u32_unaligned(p + 8);
  uint32_t w3 = ld_u32_unaligned(p + 12);  // l_read_name, mapq, bin
  uint32_t w4 = ld_u32_unaligned(p + 16);  // n_cigar, flag
  uint32_t lseq = ld_u32_unaligned(p + 20);
  int32_t nref = (int32_t)ld_u32_unaligned(p + 24);
  int32_t npos = (int32_t)ld_u32_unaligned(p + 28);
  uint32_t lname = w3 & 255, ncig = w4 & 0xFFFF;
  if (ref < -1 || ref >= n_ref || nref < -1 || nref >= n_ref) return 0;
  if (pos < -1 || npos < -1 || lname < 2) return 0;
  uint64_t need = 32ull + lname + 4ull * ncig + (lseq + 1ull) / 2 + lseq;
  if (need > bs) return 0;
  if (p[36 + lname - 1] != 0) return 0;  // read name is NUL-terminated
  uint8_t c0 = p[36];
  if (c0 < 33 || c0 > 126) return 0;
  return off + 4 + bs;
}

// first[b] for every block except block 0 of the shard (given by the caller).
__global__ void find_first_kernel(const uint8_t* __restrict__ d, const uint64_t* __restrict__ out_off, uint32_t n_blocks,
                                  uint64_t d_end, int32_t n_ref, uint64_t start_off, uint32_t* __restrict__ first) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= n_blocks) return;
  uint64_t lo = out_off[warp], hi = out_off[warp + 1];
  uint32_t res = kNoFirst;
  if (start_off >= hi) {
    // blocks wholly before the shard's first record hold nothing we own
  } else if (start_off >= lo) {
    res = (uint32_t)(start_off - lo);
  } else {
    for (uint64_t base = lo; base < hi; base += 32) {
      uint64_t c = base + lane;
      bool ok = false;
      if (c < hi) {
        uint64_t n1 = plausible(d, c, d_end, n_ref);
        if (n1) {
          ok = true;
          // confirm with up to two further records when they fit in the data
          uint64_t n2 = n1 + 36 <= d_end ? plausible(d, n1, d_end, n_ref) : n1;
          if (!n2) ok = false;
          else if (n2 != n1) {
            uint64_t n3 = n2 + 36 <= d_end ? plausible(d, n2, d_end, n_ref) : n2;
            if (!n3) ok = false;
          }
        }
      }
      uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      if (m) {
        res = (uint32_t)(base - lo) + (__ffs(m) - 1);
        break;
      }
    }
  }
  if (lane == 0) first[warp] = res;
}

// One thread per block.  EMIT=false: count + closure; EMIT=true: write the offset table.
template <bool EMIT>
__global__ void walk_kernel(const uint8_t* __restrict__ d, const uint64_t* __restrict__ out_off, uint32_t n_blocks,
                            uint64_t d_end, uint64_t end_off, const uint32_t* __restrict__ first,
                            uint32_t* __restrict__ landed, uint32_t* __restrict__ count,
                            const uint64_t* __restrict__ base, uint64_t* __restrict__ rec, ScanErr* __restrict__ err) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t f = first[b];
  if (f == kNoFirst) {
    if (!EMIT) count[b] = 0;
    return;
  }
  uint64_t lo = out_off[b], hi = out_off[b + 1];
  uint64_t off = lo + f;
  uint64_t stop = hi < end_off ? hi : end_off;
  uint32_t n = 0, max_lseq = 0;
  uint64_t w = EMIT ? base[b] : 0;
  while (off < stop) {
    if (off + 36 > d_end) { if (!EMIT) atomicAdd(&err->truncated, 1u); off = stop; break; }
    uint32_t bs = ld_u32_unaligned(d + off);
    if (bs < 32 || bs > kMaxRecordBytes) { if (!EMIT) atomicAdd(&err->bad_record, 1u); off = stop; break; }
    if (off + 4 + bs > d_end) { if (!EMIT) atomicAdd(&err->truncated, 1u); off = stop; break; }
    if (EMIT) rec[w + n] = ((uint64_t)b << 16) | (off - lo);
    else {
      // l_seq is not validated yet (the facet kernel does that): only a length the record can hold may size the quality table
      uint32_t lseq = ld_u32_unaligned(d + off + 20);
      if ((uint64_t)lseq + (lseq + 1ull) / 2 + 32 <= bs) max_lseq = lseq > max_lseq ? lseq : max_lseq;
      else atomicAdd(&err->bad_record, 1u);
    }
    ++n;
    off += 4 + bs;
  }
  if (EMIT) return;
  count[b] = n;
  atomicMax(&err->max_lseq, max_lseq);
  // closure: the walk must end exactly on the shard end or on the first record of the block it lands in
  if (off >= end_off) {
    if (off != end_off) atomicAdd(&err->chain, 1u);
    return;
  }
  uint32_t j = b + 1;
  while (j < n_blocks && out_off[j + 1] <= off) {
    if (first[j] != kNoFirst) atomicAdd(&err->chain, 1u);  // a record start claimed inside a record
    ++j;
  }
  if (j >= n_blocks) { atomicAdd(&err->truncated, 1u); return; }
  if (first[j] == kNoFirst || out_off[j] + first[j] != off) atomicAdd(&err->chain, 1u);
  else landed[j] = 1;
}

// Every block that claims a first record (other than the shard's first) must have been landed on.
__global__ void check_landed_kernel(const uint32_t* __restrict__ first, const uint32_t* __restrict__ landed,
                                    uint32_t n_blocks, uint32_t first_block, ScanErr* err) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks || b == first_block) return;
  if (first[b] != kNoFirst && !landed[b]) atomicAdd(&err->chain, 1u);
}

// Serial fallback: one thread walks the whole chain and fills first[] from scratch.
__global__ void chain_serial_kernel(const uint8_t* __restrict__ d, const uint64_t* __restrict__ out_off, uint32_t n_blocks,
                                    uint64_t d_end, uint64_t start_off, uint64_t end_off, uint32_t* __restrict__ first,
                                    ScanErr* err) {
  if (blockIdx.x || threadIdx.x) return;
  uint64_t off = start_off;
  uint32_t b = 0;
  for (uint32_t i = 0; i < n_blocks; ++i) first[i] = kNoFirst;
  while (off < end_off) {
    while (b < n_blocks && out_off[b + 1] <= off) ++b;
    if (b >= n_blocks) { atomicAdd(&err->truncated, 1u); return; }
    if (first[b] == kNoFirst) first[b] = (uint32_t)(off - out_off[b]);
    if (off + 36 > d_end) { atomicAdd(&err->truncated, 1u); return; }
    uint32_t bs = ld_u32_unaligned(d + off);
    if (bs < 32 || off + 4 + bs > d_end) { atomicAdd(&err->bad_record, 1u); return; }
    off += 4 + bs;
  }
}

// Exclusive scan of u32 counts into u64 bases; single CTA, grid-stride over chunks.
__global__ void scan_counts_kernel(const uint32_t* __restrict__ count, uint32_t n, uint64_t* __restrict__ base,
                                   uint64_t* __restrict__ total) {
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t start = 0; start < n; start += blockDim.x) {
    uint32_t i = start + threadIdx.x;
    uint64_t v = i < n ? count[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint64_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if ((int)lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    uint64_t carry = carry_s;
    uint64_t wbase = wid ? warp_sums[wid - 1] : 0;
    if (i < n) base[i] = carry + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + wbase + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

}  // namespace ngsq
