// K9 — resolve one contig's int32 difference array into what CoverageFacet::teardown
// computes with an O(L) scalar sweep over a Vec<usize> (coverage.rs:182-262):
//   * depth histogram over positions 0..=L, capped at 2048 (deeper -> pileup_too_large),
//   * per-50 kb running sums: bin 0 = {0}, bin k = (50000(k-1), 50000k], partial tail bin.
// The depth array itself is never materialised: tile sums -> scan of tile sums -> per tile
// local scan fused with the histogram (run-length aggregated, shared-memory privatised) and
// the bin sums.  HBM-bound: the difference array is read twice (4 B/position each).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "facets.cuh"

namespace ngsq {

constexpr int kCovThreads = 256;
constexpr int kCovPerThread = 16;
constexpr int kCovTile = kCovThreads * kCovPerThread;  // 4096 positions < bin size
constexpr uint32_t kCovBin = 50000;

// depth -> slot of the CTA-private histogram; [2049] collects depths beyond the reference's 2048 cap.  A negative running depth
// cannot come from the facet kernel's scatter (+1 at a start precedes its -1); should the int32 array ever wrap, stay in bounds.
__device__ __forceinline__ uint32_t cov_hist_slot(int64_t d) { return (d > 2048 || d < 0) ? 2049u : (uint32_t)d; }

// tile_sum[t] = sum of diff over tile t  (positions [t*4096, ...) clipped to n = L+1)
__global__ void __launch_bounds__(kCovThreads)
cov_tile_sums_kernel(const int32_t* __restrict__ diff, uint32_t n, const uint64_t* __restrict__ touched, int64_t* __restrict__ tile_sum) {
  if (!*touched) return;
  __shared__ int64_t ws[kCovThreads / 32];
  uint32_t t = blockIdx.x;
  uint32_t base = t * kCovTile + threadIdx.x * 4;
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < kCovPerThread / 4; ++k) {
    uint32_t i = base + k * (kCovThreads * 4);
    if (i + 3 < n) {
      int4 v = *reinterpret_cast<const int4*>(diff + i);
      s += (int64_t)v.x + v.y + v.z + v.w;
    } else {
      for (uint32_t j = i; j < n && j < i + 4; ++j) s += diff[j];
    }
  }
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t tot = 0;
    for (int w = 0; w < kCovThreads / 32; ++w) tot += ws[w];
    tile_sum[t] = tot;
  }
}

// exclusive scan of tile sums in place (single CTA)
__global__ void cov_scan_tiles_kernel(int64_t* __restrict__ tile_sum, uint32_t n_tiles, const uint64_t* __restrict__ touched) {
  if (!*touched) return;
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t start = 0; start < n_tiles; start += blockDim.x) {
    uint32_t i = start + threadIdx.x;
    int64_t v = i < n_tiles ? tile_sum[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if ((int)lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    int64_t carry = carry_s;
    int64_t wbase = wid ? warp_sums[wid - 1] : 0;
    if (i < n_tiles) tile_sum[i] = carry + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + wbase + x;
    __syncthreads();
  }
}

// Persistent CTAs loop over tiles; slot = this contig's [touched, too_large, hist[2049], bins[]].
__global__ void __launch_bounds__(kCovThreads)
cov_resolve_kernel(const int32_t* __restrict__ diff, uint32_t n, const int64_t* __restrict__ tile_base, uint32_t n_tiles,
                   uint64_t* __restrict__ slot) {
  if (!slot[COV_TOUCHED]) return;
  __shared__ uint32_t hist[2050];  // [2049] = deeper than 2048
  __shared__ int64_t ws[kCovThreads / 32];
  __shared__ unsigned long long bsum[2];
  for (uint32_t i = threadIdx.x; i < 2050; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint32_t p0 = t * kCovTile + threadIdx.x * kCovPerThread;  // first position of this thread
    int32_t v[kCovPerThread];
#pragma unroll
    for (int k = 0; k < kCovPerThread; k += 4) {
      uint32_t i = p0 + k;
      if (i + 3 < n) {
        int4 q = *reinterpret_cast<const int4*>(diff + i);
        v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[k + j] = (i + j < n) ? diff[i + j] : 0;
      }
    }
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kCovPerThread; ++k) tsum += v[k];
    // exclusive scan of thread sums across the CTA
    int64_t x = tsum;
    for (int o = 1; o < 32; o <<= 1) {
      int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if ((int)lane >= o) x += y;
    }
    if (threadIdx.x < 2) bsum[threadIdx.x] = 0;
    if (lane == 31) ws[wid] = x;
    __syncthreads();
    int64_t wbase = 0;
    for (uint32_t w = 0; w < wid; ++w) wbase += ws[w];
    int64_t depth = tile_base[t] + wbase + x - tsum;
    // positions -> depth; run-length aggregated histogram; bin sums
    const uint32_t k0 = (t * kCovTile + kCovBin - 1) / kCovBin;  // bin of the tile's first position (0 for position 0)
    unsigned long long s0 = 0, s1 = 0;
    int64_t run_d = -1;
    uint32_t run_n = 0;
#pragma unroll
    for (int k = 0; k < kCovPerThread; ++k) {
      uint32_t i = p0 + k;
      depth += v[k];
      if (i < n) {
        uint32_t bin = (i + kCovBin - 1) / kCovBin;
        if (bin == k0) s0 += (unsigned long long)depth; else s1 += (unsigned long long)depth;
        if (depth == run_d) ++run_n;
        else {
          if (run_n) atomicAdd(&hist[cov_hist_slot(run_d)], run_n);
          run_d = depth;
          run_n = 1;
        }
      }
    }
    if (run_n) atomicAdd(&hist[cov_hist_slot(run_d)], run_n);
    // 64-bit butterfly: two independent 32-bit reductions of the halves would drop the carry out of the low words once the
    // 32 lanes' sums pass 2^32 (depth of a few million over the warp's 512 positions: amplicon / rRNA data)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s0 += __shfl_down_sync(0xFFFFFFFFu, s0, o);
      s1 += __shfl_down_sync(0xFFFFFFFFu, s1, o);
    }
    if (lane == 0) {
      if (s0) atomicAdd(&bsum[0], s0);
      if (s1) atomicAdd(&bsum[1], s1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (bsum[0]) atomicAdd((unsigned long long*)&slot[COV_BINS + k0], bsum[0]);
      if (bsum[1]) atomicAdd((unsigned long long*)&slot[COV_BINS + k0 + 1], bsum[1]);
    }
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i <= 2048; i += blockDim.x)
    if (hist[i]) atomicAdd((unsigned long long*)&slot[COV_HIST + i], (unsigned long long)hist[i]);
  if (threadIdx.x == 0 && hist[2049]) atomicAdd((unsigned long long*)&slot[COV_TOO_LARGE], (unsigned long long)hist[2049]);
}

}  // namespace ngsq
