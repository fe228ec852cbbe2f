// CRC32 (IEEE 802.3, reflected, as in gzip) of every inflated BGZF block, compared with the
// block trailer — the check noodles-bgzf performs on each block it reads.
// One warp per block: every lane runs a table CRC over its own contiguous slice,
// then the 32 partial CRCs are merged with CRC(A||B) = CRC(A) * x^(8|B|) mod P  xor  CRC(B),
// a carry-less multiply modulo the CRC polynomial (no 32x32 GF(2) matrices needed).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "inflate_lane.cuh"  // BlockDesc

namespace ngsq {

constexpr uint32_t kCrcPoly = 0xEDB88320u;

__host__ __device__ inline uint32_t crc_multmodp(uint32_t a, uint32_t b) {
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
  }
  return p;
}

struct CrcTables {
  uint32_t byte_table[256];
  uint32_t x2n[32];  // x^(2^n) mod P
};

inline void crc_make_tables(CrcTables& t) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
    t.byte_table[i] = c;
  }
  uint32_t p = 1u << 30;  // x^1
  t.x2n[0] = p;
  for (int n = 1; n < 32; ++n) t.x2n[n] = p = crc_multmodp(p, p);
}

// x^(n * 2^k) mod P
__device__ __forceinline__ uint32_t crc_x2nmodp(const uint32_t* x2n, uint32_t n, uint32_t k) {
  uint32_t p = 1u << 31;  // x^0
  while (n) {
    if (n & 1) p = crc_multmodp(x2n[k & 31], p);
    n >>= 1;
    ++k;
  }
  return p;
}

// Table lookups dominate: the 256-entry table is replicated once per lane (entry i of lane l at word
// i * 32 + l, i.e. always in bank l) so the 32 data-dependent lookups of a warp never conflict, and
// every lane streams its slice with 128-bit loads (byte loads made the kernel L1-wavefront bound:
// ncu, profiles/).  Round 2 measured two more forms of this kernel — four and two independent chains per lane with
// whole-sector loads and a table of shift factors — and both lost (3.4 ms against 2.3 ms per 3.6 GB): the kernel is bound
// by the shared-memory gather pipe (l1tex data-pipe wavefronts 56 % of peak at one lookup per byte), not by the
// dependency chain, and more streams per lane only spread the working set over more L2 lines
// (profiles/round2_crc_variants.md).
constexpr int kCrcThreads = 256;
constexpr size_t kCrcSmem = 256 * 32 * 4 + 32 * 4;

__global__ void __launch_bounds__(kCrcThreads)
crc32_kernel(const uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, const uint32_t* __restrict__ expect,
             uint32_t n_blocks, const CrcTables* __restrict__ tables, uint32_t* __restrict__ n_bad) {
  extern __shared__ uint32_t crc_sm[];
  uint32_t* tab = crc_sm;             // [256][32]
  uint32_t* x2n = crc_sm + 256 * 32;  // [32]
  for (uint32_t i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = tables->byte_table[i >> 5];
  if (threadIdx.x < 32) x2n[threadIdx.x] = tables->x2n[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t* tl = tab + lane;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
#define NGSQ_CRC_BYTE(c, byte) c = tl[(((c) ^ (byte)) & 255u) << 5] ^ ((c) >> 8)
  for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
    const BlockDesc d = blocks[b];
    const uint8_t* p = out + d.out_off;
    const uint32_t n = d.isize;
    // contiguous slice per lane, a multiple of 16 bytes: every lane has the block's alignment
    const uint32_t per = ((n + 31) / 32 + 15) & ~15u;
    uint32_t lo = lane * per, hi = lo + per;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    uint32_t c = 0xFFFFFFFFu;
    uint32_t i = lo;
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p + lo) & 15)) & 15u;
    const uint32_t he = min(hi, lo + head);
    for (; i < he; ++i) NGSQ_CRC_BYTE(c, p[i]);
    for (; i + 16 <= hi; i += 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(p + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        NGSQ_CRC_BYTE(c, w[k]);
        NGSQ_CRC_BYTE(c, w[k] >> 8);
        NGSQ_CRC_BYTE(c, w[k] >> 16);
        NGSQ_CRC_BYTE(c, w[k] >> 24);
      }
    }
    for (; i < hi; ++i) NGSQ_CRC_BYTE(c, p[i]);
    c ^= 0xFFFFFFFFu;  // standard CRC of this slice (CRC of the empty string is 0)
    // shift by the bytes that follow this slice, then xor-reduce
    const uint32_t after = n - hi;
    if (hi > lo && after) c = crc_multmodp(crc_x2nmodp(x2n, after, 3), c);
    if (hi == lo) c = 0;
    for (int o = 16; o; o >>= 1) c ^= __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0 && c != expect[b]) atomicAdd(n_bad, 1u);
  }
#undef NGSQ_CRC_BYTE
}

}  // namespace ngsq
