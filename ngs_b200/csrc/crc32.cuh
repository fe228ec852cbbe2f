// CRC32 (IEEE 802.3, reflected, as in gzip) of every inflated BGZF block, compared with the
// block trailer — the check noodles-bgzf performs on each block it reads.
// One warp per block.  Every lane owns a contiguous slice of the block and runs TWO independent
// table CRCs over its halves, 64 bytes of each per loop iteration: the byte-table lookup chain (load -> xor ->
// shared-memory lookup) is latency bound, and one chain per lane left the kernel at 1.5 TB/s with 12.5 + 9 stall cycles
// per issue on the shared-memory and global scoreboards (ncu, profiles/round1_v2_kernels_ncu.md).  Four chains per lane
// were measured too: 0.9 M concurrent streams x 128-byte lines no longer fit the L2 and every line came from DRAM twice
// (profiles/round2_crc_variants.md).  The
// partial CRCs are merged with CRC(A||B) = CRC(A) * x^(8|B|) mod P xor CRC(B), a carry-less multiply
// modulo the CRC polynomial; the factors x^(8k) mod P for every k <= 65536 come from a 256 KB table
// (L2-resident) instead of a square-and-multiply chain per lane.
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

constexpr uint32_t kCrcShiftEntries = 65536 + 1;
struct CrcTables {
  uint32_t byte_table[256];
  uint32_t xpow8[kCrcShiftEntries];  // x^(8k) mod P: appending k zero bytes to a message multiplies its CRC by this
};

inline void crc_make_tables(CrcTables& t) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
    t.byte_table[i] = c;
  }
  uint32_t x8 = 1u << 31;  // x^0
  for (int k = 0; k < 8; ++k) x8 = (x8 & 1) ? (x8 >> 1) ^ kCrcPoly : x8 >> 1;  // x^8
  uint32_t p = 1u << 31;
  for (uint32_t k = 0; k < kCrcShiftEntries; ++k) {
    t.xpow8[k] = p;
    p = crc_multmodp(p, x8);
  }
}

// Table lookups dominate: the 256-entry table is replicated once per lane (entry i of lane l at word
// i * 32 + l, i.e. always in bank l) so the 32 data-dependent lookups of a warp never conflict, and
// every lane streams its quarters with 128-bit loads.
constexpr int kCrcThreads = 256;
constexpr int kCrcStreams = 2;
constexpr int kCrcStep = 64;  // bytes of a stream per loop iteration: half a 128-byte line, whole sectors
constexpr size_t kCrcSmem = 256 * 32 * 4;

__global__ void __launch_bounds__(kCrcThreads)
crc32_kernel(const uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, const uint32_t* __restrict__ expect,
             uint32_t n_blocks, const CrcTables* __restrict__ tables, uint32_t* __restrict__ n_bad) {
  extern __shared__ uint32_t crc_sm[];
  uint32_t* tab = crc_sm;  // [256][32]
  for (uint32_t i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = tables->byte_table[i >> 5];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t* tl = tab + lane;
  const uint32_t* __restrict__ xpow8 = tables->xpow8;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
#define NGSQ_CRC_BYTE(c, byte) c = tl[(((c) ^ (byte)) & 255u) << 5] ^ ((c) >> 8)
  for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
    const BlockDesc d = blocks[b];
    const uint8_t* p = out + d.out_off;
    const uint32_t n = d.isize;
    // lane slice [lo, hi): `per` bytes, a multiple of 128; part j = [lo + j*q, lo + (j+1)*q) clipped to n.
    // All part starts are multiples of 64 bytes into the block, so every stream has the block's alignment and
    // consumes whole 32-byte sectors: a lane's loads share no sector with any other lane's, and with 16-byte steps the
    // second half of every sector had left the (small) L1 before it was asked for — 4.6x DRAM traffic (ncu, profiles/).
    const uint32_t per = ((n + 31) / 32 + 127) & ~127u;
    const uint32_t q = per / kCrcStreams;
    const uint32_t lo = min(lane * per, n), hi = min(lo + per, n);
    uint32_t s_lo[kCrcStreams], s_hi[kCrcStreams], c[kCrcStreams];
#pragma unroll
    for (int j = 0; j < kCrcStreams; ++j) {
      s_lo[j] = min(lo + j * q, hi);
      s_hi[j] = min(s_lo[j] + q, hi);
      c[j] = 0xFFFFFFFFu;
    }
    const uint32_t head = (32u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 31)) & 31u;
    // head bytes up to the first sector boundary of each stream
    for (uint32_t t = 0; t < head; ++t) {
#pragma unroll
      for (int j = 0; j < kCrcStreams; ++j)
        if (s_lo[j] + t < s_hi[j]) NGSQ_CRC_BYTE(c[j], p[s_lo[j] + t]);
    }
    // whole sectors, the streams in lock step (stream 0 is never shorter than the others)
    uint32_t k_n[kCrcStreams];
#pragma unroll
    for (int j = 0; j < kCrcStreams; ++j) k_n[j] = s_lo[j] + head < s_hi[j] ? (s_hi[j] - s_lo[j] - head) / kCrcStep : 0u;
    for (uint32_t k = 0; k < k_n[0]; ++k) {
      uint4 v[kCrcStreams][kCrcStep / 16];
#pragma unroll
      for (int j = 0; j < kCrcStreams; ++j) {
#pragma unroll
        for (int t = 0; t < kCrcStep / 16; ++t) v[j][t] = make_uint4(0, 0, 0, 0);
        if (k < k_n[j]) {
          const uint4* src = reinterpret_cast<const uint4*>(p + s_lo[j] + head + kCrcStep * k);
#pragma unroll
          for (int t = 0; t < kCrcStep / 16; ++t) v[j][t] = src[t];
        }
      }
#pragma unroll
      for (int w = 0; w < kCrcStep / 4; ++w) {
#pragma unroll
        for (int sh = 0; sh < 32; sh += 8) {
#pragma unroll
          for (int j = 0; j < kCrcStreams; ++j) {
            const uint4& x = v[j][w >> 2];
            const uint32_t word = (w & 3) == 0 ? x.x : (w & 3) == 1 ? x.y : (w & 3) == 2 ? x.z : x.w;
            if (k < k_n[j]) NGSQ_CRC_BYTE(c[j], word >> sh);
          }
        }
      }
    }
    // tail bytes of each stream
#pragma unroll
    for (int j = 0; j < kCrcStreams; ++j)
      for (uint32_t i = s_lo[j] + head + kCrcStep * k_n[j]; i < s_hi[j]; ++i) NGSQ_CRC_BYTE(c[j], p[i]);
    // merge the lane's quarters (Horner), then shift the lane's CRC by the bytes that follow its slice
    uint32_t h = 0;
#pragma unroll
    for (int j = 0; j < kCrcStreams; ++j) {
      const uint32_t len = s_hi[j] - s_lo[j];
      if (len) {
        const uint32_t cj = c[j] ^ 0xFFFFFFFFu;  // standard CRC of this quarter (CRC of the empty string is 0)
        // the init / final xor of a prefix is absorbed by treating h as the CRC of the bytes before: crc(A||B) = shift(crc(A), |B|) ^ crc(B)
        h = (h ? crc_multmodp(xpow8[len], h) : 0u) ^ cj;
      }
    }
    const uint32_t after = n - hi;
    if (hi > lo && after && h) h = crc_multmodp(xpow8[after], h);
    if (hi == lo) h = 0;
    for (int o = 16; o; o >>= 1) h ^= __shfl_xor_sync(0xFFFFFFFFu, h, o);
    if (lane == 0 && h != expect[b]) atomicAdd(n_bad, 1u);
  }
#undef NGSQ_CRC_BYTE
}

}  // namespace ngsq
