// K2 — BGZF inflate in ONE kernel (reference: the inflate noodles-bgzf/miniz_oxide perform
// under bam::Reader, src/utils/formats/bam.rs:41-44, src/qc/command.rs:305 and :350).
//
//   inflate_kernel   one BGZF block per LANE: 32 independent decoders per warp instruction, each finishing its own
//                    block — Huffman decode and LZ77 copies (inflate_lane.cuh).  Persistent grid, one CTA per SM
//                    (shared-memory bound: one 396-byte table slab per lane, 18 warps = 576 decoders per SM); a warp
//                    pulls 32 consecutive blocks at a time so that the lanes of a warp meet their DEFLATE block
//                    headers together (zlib ends a block every 16383 symbols) and parse/build them in lock step.
// An integer kernel bound by instruction issue (70 % of the slots at 4.5 warps per scheduler) and by the dependency
// chain of the symbol loop; HBM traffic is C read + D written + the match sources that left the L2.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "inflate_lane.cuh"

namespace ngsq {

constexpr int kDecBatch = 32;  // loop iterations per lane between header / overrun checks
constexpr int kDecWarps = 18;  // 18 x 32 x 396 B of slabs + 256 B of static shared memory
constexpr int kDecThreads = kDecWarps * 32;
constexpr size_t kDecSmem = (size_t)kDecThreads * kSlabBytes;

__global__ void __launch_bounds__(kDecThreads, 1)
inflate_kernel(uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, uint32_t n_blocks, uint32_t* __restrict__ queue,
               uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_base_lut[64];
  const uint32_t lane = threadIdx.x & 31;
  Lane L;
  L.slab = smem_raw + (size_t)threadIdx.x * kSlabBytes;
  L.state = LS_IDLE;
  if (threadIdx.x < 64) s_base_lut[threadIdx.x] = base_lut_entry(threadIdx.x);
  __syncthreads();
  L.lut = s_base_lut;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(queue, 32u);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n_blocks) break;
    const uint32_t b = base + lane;
    if (b < n_blocks) L.begin_block(blocks[b], out);
    else L.state = LS_IDLE;
    while (__any_sync(0xFFFFFFFFu, L.state != LS_IDLE)) {
      if (L.state == LS_HEADER) L.header();
      __syncwarp();
#pragma unroll 1
      for (int it = 0; it < kDecBatch; ++it) {
        if (L.state == LS_DECODE) L.step();
        L.settle();  // the input load of this iteration's window slide, consumed only here
        if (!__any_sync(0xFFFFFFFFu, L.state == LS_DECODE)) break;
      }
      if (L.state == LS_DECODE && L.overran()) L.end_block(kBlkBadStream);
    }
    if (b < n_blocks && L.err) status[b] = L.err;
  }
}

}  // namespace ngsq
