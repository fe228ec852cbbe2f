// K2 — DEFLATE inflate of independent BGZF blocks (reference: the inflate that
// noodles-bgzf/miniz_oxide perform under bam::Reader, src/utils/formats/bam.rs:41-44).
//
// Shape: a group of G lanes (G = 4/8/16/32; 32 = the warp-per-block form) owns one
// BGZF block at a time and pulls the next block id from a global atomic queue, so
// the grid is persistent and block-size skew does not matter.  Within a group:
//   * lane 0 ("leader") owns the bit reader and decodes Huffman symbols serially
//     (that part of DEFLATE is inherently bit-serial); literals are stored by the
//     leader as they are decoded;
//   * when the leader hits a match it hands (length, distance) to the whole group
//     through one shuffle and all G lanes copy the LZ77 match;
//   * dynamic-Huffman tables are built cooperatively by the G lanes into a
//     per-group shared-memory slab (10-bit literal/length LUT, 8-bit distance LUT,
//     canonical count/sorted-symbol arrays for the rare longer codes).
// Several groups share a warp: the symbol loop is the same instruction stream for
// every group, so G < 32 multiplies the number of concurrently decoding lanes per
// issued instruction.  This kernel is issue/latency-bound, not HBM-bound.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "inflate_lane.cuh"  // BlockDesc, kBlk* status codes

namespace ngsq {

constexpr int kLLBits = 10;
constexpr int kDBits = 8;
constexpr int kInflateThreads = 128;

struct __align__(16) DecSmem {
  uint16_t ll_lut[1 << kLLBits];  // (sym << 4) | len, 0 = longer than kLLBits / unused
  uint16_t d_lut[1 << kDBits];
  uint16_t ll_sorted[288];        // symbols in canonical order (slow path)
  uint16_t d_sorted[32];
  uint32_t ll_count[16];
  uint32_t d_count[16];
  uint32_t first_code[16];
  uint32_t offs[16];
  uint32_t run[16];
  uint16_t ll_lim[16], d_lim[16];  // slow path: (first_code[l] + count[l]) << (15 - l)
  int16_t ll_base[16], d_base[16]; // slow path: offs[l] - first_code[l]
  uint32_t tok_pl[8];              // queued matches: pos | (len - 3) << 16
  uint32_t tok_d[8];               //                 dist | dependent << 31
  uint8_t lens[344];              // [0,320) litlen+dist code lengths, [320,339) code-length code lengths
  uint8_t cl_lut[128];            // (sym << 3) | len
};

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Canonical Huffman tables from code lengths, built by the G lanes of a group.
template <int G>
__device__ __forceinline__ void build_table(const uint8_t* lens, int n, uint16_t* lut, int lut_bits, uint16_t* sorted,
                                            uint32_t* count, uint16_t* lim, int16_t* base, DecSmem* s, uint32_t gmask,
                                            int lig, int lane) {
  for (int i = lig; i < 16; i += G) { count[i] = 0; s->run[i] = 0; }
  for (int i = lig; i < (1 << lut_bits) / 2; i += G) reinterpret_cast<uint32_t*>(lut)[i] = 0;
  __syncwarp(gmask);
  for (int k = lig; k < n; k += G) {
    uint32_t l = lens[k];
    if (l) atomicAdd(&count[l], 1u);
  }
  __syncwarp(gmask);
  if (lig == 0) {
    uint32_t code = 0, off = 0, prev = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code + prev) << 1;
      prev = count[l];
      s->first_code[l] = code;
      s->offs[l] = off;
      lim[l] = (uint16_t)min((code + prev) << (15 - l), 0xFFFFu);
      base[l] = (int16_t)((int)off - (int)code);
      off += prev;
    }
  }
  __syncwarp(gmask);
  for (int c = 0; c < n; c += G) {
    int k = c + lig;
    uint32_t l = k < n ? lens[k] : 0;
    uint32_t m = __match_any_sync(gmask, l);
    uint32_t r = __popc(m & lanemask_lt());
    uint32_t base = s->run[l];
    __syncwarp(gmask);
    if (l && (m >> lane) == 1u) s->run[l] = base + __popc(m);
    __syncwarp(gmask);
    if (l) {
      uint32_t rank = base + r;
      uint32_t code = s->first_code[l] + rank;
      sorted[s->offs[l] + rank] = (uint16_t)k;
      if ((int)l <= lut_bits) {
        uint32_t rev = __brev(code) >> (32 - l);
        uint16_t e = (uint16_t)((k << 4) | l);
        for (uint32_t j = rev; j < (1u << lut_bits); j += (1u << l)) lut[j] = e;
      }
    }
  }
  __syncwarp(gmask);
}

// Canonical decode for codes longer than the LUT width (leader only): left-align the next 15 bits
// MSB-first; canonical codes are ordered, so the first length whose limit exceeds them is the length.
template <int LUT_BITS>
__device__ __forceinline__ int slow_decode(uint32_t bits, const uint16_t* lim, const int16_t* base, const uint16_t* sorted,
                                           int& len_out) {
  const uint32_t x = __brev(bits) >> 17;
#pragma unroll
  for (int l = LUT_BITS + 1; l <= 15; ++l) {
    if (x < lim[l]) {
      len_out = l;
      return sorted[(int)(x >> (15 - l)) + base[l]];
    }
  }
  len_out = 0;
  return -1;
}

struct BitReader {
  uint64_t bb;           // bit buffer, LSB first
  int bc;                // valid bits in bb
  uint32_t nw0, nw1;     // two prefetched words: an L1 miss on the input stream has ~64 bits of slack
  const uint32_t* wptr;  // next aligned word to prefetch
  __device__ __forceinline__ void init(const uint8_t* p) {
    uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
    wptr = reinterpret_cast<const uint32_t*>(p - mis);
    uint32_t w = __ldg(wptr++);
    bb = (uint64_t)(w >> (8 * mis));
    bc = 32 - 8 * (int)mis;
    nw0 = __ldg(wptr++);
    nw1 = __ldg(wptr++);
    refill();
  }
  __device__ __forceinline__ void refill() {
    if (bc <= 32) {
      bb |= (uint64_t)nw0 << bc;
      bc += 32;
      nw0 = nw1;
      nw1 = __ldg(wptr++);
    }
  }
  __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)bb; }
  __device__ __forceinline__ void drop(int n) { bb >>= n; bc -= n; }
  __device__ __forceinline__ uint32_t take(int n) {
    uint32_t v = (uint32_t)bb & ((1u << n) - 1u);
    drop(n);
    return v;
  }
  // address of the next unread byte once the reader is byte-aligned
  __device__ __forceinline__ const uint8_t* byte_ptr() const {
    return reinterpret_cast<const uint8_t*>(wptr) - 8 - (bc >> 3);
  }
};

constexpr int kTokens = 8;       // match tokens a leader may queue per batch
constexpr int kBatchIters = 40;  // symbols a leader may decode per batch
enum : int { ST_BLOCK = 0, ST_HEADER = 1, ST_DECODE = 2, ST_FINISH = 3, ST_DONE = 4 };

template <int G>
struct LeaderMask;
template <> struct LeaderMask<32> { static constexpr uint32_t v = 0x00000001u; };
template <> struct LeaderMask<16> { static constexpr uint32_t v = 0x00010001u; };
template <> struct LeaderMask<8> { static constexpr uint32_t v = 0x01010101u; };
template <> struct LeaderMask<4> { static constexpr uint32_t v = 0x11111111u; };
template <> struct LeaderMask<2> { static constexpr uint32_t v = 0x55555555u; };

// one lane copies one LZ77 match; sources never include bytes of this match (periodic extension)
__device__ __forceinline__ void lane_copy(uint8_t* dst, uint32_t mlen, uint32_t dist) {
  const uint8_t* src = dst - dist;
  if (dist >= mlen) {
    uint32_t k = 0;
    for (; k + 4 <= mlen; k += 4) {
      uint8_t a = src[k], b = src[k + 1], c = src[k + 2], d = src[k + 3];
      dst[k] = a; dst[k + 1] = b; dst[k + 2] = c; dst[k + 3] = d;
    }
    for (; k < mlen; ++k) dst[k] = src[k];
  } else if (dist == 1) {
    uint8_t v = src[0];
    for (uint32_t k = 0; k < mlen; ++k) dst[k] = v;
  } else {
    uint32_t sidx = 0;
    for (uint32_t k = 0; k < mlen; ++k) {
      dst[k] = src[sidx];
      sidx = sidx + 1 == dist ? 0 : sidx + 1;
    }
  }
}

// Persistent kernel: every warp runs one flat state machine; all 32 lanes reconverge at the loop
// top each iteration, so the 32/G leaders of a warp execute the symbol loop in lock step.
template <int G>
__global__ void __launch_bounds__(kInflateThreads)
inflate_kernel(uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks,
               uint32_t n_blocks, uint32_t* __restrict__ queue, uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int lig = threadIdx.x % G;
  const int gid = threadIdx.x / G;
  const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane - lig));
  const int leader_lane = lane - lig;
  constexpr uint32_t kLeaders = LeaderMask<G>::v;
  DecSmem* s = reinterpret_cast<DecSmem*>(smem_raw) + gid;

  BitReader br;
  br.bb = 0; br.bc = 0; br.nw0 = 0; br.nw1 = 0; br.wptr = nullptr;
  uint8_t* obase = nullptr;  // output of the current block
  const uint8_t* in_end = nullptr;  // end of the block's DEFLATE payload (leader)
  uint32_t pos = 0, isize = 0, blk = 0;
  int state = ST_BLOCK;
  bool bfinal = false;
  uint32_t err = 0;

  for (;;) {
    __syncwarp();
    if (state == ST_BLOCK) {
      uint32_t b = 0;
      if (lig == 0) b = atomicAdd(queue, 1u);
      b = __shfl_sync(gmask, b, leader_lane);
      if (b >= n_blocks) {
        state = ST_DONE;
      } else {
        blk = b;
        BlockDesc d = blocks[b];
        obase = out + d.out_off;
        isize = d.isize;
        pos = 0;
        err = 0;
        bfinal = false;
        in_end = reinterpret_cast<const uint8_t*>(d.in_off) + d.clen;
        if (lig == 0) br.init(reinterpret_cast<const uint8_t*>(d.in_off));
        state = ST_HEADER;
      }
    }
    if (state == ST_HEADER) {
      // ---- deflate block header (leader), tables (group) ----
      uint32_t hdr = 0;  // bits 0-1 btype, 2 bfinal, 3.. hlit / hdist packed for the group
      if (lig == 0) {
        br.refill();
        uint32_t h = br.take(3);
        uint32_t btype = h >> 1;
        hdr = btype | ((h & 1) << 2);
        if (btype == 2) {
          br.refill();
          uint32_t v = br.take(14);
          uint32_t hlit = (v & 31) + 257, hdist = ((v >> 5) & 31) + 1, hclen = ((v >> 10) & 15) + 4;
          const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
          uint8_t* cl = s->lens + 320;
          for (int i = 0; i < 19; ++i) cl[i] = 0;
          for (uint32_t i = 0; i < hclen; ++i) {
            br.refill();
            cl[order[i]] = (uint8_t)br.take(3);
          }
          // code-length code: canonical, 7-bit LUT
          uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8];
          for (int i = 0; i < 19; ++i) cnt[cl[i]]++;
          cnt[0] = 0;
          uint32_t code = 0;
          for (int l = 1; l <= 7; ++l) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
          for (int i = 0; i < 32; ++i) reinterpret_cast<uint32_t*>(s->cl_lut)[i] = 0;
          for (int k = 0; k < 19; ++k) {
            uint32_t l = cl[k];
            if (!l) continue;
            uint32_t c = next[l]++;
            uint32_t rev = __brev(c) >> (32 - l);
            for (uint32_t j = rev; j < 128; j += (1u << l)) s->cl_lut[j] = (uint8_t)((k << 3) | l);
          }
          uint32_t total = hlit + hdist, i = 0;
          bool bad = hlit > 286 || hdist > 30;
          while (i < total && !bad) {
            br.refill();
            uint32_t e = s->cl_lut[br.peek32() & 127];
            uint32_t l = e & 7, sym = e >> 3;
            if (!l) { bad = true; break; }
            br.drop(l);
            if (sym < 16) { s->lens[i++] = (uint8_t)sym; continue; }
            uint32_t rep, val = 0;
            if (sym == 16) { if (i == 0) { bad = true; break; } val = s->lens[i - 1]; rep = 3 + br.take(2); }
            else if (sym == 17) rep = 3 + br.take(3);
            else rep = 11 + br.take(7);
            if (i + rep > total) { bad = true; break; }
            for (uint32_t k = 0; k < rep; ++k) s->lens[i++] = (uint8_t)val;
          }
          if (!bad && s->lens[256] == 0) bad = true;
          hdr |= (hlit << 3) | (hdist << 12);
          if (bad) hdr = 3;  // reserved btype = error
        }
      }
      hdr = __shfl_sync(gmask, hdr, leader_lane);
      uint32_t btype = hdr & 3;
      bfinal = (hdr >> 2) & 1;
      if (btype == 3) {
        err = kBlkBadStream;
        state = ST_FINISH;
      } else if (btype == 0) {
        // stored block: LEN/NLEN on the next byte boundary, then raw bytes
        uint32_t len = 0;
        unsigned long long src = 0;
        if (lig == 0) {
          br.drop(br.bc & 7);
          br.refill();
          uint32_t v = br.take(16);
          br.refill();
          uint32_t nv = br.take(16);
          if ((v ^ nv) != 0xFFFFu) len = 0xFFFFFFFFu; else len = v;
          src = (unsigned long long)br.byte_ptr();
        }
        len = __shfl_sync(gmask, len, leader_lane);
        src = __shfl_sync(gmask, src, leader_lane);
        if (len == 0xFFFFFFFFu) err = kBlkBadStream;
        else if (pos + len > isize) err = kBlkOverrun;
        else {
          const uint8_t* sp = reinterpret_cast<const uint8_t*>(src);
          for (uint32_t k = lig; k < len; k += G) obase[pos + k] = sp[k];
          pos += len;
          if (lig == 0) br.init(sp + len);
          __syncwarp(gmask);
        }
        state = (bfinal || err) ? ST_FINISH : ST_HEADER;
      } else {
        int n_ll, n_d;
        if (btype == 1) {
          for (int k = lig; k < 288; k += G) s->lens[k] = k < 144 ? 8 : k < 256 ? 9 : k < 280 ? 7 : 8;
          for (int k = lig; k < 32; k += G) s->lens[288 + k] = 5;
          n_ll = 288; n_d = 32;
        } else {
          n_ll = (hdr >> 3) & 511; n_d = (hdr >> 12) & 63;
        }
        __syncwarp(gmask);
        build_table<G>(s->lens, n_ll, s->ll_lut, kLLBits, s->ll_sorted, s->ll_count, s->ll_lim, s->ll_base, s, gmask, lig, lane);
        build_table<G>(s->lens + n_ll, n_d, s->d_lut, kDBits, s->d_sorted, s->d_count, s->d_lim, s->d_base, s, gmask, lig, lane);
        state = ST_DECODE;
      }
    }
    __syncwarp();
    // ---- phase A: every leader decodes up to kBatchIters symbols in lock step; literals are stored
    // at once, matches are queued as tokens (at most kTokens) ----
    uint32_t res = 0;  // bits 0-16 new pos, 17-20 tokens queued, 21-22 end kind (1 EOB, 2 bad stream, 3 overrun)
    if (lig == 0) {
      bool active = state == ST_DECODE;
      uint32_t p = pos, ntok = 0, endk = 0, first_dst = 0;
      for (int it = 0; it < kBatchIters; ++it) {
        if (active) {
          br.refill();
          uint32_t e = s->ll_lut[br.peek32() & ((1u << kLLBits) - 1u)];
          int len = e & 15;
          int sym = e >> 4;
          if (len == 0) sym = slow_decode<kLLBits>(br.peek32(), s->ll_lim, s->ll_base, s->ll_sorted, len);
          if (sym < 0) { endk = 2; active = false; }
          else {
            br.drop(len);
            if (sym < 256) {
              if (p >= isize) { endk = 3; active = false; }
              else obase[p++] = (uint8_t)sym;
            } else if (sym == 256) {
              endk = 1; active = false;
            } else {
              uint32_t idx = sym - 257, mlen, dist = 0;
              if (idx < 8) mlen = 3 + idx;
              else if (idx >= 28) mlen = 258;
              else {
                uint32_t eb = (idx - 4) >> 2;
                mlen = 3 + ((4 + (idx & 3)) << eb) + br.take(eb);
              }
              br.refill();
              uint32_t de = s->d_lut[br.peek32() & ((1u << kDBits) - 1u)];
              int dl = de & 15;
              int ds = de >> 4;
              if (dl == 0) ds = slow_decode<kDBits>(br.peek32(), s->d_lim, s->d_base, s->d_sorted, dl);
              if (ds < 0 || ds > 29 || idx > 28) { endk = 2; active = false; }
              else {
                br.drop(dl);
                if (ds < 4) dist = 1 + ds;
                else {
                  uint32_t eb = (ds >> 1) - 1;
                  dist = 1 + ((2 + (ds & 1)) << eb) + br.take(eb);
                }
                if (dist > p || p + mlen > isize) { endk = 3; active = false; }
                else {
                  // a token depends on this batch when its source reaches into an earlier queued match
                  uint32_t dep = (ntok && p - dist + mlen > first_dst) ? 0x80000000u : 0u;
                  if (!ntok) first_dst = p;
                  s->tok_pl[ntok] = p | ((mlen - 3) << 16);
                  s->tok_d[ntok] = dist | dep;
                  ++ntok;
                  p += mlen;
                  if (ntok == kTokens) active = false;
                }
              }
            }
          }
        }
        if (!__any_sync(kLeaders, active)) break;
      }
      // a malformed stream must not run away over the input: the reader may be at most its prefetch ahead
      if (state == ST_DECODE && !endk && reinterpret_cast<const uint8_t*>(br.wptr) > in_end + 24) endk = 2;
      res = p | (ntok << 17) | (endk << 21);
    }
    res = __shfl_sync(gmask, res, leader_lane);
    __syncwarp();  // literal stores and tokens are visible to every lane
    // ---- phase B: copies.  B1: one lane per independent token; B2: dependent tokens in order ----
    {
      const uint32_t ntok = (res >> 17) & 15;
      for (uint32_t j = lig; j < ntok; j += G) {
        uint32_t d = s->tok_d[j];
        if (!(d & 0x80000000u)) {
          uint32_t pl = s->tok_pl[j];
          lane_copy(obase + (pl & 0xFFFF), (pl >> 16) + 3, d);
        }
      }
      __syncwarp();
      for (uint32_t j = 1; j < ntok; ++j) {
        uint32_t d = s->tok_d[j];
        if (d & 0x80000000u) {
          d &= 0x7FFFFFFFu;
          uint32_t pl = s->tok_pl[j];
          uint32_t mlen = (pl >> 16) + 3;
          uint8_t* dst = obase + (pl & 0xFFFF);
          const uint8_t* src = dst - d;
          if (d >= mlen) {
            for (uint32_t k = lig; k < mlen; k += G) dst[k] = src[k];
          } else if (d == 1) {
            uint8_t v = src[0];
            for (uint32_t k = lig; k < mlen; k += G) dst[k] = v;
          } else {
            for (uint32_t k = lig; k < mlen; k += G) dst[k] = src[k % d];
          }
          __syncwarp(gmask);
        }
      }
      if (state == ST_DECODE) {
        pos = res & 0x1FFFF;
        uint32_t endk = (res >> 21) & 3;
        if (endk == 1) state = bfinal ? ST_FINISH : ST_HEADER;
        else if (endk == 2) { err = kBlkBadStream; state = ST_FINISH; }
        else if (endk == 3) { err = kBlkOverrun; state = ST_FINISH; }
      }
    }
    if (state == ST_FINISH) {
      if (!err && pos != isize) err = kBlkIsize;
      if (lig == 0 && err) status[blk] = err;
      state = ST_BLOCK;
    }
    if (__all_sync(0xFFFFFFFFu, state == ST_DONE)) break;
  }
}

}  // namespace ngsq
