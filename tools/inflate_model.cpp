// CPU model of the inflate kernel (test tooling; NOT part of the product library).
// Compiles ngs_b200/csrc/inflate_lane.cuh — the per-lane decoder with the fused LZ77 copies that inflate_kernel runs — for
// the host, drives it exactly as the kernel does (header / step / settle) over every BGZF block of a file at several
// output alignments, checks the bytes against zlib and prints the symbol statistics that drive kernel design decisions.
//   g++ -O2 -std=c++17 -DNGSQ_HOST_MODEL -o /tmp/inflate_model tools/inflate_model.cpp -lz
//   /tmp/inflate_model file.bam [max_blocks]
//   /tmp/inflate_model file.bam --fuzz N     N corrupted copies of every block: the decoder must fail
//                                            or finish, never write outside its block, never run away
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../ngs_b200/csrc/inflate_lane.cuh"

using namespace ngsq;

static uint32_t g_lut[64];

// one block through the lane code; returns the number of loop iterations (0xFFFFFFFF: did not terminate)
static uint64_t run_block(const BlockDesc& d, uint8_t* out, uint8_t* slab, InflateCounters* ctr, uint32_t* err, uint64_t max_steps) {
  Lane L;
  L.slab = slab;
  L.lut = g_lut;
  L.ctr = ctr;
  L.begin_block(d, out);
  uint64_t steps = 0;
  while (L.state != LS_IDLE && steps < max_steps) {
    if (L.state == LS_HEADER) L.header();
    else {
      L.step();
      L.settle();
      if (L.state == LS_DECODE && L.overran()) L.end_block(kBlkBadStream);
    }
    ++steps;
  }
  *err = L.err;
  return L.state == LS_IDLE ? steps : ~0ull;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s file.bam [max_blocks]\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  fseek(f, 0, SEEK_END);
  size_t n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(n + 1024, 0);
  if (fread(buf.data(), 1, n, f) != n) { perror("read"); return 2; }
  fclose(f);
  size_t max_blocks = argc > 2 ? strtoull(argv[2], nullptr, 10) : ~size_t(0);
  int fuzz = 0;
  if (argc > 3 && !strcmp(argv[2], "--fuzz")) { fuzz = atoi(argv[3]); max_blocks = ~size_t(0); }
  uint64_t fuzz_runs = 0, fuzz_failed = 0, fuzz_ok = 0, rng = 0x9E3779B97F4A7C15ull;

  InflateCounters ctr;
  for (uint32_t i = 0; i < 64; ++i) g_lut[i] = base_lut_entry(i);
  std::vector<uint8_t> slab(kSlabBytes);
  constexpr size_t kPad = 64;
  std::vector<uint8_t> out(kPad + 65536 + kPad), ref(65536);
  uint64_t n_blocks = 0, total_out = 0, total_in = 0, bad = 0, iterations = 0;
  size_t o = 0;
  while (o + 18 <= n && n_blocks < max_blocks) {
    const uint8_t* h = &buf[o];
    if (h[0] != 0x1f || h[1] != 0x8b) { fprintf(stderr, "bad magic at %zu\n", o); return 1; }
    uint32_t xlen = h[10] | (h[11] << 8);
    uint32_t bsize = h[16] | (h[17] << 8);
    size_t total = (size_t)bsize + 1;
    uint32_t isize;
    memcpy(&isize, h + total - 4, 4);
    uint32_t crc;
    memcpy(&crc, h + total - 8, 4);
    uint32_t hdr = 12 + xlen;
    uint32_t clen = (uint32_t)total - hdr - 8;
    for (int fz = 0; fz < fuzz && isize; ++fz) {
      // corrupt 1-3 bytes of a private copy of the payload (with the reader's look-ahead padding)
      std::vector<uint8_t> pay(clen + 1024, 0);
      memcpy(pay.data(), h + hdr, clen);
      int nflip = 1 + (int)(rng >> 62) % 3;
      for (int k = 0; k < nflip; ++k) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        pay[(rng >> 33) % clen] ^= (uint8_t)(1u << ((rng >> 20) & 7));
      }
      BlockDesc d{};
      d.in_off = (uint64_t)(uintptr_t)pay.data();
      d.out_off = kPad - 16 + (fz % 19);
      d.clen = clen;
      d.isize = isize;
      memset(out.data(), 0xAA, out.size());
      uint32_t err = 0;
      const uint64_t steps = run_block(d, out.data(), slab.data(), nullptr, &err, 4000000);
      if (steps == ~0ull) { fprintf(stderr, "fuzz: block at %zu did not terminate\n", o); bad++; }
      uint8_t* ob = out.data() + d.out_off;
      // the neighbouring blocks' bytes: untouched whatever the stream said
      for (size_t g = 0; g < d.out_off; ++g)
        if (out[g] != 0xAA) { fprintf(stderr, "fuzz: block at %zu wrote %zu bytes before its range\n", o, d.out_off - g); bad++; break; }
      for (size_t g = d.out_off + isize; g < out.size(); ++g)
        if (out[g] != 0xAA) { fprintf(stderr, "fuzz: block at %zu wrote %zu bytes behind its range\n", o, g - d.out_off - isize); bad++; break; }
      (void)ob;
      if (!err) fuzz_ok++; else fuzz_failed++;
      fuzz_runs++;
    }
    if (isize && !fuzz) {
      // zlib
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      inflateInit2(&zs, -15);
      zs.next_in = const_cast<uint8_t*>(h + hdr);
      zs.avail_in = clen;
      zs.next_out = ref.data();
      zs.avail_out = 65536;
      int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END || zs.total_out != isize) { fprintf(stderr, "zlib failed at %zu\n", o); return 1; }
      if (crc32(0, ref.data(), isize) != crc) { fprintf(stderr, "crc mismatch at %zu\n", o); bad++; }
      // every alignment of the block start within a 16-byte chunk exercises the edge chunks and the piece loads
      const int aligns[] = {0, 1, 3, 7, 8, 13, 15};
      for (int ai = 0; ai < 7; ++ai) {
        const int mis = aligns[(ai + n_blocks) % 7];
        if (ai >= 3 && n_blocks >= 64) break;  // all seven on the first blocks, three per block afterwards
        BlockDesc d{};
        d.in_off = (uint64_t)(uintptr_t)(h + hdr);
        d.out_off = kPad - 16 + mis;
        d.clen = clen;
        d.isize = isize;
        memset(out.data(), 0xAA, out.size());
        uint32_t err = 0;
        const uint64_t steps = run_block(d, out.data(), slab.data(), ai == 0 ? &ctr : nullptr, &err, ~0ull >> 1);
        if (ai == 0) iterations += steps;
        if (err) { fprintf(stderr, "block at %zu: err %u\n", o, err); bad++; break; }
        uint8_t* ob = out.data() + d.out_off;
        for (size_t g = 0; g < d.out_off; ++g)
          if (out[g] != 0xAA) { fprintf(stderr, "block at %zu: wrote before its range (mis %d)\n", o, mis); bad++; break; }
        for (size_t g = d.out_off + isize; g < out.size(); ++g)
          if (out[g] != 0xAA) { fprintf(stderr, "block at %zu: wrote behind its range (mis %d)\n", o, mis); bad++; break; }
        if (memcmp(ref.data(), ob, isize) != 0) {
          uint32_t k = 0;
          while (ref[k] == ob[k]) ++k;
          fprintf(stderr, "block at %zu (mis %d): MISMATCH at byte %u of %u\n", o, mis, k, isize);
          bad++;
        }
      }
      n_blocks++;
      total_out += isize;
      total_in += clen;
    }
    o += total;
  }
  if (fuzz) {
    printf("fuzz: %llu corrupted blocks, %llu rejected, %llu decoded, bad %llu\n", (unsigned long long)fuzz_runs,
           (unsigned long long)fuzz_failed, (unsigned long long)fuzz_ok, (unsigned long long)bad);
    return bad ? 1 : 0;
  }
  printf("blocks %llu, in %llu, out %llu (ratio %.2f), bad %llu\n", (unsigned long long)n_blocks, (unsigned long long)total_in,
         (unsigned long long)total_out, total_in ? (double)total_out / total_in : 0.0, (unsigned long long)bad);
  if (!ctr.symbols) return bad ? 1 : 0;
  printf("symbols %llu (%.3f per out byte): literals %llu (%.1f%%), matches %llu (%.1f%%), avg match len %.2f, match bytes %.1f%% of output\n",
         (unsigned long long)ctr.symbols, (double)ctr.symbols / total_out, (unsigned long long)ctr.literals, 100.0 * ctr.literals / ctr.symbols,
         (unsigned long long)ctr.matches, 100.0 * ctr.matches / ctr.symbols, (double)ctr.match_bytes / (ctr.matches ? ctr.matches : 1),
         100.0 * ctr.match_bytes / total_out);
  printf("bits per symbol %.2f\n", 8.0 * total_in / ctr.symbols);
  printf("ll code length %%:");
  for (int l = 1; l <= 15; ++l) printf(" %d:%.1f", l, 100.0 * ctr.ll_len_hist[l] / ctr.symbols);
  printf("\ndist code length %%:");
  for (int l = 1; l <= 15; ++l) printf(" %d:%.1f", l, 100.0 * ctr.d_len_hist[l] / (ctr.matches ? ctr.matches : 1));
  printf("\n");
  printf("deflate blocks per BGZF block %.2f (stored %llu, fixed %llu); symbols per deflate block %.0f\n", (double)ctr.headers / n_blocks,
         (unsigned long long)ctr.stored, (unsigned long long)ctr.fixed, (double)ctr.symbols / ctr.headers);
  uint64_t le4 = 0, le8 = 0, le16 = 0;
  for (int l = 3; l <= 258; ++l) { if (l <= 4) le4 += ctr.match_len_hist[l]; if (l <= 8) le8 += ctr.match_len_hist[l]; if (l <= 16) le16 += ctr.match_len_hist[l]; }
  const double nm = ctr.matches ? (double)ctr.matches : 1.0;
  printf("match length: <= 4 %.1f%%, <= 8 %.1f%%, <= 16 %.1f%%\n", 100 * le4 / nm, 100 * le8 / nm, 100 * le16 / nm);
  printf("copy: %.3f pieces per match, %.1f%% of the loop iterations only copy, open chunk stored before %.1f%% of the pieces\n",
         ctr.pieces / nm, 100.0 * ctr.copy_iterations / (iterations ? iterations : 1), 100.0 * ctr.open_chunk_stores / (ctr.pieces ? ctr.pieces : 1));
  printf("stores: %llu chunk + %llu edge = %.3f per output byte; loop iterations %.3f per symbol\n", (unsigned long long)ctr.chunk_stores,
         (unsigned long long)ctr.edge_stores, (double)(ctr.chunk_stores + ctr.edge_stores) / total_out, (double)iterations / ctr.symbols);
  return bad ? 1 : 0;
}
