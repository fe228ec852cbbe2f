// TEST DOUBLE of the integer getters of include/ngs_cuda.h (tests/ only, never linked into the product):
// serves the integers stored in the file named by $NGSQ_FAKE_INTS (written by tests/test_host_results_cpu.py
// from the oracle), so that the product's host-side facets (ngs_b200/host/facets.hpp: summarize / teardown /
// aggregate) and its JSON writer (results.hpp) can be checked on a machine without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ngs_cuda.h"

namespace {
struct Contig { uint64_t touched, n_bins, too_large; std::vector<uint64_t> hist, bin_sums; };
struct Ints {
  std::vector<uint64_t> general, tlen, gc, nuc, rec, quality;
  uint64_t processed = 0, ignored = 0, n_pos = 0, nonsensical = 0;
  std::vector<Contig> contigs;
  bool loaded = false;
} g;

std::vector<uint64_t> rd(FILE* f, size_t n) {
  std::vector<uint64_t> v(n);
  if (n && fread(v.data(), 8, n, f) != n) { fprintf(stderr, "fake engine: short ints file\n"); exit(2); }
  return v;
}
void load() {
  if (g.loaded) return;
  const char* p = getenv("NGSQ_FAKE_INTS");
  FILE* f = p ? fopen(p, "rb") : nullptr;
  if (!f) { fprintf(stderr, "fake engine: NGSQ_FAKE_INTS not readable\n"); exit(2); }
  g.general = rd(f, 34); g.tlen = rd(f, 1025); g.processed = rd(f, 1)[0]; g.ignored = rd(f, 1)[0];
  g.gc = rd(f, 101); g.nuc = rd(f, 3); g.rec = rd(f, 3);
  g.n_pos = rd(f, 1)[0]; g.quality = rd(f, g.n_pos * 94);
  g.nonsensical = rd(f, 1)[0];
  uint64_t n_ref = rd(f, 1)[0];
  for (uint64_t c = 0; c < n_ref; ++c) {
    Contig k;
    auto h = rd(f, 3);
    k.touched = h[0]; k.n_bins = h[1]; k.too_large = h[2];
    k.hist = rd(f, 2049); k.bin_sums = rd(f, k.n_bins);
    g.contigs.push_back(std::move(k));
  }
  fclose(f);
  g.loaded = true;
}
}  // namespace

extern "C" {
const char* ngsq_last_error(ngsq_engine*) { return "fake engine"; }
int ngsq_get_general(ngsq_engine*, uint64_t out[34]) { load(); memcpy(out, g.general.data(), 34 * 8); return 0; }
int ngsq_get_tlen(ngsq_engine*, uint64_t hist[1025], uint64_t* processed, uint64_t* ignored) {
  load(); memcpy(hist, g.tlen.data(), 1025 * 8); *processed = g.processed; *ignored = g.ignored; return 0;
}
int ngsq_get_gc(ngsq_engine*, uint64_t hist[101], uint64_t nuc[3], uint64_t rec[3]) {
  load(); memcpy(hist, g.gc.data(), 101 * 8); memcpy(nuc, g.nuc.data(), 24); memcpy(rec, g.rec.data(), 24); return 0;
}
int ngsq_get_quality(ngsq_engine*, uint64_t* out, size_t cap_positions, uint32_t* n_positions) {
  load();
  if (n_positions) *n_positions = (uint32_t)g.n_pos;
  if (!out) return 0;
  if (cap_positions < g.n_pos) return NGSQ_E_ARG;
  memcpy(out, g.quality.data(), g.n_pos * 94 * 8);
  return 0;
}
int ngsq_get_coverage_contig(ngsq_engine*, uint32_t ref, ngsq_cov_ints* out, uint64_t* bin_sums, size_t cap) {
  load();
  memset(out, 0, sizeof *out);
  if (ref >= g.contigs.size()) return NGSQ_E_ARG;
  const Contig& k = g.contigs[ref];
  out->touched = (uint32_t)k.touched;
  if (!k.touched) return 0;
  out->n_bins = (uint32_t)k.n_bins;
  out->pileup_too_large = k.too_large;
  memcpy(out->hist, k.hist.data(), 2049 * 8);
  if (bin_sums) { if (cap < k.n_bins) return NGSQ_E_ARG; memcpy(bin_sums, k.bin_sums.data(), k.n_bins * 8); }
  return 0;
}
int ngsq_get_coverage_global(ngsq_engine*, uint64_t* nonsensical_records) { load(); *nonsensical_records = g.nonsensical; return 0; }
}

// ---- the two "next" facets: integers from $NGSQ_FAKE_NEXT = 9 feature counters, read_one[513], read_two[513], vaf[101], records
namespace {
std::vector<uint64_t> g_next;
void load_next() {
  if (!g_next.empty()) return;
  const char* p = getenv("NGSQ_FAKE_NEXT");
  FILE* f = p ? fopen(p, "rb") : nullptr;
  if (!f) { fprintf(stderr, "fake engine: NGSQ_FAKE_NEXT not readable\n"); exit(2); }
  g_next = rd(f, 9 + 513 + 513 + 101 + 1);
  fclose(f);
}
}  // namespace
extern "C" {
int ngsq_get_features(ngsq_engine*, uint64_t counts[9]) { load_next(); memcpy(counts, g_next.data(), 72); return 0; }
int ngsq_get_edits(ngsq_engine*, uint64_t read_one[513], uint64_t read_two[513], uint64_t vaf[101], uint64_t* records) {
  load_next();
  memcpy(read_one, &g_next[9], 513 * 8); memcpy(read_two, &g_next[9 + 513], 513 * 8); memcpy(vaf, &g_next[9 + 1026], 101 * 8);
  if (records) *records = g_next[9 + 1127];
  return 0;
}
// per-position counters from $NGSQ_FAKE_VAF: per reference id, in order, [u32 n][n x u32 refs][n x u32 alts]
int ngsq_get_edit_positions(ngsq_engine*, uint32_t ref, uint32_t* refs, uint32_t* alts, uint64_t n) {
  const char* p = getenv("NGSQ_FAKE_VAF");
  FILE* f = p ? fopen(p, "rb") : nullptr;
  if (!f) { fprintf(stderr, "fake engine: NGSQ_FAKE_VAF not readable\n"); exit(2); }
  for (uint32_t c = 0;; ++c) {
    uint32_t m = 0;
    if (fread(&m, 4, 1, f) != 1) { fprintf(stderr, "fake engine: no counters for reference %u\n", ref); exit(2); }
    std::vector<uint32_t> r(m), a(m);
    if (m && (fread(r.data(), 4, m, f) != m || fread(a.data(), 4, m, f) != m)) { fprintf(stderr, "fake engine: short VAF file\n"); exit(2); }
    if (c == ref) {
      if (m != n) { fclose(f); return -1; }
      memcpy(refs, r.data(), (size_t)m * 4); memcpy(alts, a.data(), (size_t)m * 4);
      fclose(f);
      return 0;
    }
  }
}
int ngsq_set_feature_model(ngsq_engine*, const uint8_t*, const uint8_t*) { return 0; }
int ngsq_set_features(ngsq_engine*, uint32_t, uint32_t, const uint32_t*, const uint32_t*, const uint8_t*) { return 0; }
int ngsq_set_reference_bases(ngsq_engine*, uint32_t, const uint8_t*, uint64_t) { return 0; }
}
