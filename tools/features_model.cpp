// CPU model of the Genomic Features kernel (test tooling; NOT part of the product library): runs the per-record
// function the CUDA kernel calls per lane (ngs_b200/csrc/features.cuh: features_record) over every record of a BAM
// against a GFF, with the per-class sorted start / stop arrays built the way the engine builds them.
//   g++ -O2 -std=c++17 -o /tmp/features_model tools/features_model.cpp -lz
//   /tmp/features_model file.bam model.gff 5UTR 3UTR CDS exon gene [n_records]   -> nine counters, or "error <kind>"
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../ngs_b200/csrc/features.cuh"

using namespace ngsq;

static std::vector<uint8_t> slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b(n);
  if (n && fread(b.data(), 1, n, f) != (size_t)n) { perror("read"); exit(2); }
  fclose(f);
  return b;
}

static bool is_primary(const std::string& n) {  // utils/genome.rs:59-83 over grch38_no_alt.rs:46-283, as a name rule
  if (n.rfind("chr", 0) != 0) return false;
  const std::string s = n.substr(3);
  if (s == "X" || s == "Y") return true;
  if (!s.empty() && s[0] != '0' && s.find_first_not_of("0123456789") == std::string::npos && s.size() <= 2) { int v = atoi(s.c_str()); if (v >= 1 && v <= 22) return true; }
  if (n.rfind("chrUn_", 0) == 0) return true;
  return n.size() > 7 && n.compare(n.size() - 7, 7, "_random") == 0;
}

int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "usage: %s file.bam model.gff 5UTR 3UTR CDS exon gene [n_records]\n", argv[0]); return 2; }
  std::vector<uint8_t> bam = slurp(argv[1]), gff = slurp(argv[2]);
  const std::string names[5] = {argv[3], argv[4], argv[5], argv[6], argv[7]};
  const uint64_t max_records = argc > 8 ? strtoull(argv[8], nullptr, 10) : 0;
  uint8_t slot_class[8] = {0};
  for (int j = 0; j < 5; ++j) { int c = j; for (int i = j - 1; i >= 0; --i) if (names[i] == names[j]) c = i; slot_class[j] = (uint8_t)c; }
  std::vector<uint8_t> s;
  for (size_t o = 0; o + 18 <= bam.size();) {
    const size_t total = (size_t)(bam[o + 16] | (bam[o + 17] << 8)) + 1;
    uint32_t isize;
    memcpy(&isize, &bam[o + total - 4], 4);
    const size_t at = s.size();
    s.resize(at + isize);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    inflateInit2(&zs, -15);
    zs.next_in = &bam[o + 18];
    zs.avail_in = (uInt)(total - 26);
    zs.next_out = s.data() + at;
    zs.avail_out = isize;
    if (isize && inflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "inflate failed\n"); return 2; }
    inflateEnd(&zs);
    o += total;
  }
  uint32_t l_text, n_ref;
  memcpy(&l_text, &s[4], 4);
  size_t p = 8 + l_text;
  memcpy(&n_ref, &s[p], 4);
  p += 4;
  std::vector<std::string> refs(n_ref);
  for (uint32_t c = 0; c < n_ref; ++c) {
    uint32_t ln;
    memcpy(&ln, &s[p], 4);
    refs[c].assign((const char*)&s[p + 4], ln - 1);
    p += 8 + ln;
  }
  // the caller's part of ngsq_set_features: class = smallest slot whose name equals the type; per class sorted starts / stops
  std::vector<std::vector<uint32_t>> starts(n_ref * 5), stops(n_ref * 5);
  for (size_t i = 0; i < gff.size();) {
    size_t eol = i;
    while (eol < gff.size() && gff[eol] != '\n') ++eol;
    if (eol > i && gff[i] != '#') {
      std::vector<std::string> f;
      size_t b = i;
      for (size_t q = i; q <= eol; ++q) if (q == eol || gff[q] == '\t') { f.emplace_back((const char*)&gff[b], q - b); b = q + 1; }
      if (f.size() >= 9) {
        int cls = -1;
        for (int j = 0; j < 5; ++j) if (f[2] == names[j]) { cls = j; break; }
        for (uint32_t c = 0; c < n_ref && cls >= 0; ++c)
          if (refs[c] == f[0] && is_primary(refs[c])) { starts[c * 5 + cls].push_back((uint32_t)strtoul(f[3].c_str(), nullptr, 10)); stops[c * 5 + cls].push_back((uint32_t)strtoul(f[4].c_str(), nullptr, 10)); }
      }
    }
    i = eol + 1;
  }
  std::vector<FeatureContig> contigs(n_ref);
  for (uint32_t c = 0; c < n_ref; ++c) {
    memset(&contigs[c], 0, sizeof(FeatureContig));
    contigs[c].primary = is_primary(refs[c]);
    for (int k = 0; k < 5; ++k) {
      std::sort(starts[c * 5 + k].begin(), starts[c * 5 + k].end());
      std::sort(stops[c * 5 + k].begin(), stops[c * 5 + k].end());
      contigs[c].n[k] = (uint32_t)starts[c * 5 + k].size();
      contigs[c].starts[k] = starts[c * 5 + k].data();
      contigs[c].stops[k] = stops[c * 5 + k].data();
    }
  }
  uint64_t res[F_WORDS] = {0}, r = 0;
  uint32_t err = 0;
  while (p + 36 <= s.size() && (max_records == 0 || r < max_records)) {
    uint32_t bs;
    memcpy(&bs, &s[p], 4);
    uint32_t bits = 0;
    const uint32_t st = features_record(&s[p], (int32_t)n_ref, contigs.data(), slot_class, &bits);
    if (st) err = err > st ? err : st;
    else for (int k = 0; k < 9; ++k) res[k] += (bits >> k) & 1u;
    p += 4 + (size_t)bs;
    ++r;
  }
  if (err) { printf("error %u\n", err); return 0; }
  for (int k = 0; k < 9; ++k) printf("%llu%c", (unsigned long long)res[k], k == 8 ? '\n' : ' ');
  return 0;
}
