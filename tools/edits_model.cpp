// CPU model of the Edits kernel (test tooling; NOT part of the product library): runs the per-record function the
// CUDA kernel calls per lane (ngs_b200/csrc/edits.cuh: edits_record, edits_encode, edits_vaf_bin) over every record
// of a BAM against a FASTA and prints the facet's integers, for comparison with the oracle.
//   g++ -O2 -std=c++17 -o /tmp/edits_model tools/edits_model.cpp -lz
//   /tmp/edits_model file.bam ref.fa      -> "records N" / "read_one ..." / "read_two ..." / "vaf ..."  or  "error <code>"
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../ngs_b200/csrc/edits.cuh"

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

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s file.bam ref.fa\n", argv[0]); return 2; }
  std::vector<uint8_t> bam = slurp(argv[1]), fa = slurp(argv[2]);
  // inflate every BGZF member
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
  // header
  uint32_t l_text, n_ref;
  memcpy(&l_text, &s[4], 4);
  size_t p = 8 + l_text;
  memcpy(&n_ref, &s[p], 4);
  p += 4;
  std::vector<std::string> names(n_ref);
  std::vector<EditsContig> contigs(n_ref);
  std::vector<std::vector<uint8_t>> codes(n_ref);
  std::vector<std::vector<uint32_t>> bits(n_ref), prefix(n_ref);
  uint64_t pos_total = 0;
  for (uint32_t c = 0; c < n_ref; ++c) {
    uint32_t ln, L;
    memcpy(&ln, &s[p], 4);
    names[c].assign((const char*)&s[p + 4], ln - 1);
    memcpy(&L, &s[p + 4 + ln], 4);
    p += 8 + ln;
    EditsContig& C = contigs[c];
    memset(&C, 0, sizeof C);
    C.hdr_len = L;
    C.pos_off = pos_total;
    pos_total += (uint64_t)L + 1;
    // FASTA record whose name (first word after '>') equals the contig's
    std::vector<uint8_t> letters;
    bool hit = false;
    for (size_t i = 0; i < fa.size();) {
      size_t eol = i;
      while (eol < fa.size() && fa[eol] != '\n') ++eol;
      if (fa[i] == '>') {
        if (hit) break;
        size_t e = i + 1;
        while (e < eol && fa[e] != ' ' && fa[e] != '\t' && fa[e] != '\r') ++e;
        hit = names[c] == std::string((const char*)&fa[i + 1], e - i - 1);
      } else if (hit) {
        for (size_t q = i; q < eol; ++q) if (fa[q] != '\r') letters.push_back(fa[q]);
      }
      i = eol + 1;
    }
    if (!hit) continue;
    C.code_len = letters.size();
    codes[c].resize(letters.size() + 1);
    bits[c].resize(letters.size() / 32 + 1);
    prefix[c].resize(letters.size() / 32 + 1);
    edits_encode(letters.data(), letters.size(), codes[c].data(), bits[c].data(), prefix[c].data());
    C.codes = codes[c].data();
    C.bad_bits = bits[c].data();
    C.bad_prefix = prefix[c].data();
  }
  std::vector<uint32_t> refs(pos_total, 0), alts(pos_total, 0);
  std::vector<uint64_t> res(E_WORDS, 0);
  uint32_t err = 0;
  while (p + 36 <= s.size()) {
    uint32_t bs;
    memcpy(&bs, &s[p], 4);
    const uint8_t* rec = &s[p];
    const int32_t ref = (int32_t)ed_ld32(rec + 4);
    const uint64_t poff = (ref >= 0 && ref < (int32_t)n_ref) ? contigs[ref].pos_off : 0;
    uint32_t e = 0;
    bool first = false;
    const uint32_t st = edits_record(rec, (int32_t)n_ref, contigs.data(), &e, &first,
                                     [&](uint64_t q, bool is_edit) { (is_edit ? alts : refs)[poff + q]++; });
    if (st == kEdCounted) { res[(first ? E_READ_ONE : E_READ_TWO) + e]++; res[E_RECORDS]++; }
    else if (st != kEdSkipped) err = err > st ? err : st;
    p += 4 + (size_t)bs;
  }
  if (err) { printf("error %u\n", err); return 0; }
  for (uint64_t i = 0; i < pos_total; ++i)
    if (refs[i] + alts[i]) res[E_VAF + edits_vaf_bin(refs[i], alts[i])]++;
  printf("records %llu\n", (unsigned long long)res[E_RECORDS]);
  const char* nm[3] = {"read_one", "read_two", "vaf"};
  const uint32_t off[3] = {E_READ_ONE, E_READ_TWO, E_VAF}, cnt[3] = {513, 513, 101};
  for (int k = 0; k < 3; ++k) {
    printf("%s", nm[k]);
    for (uint32_t i = 0; i < cnt[k]; ++i) printf(" %llu", (unsigned long long)res[off[k] + i]);
    printf("\n");
  }
  return 0;
}
