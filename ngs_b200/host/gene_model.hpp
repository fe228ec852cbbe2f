// Host-side inputs of the two optional facets (reference: src/utils/formats/gff.rs + the GFF walk of
// src/qc/record_based/features.rs:270-355; src/utils/formats/fasta.rs + EditsFacet::setup, edits.rs:175-215).
// Plain-text or gzip files (zlib's gzFile reads both), parsed the way noodles' readers are used there.
#pragma once
#include <zlib.h>

#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "genome.hpp"

namespace ngs {

inline std::string slurp_maybe_gz(const std::string& path, const char* what) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error(std::string("opening ") + what + " file: " + path);
  std::string out;
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof buf)) > 0) out.append(buf, (size_t)n);
  const bool bad = n < 0;
  gzclose(f);
  if (bad) throw std::runtime_error(std::string("reading ") + what + " file: " + path);
  return out;
}

// command.rs:78-101 — slots 0 five_prime_utr, 1 three_prime_utr, 2 coding_sequence, 3 exon, 4 gene
struct FeatureNames {
  std::array<std::string, 5> slot{"five_prime_UTR", "three_prime_UTR", "CDS", "exon", "gene"};
  // smallest slot index whose name equals slot j's (names may coincide)
  std::array<uint8_t, 5> slot_class() const {
    std::array<uint8_t, 5> c{};
    for (int j = 0; j < 5; ++j) { c[j] = (uint8_t)j; for (int i = j - 1; i >= 0; --i) if (slot[i] == slot[j]) c[j] = (uint8_t)i; }
    return c;
  }
};

struct ContigFeatures { std::vector<uint32_t> start, stop; std::vector<uint8_t> cls; };

// features.rs:287-345: every GFF record is read (a malformed one aborts: `result.unwrap()`); records on a sequence of
// the genome's primary assembly get their strand parsed ("+" / "-" only, features/utils.rs:33-43) BEFORE their type is
// looked at; those whose type is one of the five names are kept, filed under the FIRST slot with that name.
inline std::map<std::string, ContigFeatures> read_gene_model(const std::string& gff_text, const FeatureNames& names, const ReferenceGenome& genome) {
  std::map<std::string, ContigFeatures> out;
  std::vector<Sequence> primary = get_primary_assembly(genome);
  auto is_primary = [&](const std::string& s) { for (auto& p : primary) if (p.name == s) return true; return false; };
  size_t i = 0;
  while (i < gff_text.size()) {
    size_t eol = gff_text.find('\n', i);
    if (eol == std::string::npos) eol = gff_text.size();
    size_t n = eol - i;
    if (n && gff_text[i + n - 1] == '\r') --n;
    const std::string line = gff_text.substr(i, n);
    i = eol + 1;
    if (line.rfind("##FASTA", 0) == 0) break;  // records() ends at the FASTA section
    if (!line.empty() && line[0] == '#') continue;
    std::vector<std::string> f;
    size_t b = 0;
    for (size_t q = 0; q <= line.size(); ++q) if (q == line.size() || line[q] == '\t') { f.push_back(line.substr(b, q - b)); b = q + 1; }
    auto position = [](const std::string& s, uint32_t* v) {
      if (s.empty() || s.size() > 10 || s.find_first_not_of("0123456789") != std::string::npos) return false;
      unsigned long long x = std::stoull(s);
      if (x == 0 || x > 0xFFFFFFFFull) return false;
      *v = (uint32_t)x;
      return true;
    };
    uint32_t start = 0, stop = 0;
    if (f.size() != 9 || !position(f[3], &start) || !position(f[4], &stop)) throw std::runtime_error("invalid GFF record: " + line);
    if (!is_primary(f[0])) continue;
    if (f[6] != "+" && f[6] != "-") throw std::runtime_error("attempted to parse strand from value: " + f[6]);
    for (int j = 0; j < 5; ++j)
      if (f[2] == names.slot[j]) {
        ContigFeatures& c = out[f[0]];
        c.start.push_back(start); c.stop.push_back(stop); c.cls.push_back((uint8_t)j);
        break;
      }
  }
  return out;
}

// fasta records: name = first word after '>', sequence = the following lines joined (line ends removed)
inline std::map<std::string, std::string> read_fasta(const std::string& text) {
  std::map<std::string, std::string> out;
  std::string* cur = nullptr;
  size_t i = 0;
  while (i < text.size()) {
    size_t eol = text.find('\n', i);
    if (eol == std::string::npos) eol = text.size();
    size_t n = eol - i;
    if (n && text[i + n - 1] == '\r') --n;
    if (n && text[i] == '>') {
      size_t e = i + 1;
      while (e < i + n && text[e] != ' ' && text[e] != '\t') ++e;
      cur = &out[text.substr(i + 1, e - i - 1)];
      cur->clear();
    } else if (cur) {
      cur->append(text, i, n);
    }
    i = eol + 1;
  }
  return out;
}

}  // namespace ngs
