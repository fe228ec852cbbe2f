// Built-in reference-genome name tables (reference: src/utils/genome.rs:32-126 and
// src/utils/genome/ncbi/grch38_no_alt.rs:46-283).  Only GRCh38_no_alt_AnalysisSet is built in
// without the reference's `extended-reference-genomes` cargo feature (genome.rs:32-44).
#pragma once
#include <algorithm>
#include <cctype>
#include <memory>
#include <string>
#include <vector>

namespace ngs {

struct Sequence {
  std::string name;
  char kind;  // C chromosome, M mitochondrion, E ebv, L unlocalized, P unplaced
};

class ReferenceGenome {
 public:
  virtual ~ReferenceGenome() = default;
  virtual const char* name() const = 0;
  virtual const std::vector<Sequence>& sequences() const = 0;
};

class GRCh38NoAltAnalysisSet : public ReferenceGenome {
 public:
  const char* name() const override { return "GRCh38_no_alt_AnalysisSet"; }
  const std::vector<Sequence>& sequences() const override {
    static const std::vector<Sequence> table = {
#define NGSQ_SEQ(n, k) {n, k},
#include "grch38_no_alt_names.inc"
#undef NGSQ_SEQ
    };
    return table;
  }
};

inline bool eq_ignore_ascii_case(const std::string& a, const std::string& b) {
  return a.size() == b.size() && std::equal(a.begin(), a.end(), b.begin(), [](char x, char y) { return std::tolower((unsigned char)x) == std::tolower((unsigned char)y); });
}

// genome.rs:48-52
inline std::shared_ptr<ReferenceGenome> get_reference_genome(const std::string& s) {
  auto g = std::make_shared<GRCh38NoAltAnalysisSet>();
  if (eq_ignore_ascii_case(s, g->name())) return g;
  return nullptr;
}
// genome.rs:59-83: autosomes + sex + alt contigs (none here) + unlocalized + unplaced
inline std::vector<Sequence> get_primary_assembly(const ReferenceGenome& g) {
  std::vector<Sequence> out;
  for (auto& s : g.sequences()) if (s.kind == 'C' || s.kind == 'L' || s.kind == 'P') out.push_back(s);
  return out;
}
// genome.rs:86-126
inline std::vector<Sequence> get_all_sequences(const ReferenceGenome& g) { return g.sequences(); }

}  // namespace ngs
