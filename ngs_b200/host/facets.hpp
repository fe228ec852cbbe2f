// Facet plugin surface of `ngs qc` (reference: src/qc.rs:151-229) with CUDA-backed facets.
//
// Names, lifecycle methods and the registry mirror the reference one to one.  The per-record
// `process(&Record)` granularity is exactly what presupposes CPU inflate + decode, so on the
// CUDA path `process` is a no-op: the engine runs both hot loops of app() on the GPU and each
// facet pulls its exact integer state through the C ABI (`ingest`).  summarize()/teardown()/
// aggregate() then run the reference's own float arithmetic on the host, in the same order.
#pragma once
#include <charconv>
#include <fstream>
#include <optional>
#include <filesystem>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ngs_cuda.h"
#include "gene_model.hpp"
#include "genome.hpp"
#include "results.hpp"

namespace ngs {

// src/qc.rs:133-143
enum class ComputationalLoad { Light, Moderate, Heavy };
inline const char* to_string(ComputationalLoad l) { return l == ComputationalLoad::Light ? "Light" : l == ComputationalLoad::Moderate ? "Moderate" : "Heavy"; }

struct Record;  // the CPU record type does not exist on this path (decoding happens on the device)
struct ReferenceSequence { std::string name; uint32_t length; };

inline void check(ngsq_engine* e, int rc) { if (rc) throw std::runtime_error(ngsq_last_error(e)); }

// src/qc.rs:151-176
class RecordBasedQualityControlFacet {
 public:
  virtual ~RecordBasedQualityControlFacet() = default;
  virtual const char* name() const = 0;
  virtual ComputationalLoad computational_load() const = 0;
  virtual void process(const Record&) {}          // no-op: the engine processed every record on the GPU
  virtual void ingest(ngsq_engine* e) = 0;        // exact integers from the device
  virtual void summarize() = 0;
  virtual void aggregate(Results& results) const = 0;
};

// src/qc.rs:184-229
class SequenceBasedQualityControlFacet {
 public:
  virtual ~SequenceBasedQualityControlFacet() = default;
  virtual const char* name() const = 0;
  virtual ComputationalLoad computational_load() const = 0;
  virtual bool supports_sequence_name(const std::string& name) const = 0;
  virtual void setup(const ReferenceSequence&) {}
  virtual void process(const ReferenceSequence&, const Record&) {}  // no-op, see above
  virtual void ingest(ngsq_engine* e, uint32_t ref_id, const ReferenceSequence& seq) = 0;
  virtual void teardown(const ReferenceSequence& seq) = 0;
  virtual void aggregate(Results& results) = 0;
  virtual void ingest_global(ngsq_engine*) {}
};

// ---- General (src/qc/record_based/general.rs) ----
class GeneralMetricsFacet : public RecordBasedQualityControlFacet {
 public:
  GeneralMetrics metrics;
  const char* name() const override { return "General"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Light; }
  void ingest(ngsq_engine* e) override {
    uint64_t g[34];
    check(e, ngsq_get_general(e, g));
    auto& r = metrics.records;
    r.total = g[0]; r.unmapped = g[1]; r.duplicate = g[2];
    r.designation.primary = g[3]; r.designation.secondary = g[4]; r.designation.supplementary = g[5];
    r.primary_mapped = g[6]; r.primary_duplicate = g[7]; r.paired = g[8]; r.read_1 = g[9]; r.read_2 = g[10];
    r.proper_pair = g[11]; r.singleton = g[12]; r.mate_mapped = g[13];
    r.mate_reference_sequence_id_mismatch = g[14]; r.mate_reference_sequence_id_mismatch_hq = g[15];
    static const char* kinds[9] = {"M", "I", "D", "N", "S", "H", "P", "=", "X"};  // op.kind().to_string()
    for (int k = 0; k < 9; ++k) {
      if (g[16 + k]) metrics.cigar.read_one_cigar_ops[kinds[k]] = g[16 + k];  // entries exist only for kinds seen
      if (g[25 + k]) metrics.cigar.read_two_cigar_ops[kinds[k]] = g[25 + k];
    }
  }
  void summarize() override {  // general.rs:126-153
    const auto& r = metrics.records;
    GeneralSummaryMetrics s;
    s.duplication_pct = (double)r.duplicate / (double)r.total * 100.0;
    s.mapped_pct = (1.0 - (double)r.unmapped / (double)r.total) * 100.0;
    s.mate_reference_sequence_id_mismatch_pct = (double)r.mate_reference_sequence_id_mismatch / (double)r.total * 100.0;
    s.mate_reference_sequence_id_mismatch_hq_pct = (double)r.mate_reference_sequence_id_mismatch_hq / (double)r.total * 100.0;
    metrics.summary = s;
  }
  void aggregate(Results& results) const override { results.general = metrics; }
};

// ---- Template Length (src/qc/record_based/template_length.rs) ----
class TemplateLengthFacet : public RecordBasedQualityControlFacet {
 public:
  TemplateLengthMetrics m;
  static TemplateLengthFacet with_capacity(uint64_t capacity) { return TemplateLengthFacet(capacity); }
  const char* name() const override { return "Template Length"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Light; }
  void ingest(ngsq_engine* e) override {
    if (m.histogram.range_stop() != 1024) throw std::runtime_error("the CUDA engine bins template lengths 0..=1024 (src/qc.rs:62)");
    uint64_t h[1025], p = 0, i = 0;
    check(e, ngsq_get_tlen(e, h, &p, &i));
    m.histogram.fill_from(h, 1025);
    m.records.processed = p;
    m.records.ignored = i;
  }
  void summarize() override {  // template_length.rs:89-100
    TlenSummaryMetrics s;
    s.template_length_unknown_pct = ((double)m.histogram.get(0) / ((double)m.records.processed + (double)m.records.ignored)) * 100.0;
    s.template_length_out_of_range_pct = ((double)m.records.ignored / ((double)m.records.processed + (double)m.records.ignored)) * 100.0;
    m.summary = s;
  }
  void aggregate(Results& results) const override { results.template_length = m; }

 private:
  explicit TemplateLengthFacet(uint64_t capacity) : m(capacity) {}
};

// ---- GC Content (src/qc/record_based/gc_content.rs) ----
class GCContentFacet : public RecordBasedQualityControlFacet {
 public:
  GCContentMetrics metrics;
  const char* name() const override { return "GC Content"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Light; }
  void ingest(ngsq_engine* e) override {
    uint64_t h[101], nuc[3], rec[3];
    check(e, ngsq_get_gc(e, h, nuc, rec));
    metrics.histogram.fill_from(h, 101);
    metrics.nucleobases.total_gc_count = nuc[0]; metrics.nucleobases.total_at_count = nuc[1]; metrics.nucleobases.total_other_count = nuc[2];
    metrics.records.processed = rec[0]; metrics.records.ignored_flags = rec[1]; metrics.records.ignored_too_short = rec[2];
  }
  void summarize() override {  // gc_content.rs:102-122 (integer sums first, then `as f64`)
    const auto& n = metrics.nucleobases;
    const auto& r = metrics.records;
    GCSummaryMetrics s;
    s.gc_content_pct = ((double)n.total_gc_count / (double)(n.total_gc_count + n.total_at_count + n.total_other_count)) * 100.0;
    s.ignored_flags_pct = ((double)r.ignored_flags / (double)(r.ignored_flags + r.ignored_too_short + r.processed)) * 100.0;
    s.ignored_too_short_pct = ((double)r.ignored_too_short / (double)(r.ignored_flags + r.ignored_too_short + r.processed)) * 100.0;
    metrics.summary = s;
  }
  void aggregate(Results& results) const override { results.gc_content = metrics; }
};

// ---- Quality Score (src/qc/record_based/quality_scores.rs) ----
constexpr uint64_t MAX_SCORE = 93;  // quality_scores.rs:26
class QualityScoreFacet : public RecordBasedQualityControlFacet {
 public:
  QualityScoreMetrics m;
  const char* name() const override { return "Quality Score"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Moderate; }
  void ingest(ngsq_engine* e) override {
    uint32_t n = 0;
    check(e, ngsq_get_quality(e, nullptr, 0, &n));
    std::vector<uint64_t> q((size_t)n * 94 + 1);
    check(e, ngsq_get_quality(e, q.data(), n, &n));
    // positions 1..=n exist: a record of length n with qualities creates every entry up to n (quality_scores.rs:38-46)
    for (uint32_t p = 0; p < n; ++p) {
      Histogram h = Histogram::zero_based_with_capacity(MAX_SCORE);
      h.fill_from(&q[(size_t)p * 94], 94);
      m.scores.emplace((uint64_t)p + 1, std::move(h));
    }
  }
  void summarize() override {}
  void aggregate(Results& results) const override { results.quality_scores = m; }
};

// ---- Coverage (src/qc/sequence_based/coverage.rs) ----
class CoverageFacet : public SequenceBasedQualityControlFacet {
 public:
  CoverageFacet(const ReferenceGenome& genome, uint64_t bin_size) : primary_assembly_(get_primary_assembly(genome)), bin_size_(bin_size) {
    if (bin_size != 50000) throw std::runtime_error("the CUDA engine resolves coverage in 50,000-bp bins (src/qc.rs:86-89)");
  }
  const char* name() const override { return "Coverage"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Moderate; }
  bool supports_sequence_name(const std::string& name) const override {  // coverage.rs:133-138
    for (auto& s : primary_assembly_) if (s.name == name) return true;
    return false;
  }
  void ingest(ngsq_engine* e, uint32_t ref_id, const ReferenceSequence& seq) override {
    Pending p;
    p.bin_sums.resize((size_t)seq.length / 50000 + 3);
    check(e, ngsq_get_coverage_contig(e, ref_id, &p.ints, p.bin_sums.data(), p.bin_sums.size()));
    p.bin_sums.resize(p.ints.n_bins);
    pending_[seq.name] = std::move(p);
  }
  void ingest_global(ngsq_engine* e) override { check(e, ngsq_get_coverage_global(e, &metrics_.ignored.nonsensical_records)); }
  // coverage.rs:182-262, fed with the exact depth histogram / bin sums instead of a per-position sweep
  void teardown(const ReferenceSequence& seq) override {
    auto it = pending_.find(seq.name);
    if (it == pending_.end() || !it->second.ints.touched) return;  // no record was returned for this sequence
    const Pending& p = it->second;
    Histogram coverages = Histogram::zero_based_with_capacity(2048);
    coverages.fill_from(p.ints.hist, 2049);
    uint64_t ignored = p.ints.pileup_too_large;
    std::vector<double>& bins = metrics_.mean_coverage_per_bin[seq.name];
    uint64_t full_bins = 1 + seq.length / bin_size_;  // i = 0, 50000, 100000, ... <= L
    for (uint64_t k = 0; k < full_bins; ++k) bins.push_back((double)p.bin_sums[k] / (double)bin_size_);
    uint64_t modulo = seq.length % bin_size_;
    if (modulo != 0) bins.push_back((double)p.bin_sums[full_bins] / (double)modulo);
    double mean = coverages.mean();
    double median = coverages.median().value();
    double median_over_mean = median / mean;
    for (uint64_t i = 0; i <= 2048; ++i) metrics_.coverage_distribution.increment_by(i, coverages.get(i));
    metrics_.mean_coverage[seq.name] = mean;
    metrics_.median_coverage[seq.name] = median;
    metrics_.median_over_mean_coverage[seq.name] = median_over_mean;
    metrics_.ignored.pileup_too_large_positions[seq.name] = ignored;
  }
  void aggregate(Results& results) override {  // coverage.rs:264-287 (f32 arithmetic)
    uint64_t total_positions = metrics_.coverage_distribution.sum();
    for (auto& kv : metrics_.ignored.pileup_too_large_positions) total_positions += kv.second;
    static const uint64_t COVERAGES_TO_CHECK[6] = {10, 20, 30, 40, 50, 60};
    for (uint64_t c : COVERAGES_TO_CHECK) {
      std::string k = std::to_string(c) + "x";
      uint64_t at_least = metrics_.coverage_distribution.count_from_top_until(c);
      float v = ((float)at_least / (float)total_positions) * 100.0f;
      metrics_.genome_covered_by[k] = v;
      metrics_.covered_by_order.push_back(k);
    }
    results.coverage = metrics_;
  }

 private:
  struct Pending { ngsq_cov_ints ints{}; std::vector<uint64_t> bin_sums; };
  std::map<std::string, Pending> pending_;
  CoverageMetrics metrics_;
  std::vector<Sequence> primary_assembly_;
  uint64_t bin_size_;
};

// ---- Genomic Features (src/qc/record_based/features.rs): process() runs on the device (features.cuh) ----
class GenomicFeaturesFacet : public RecordBasedQualityControlFacet {
 public:
  FeaturesMetrics metrics;
  // features.rs:270-355: reads and files the gene model; throws where the reference errors or panics
  static GenomicFeaturesFacet try_from(const std::string& src, const FeatureNames& names, const ReferenceGenome& genome) {
    GenomicFeaturesFacet f;
    f.names_ = names;
    f.model_ = read_gene_model(slurp_maybe_gz(src, "GFF"), names, genome);
    for (auto& s : get_primary_assembly(genome)) f.primary_.push_back(s.name);
    return f;
  }
  const char* name() const override { return "Genomic Features"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Moderate; }
  // hands the model to an engine (after ngsq_set_references, before the first submit)
  void upload(ngsq_engine* e, const std::vector<ReferenceSequence>& header) const {
    std::vector<uint8_t> primary(header.size(), 0);
    for (size_t c = 0; c < header.size(); ++c)
      for (auto& p : primary_) if (p == header[c].name) primary[c] = 1;
    const auto sc = names_.slot_class();
    check(e, ngsq_set_feature_model(e, sc.data(), primary.data()));
    for (size_t c = 0; c < header.size(); ++c) {
      auto it = model_.find(header[c].name);
      if (it == model_.end() || !primary[c]) continue;
      const ContigFeatures& m = it->second;
      check(e, ngsq_set_features(e, (uint32_t)c, (uint32_t)m.start.size(), m.start.data(), m.stop.data(), m.cls.data()));
    }
  }
  void ingest(ngsq_engine* e) override {
    uint64_t c[9];
    check(e, ngsq_get_features(e, c));
    metrics.utr_five_prime_count = c[0]; metrics.utr_three_prime_count = c[1]; metrics.coding_sequence_count = c[2];
    metrics.intergenic_count = c[3]; metrics.exonic_count = c[4]; metrics.intronic_count = c[5];
    metrics.processed = c[6]; metrics.ignored_flags = c[7]; metrics.ignored_nonprimary_chromosome = c[8];
  }
  void summarize() override {  // features.rs:244-262 (integer sum first, then `as f64`)
    const double total = (double)(metrics.ignored_flags + metrics.ignored_nonprimary_chromosome + metrics.processed);
    metrics.summary = std::make_pair(((double)metrics.ignored_flags / total) * 100.0, ((double)metrics.ignored_nonprimary_chromosome / total) * 100.0);
  }
  void aggregate(Results& results) const override { results.features = metrics; }
  const std::map<std::string, ContigFeatures>& model() const { return model_; }

 private:
  FeatureNames names_;
  std::map<std::string, ContigFeatures> model_;
  std::vector<std::string> primary_;
};

// ---- Edits (src/qc/sequence_based/edits.rs): process() and the VAF histogram of teardown() run on the device ----
class EditsFacet : public SequenceBasedQualityControlFacet {
 public:
  EditMetrics metrics;
  // edits.rs:117-161
  static EditsFacet try_from(const std::string& reference_fasta, const std::optional<std::string>& vaf_file_path = std::nullopt) {
    EditsFacet f;
    f.sequences_ = read_fasta(slurp_maybe_gz(reference_fasta, "reference FASTA"));
    if (vaf_file_path) {
      if (std::filesystem::exists(*vaf_file_path))
        throw std::runtime_error("refusing to overwrite existing VAF file: " + *vaf_file_path + ". Please delete and rerun if you'd like to replace it.");
      f.vaf_file_ = std::make_unique<std::ofstream>(*vaf_file_path, std::ios::binary);
      if (!*f.vaf_file_) throw std::runtime_error("creating VAF file");
      *f.vaf_file_ << "Sequence\tPosition\tVAF\n";
      if (!*f.vaf_file_) throw std::runtime_error("writing VAF file header");
    }
    return f;
  }
  const char* name() const override { return "Edits"; }
  ComputationalLoad computational_load() const override { return ComputationalLoad::Heavy; }
  bool supports_sequence_name(const std::string&) const override { return true; }  // edits.rs:178-180
  void upload(ngsq_engine* e, const std::vector<ReferenceSequence>& header) const {
    for (size_t c = 0; c < header.size(); ++c) {
      auto it = sequences_.find(header[c].name);
      if (it == sequences_.end()) throw std::runtime_error("sequence " + header[c].name + " not found in reference FASTA.");  // edits.rs:205-207
      check(e, ngsq_set_reference_bases(e, (uint32_t)c, reinterpret_cast<const uint8_t*>(it->second.data()), it->second.size()));
    }
  }
  // with several devices every engine holds the per-position counts of its own (contig-exclusive) shard
  void set_engines(const std::vector<ngsq_engine*>& engines) { engines_ = engines; }
  bool writes_vaf_file() const { return (bool)vaf_file_; }
  void ingest_global(ngsq_engine* e) override {
    std::vector<uint64_t> one(513), two(513), vaf(101);
    uint64_t records = 0;
    check(e, ngsq_get_edits(e, one.data(), two.data(), vaf.data(), &records));
    metrics.read_one_edits.fill_from(one.data(), 513);
    metrics.read_two_edits.fill_from(two.data(), 513);
    metrics.vaf_histogram.fill_from(vaf.data(), 101);
  }
  // teardown of one sequence (edits.rs:305-340).  The VAF histogram came from the device (edits_vaf_kernel); the VAF file,
  // when asked for, is written here from the sequence's per-position counters: one line per position a record's `M` covered.
  void ingest(ngsq_engine* root, uint32_t ref, const ReferenceSequence& seq) override {
    if (!vaf_file_) return;
    const uint64_t n = (uint64_t)seq.length + 1;
    std::vector<uint32_t> refs(n, 0), alts(n, 0), r(n), a(n);
    std::vector<ngsq_engine*> engines = engines_.empty() ? std::vector<ngsq_engine*>{root} : engines_;
    for (ngsq_engine* e : engines) {
      check(e, ngsq_get_edit_positions(e, ref, r.data(), a.data(), n));
      for (uint64_t i = 0; i < n; ++i) { refs[i] += r[i]; alts[i] += a[i]; }
    }
    std::string out;
    char buf[64];
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t total = (uint64_t)refs[i] + alts[i];
      if (!total) continue;
      out += seq.name;
      out += '\t';
      out += std::to_string(i);
      out += '\t';
      out.append(buf, format_f32_display(buf, sizeof buf, (float)alts[i] / (float)total));
      out += '\n';
      if (out.size() > (1u << 20)) { vaf_file_->write(out.data(), (std::streamsize)out.size()); out.clear(); }
    }
    vaf_file_->write(out.data(), (std::streamsize)out.size());
    if (!*vaf_file_) throw std::runtime_error("writing VAF file");
  }
  void teardown(const ReferenceSequence&) override {}
  void aggregate(Results& results) override {          // edits.rs:336-344
    metrics.summary = std::make_pair(metrics.read_one_edits.mean(), metrics.read_two_edits.mean());
    if (vaf_file_) vaf_file_->flush();
    results.edits = metrics;
  }
  // `{}` of an f32 in Rust: the shortest decimal that reads back as the same f32, positional, no exponent, no trailing ".0"
  // (edits.rs:338).  std::to_chars in fixed notation without a precision is that representation.
  static size_t format_f32_display(char* buf, size_t cap, float v) {
    auto res = std::to_chars(buf, buf + cap, v, std::chars_format::fixed);
    return (size_t)(res.ptr - buf);
  }

 private:
  std::map<std::string, std::string> sequences_;
  std::unique_ptr<std::ofstream> vaf_file_;
  std::vector<ngsq_engine*> engines_;
};

// src/qc.rs:44-126 — default facet set and the `--only` filter.  Genomic Features (needs a GFF) and Edits (needs a
// reference FASTA) join it when their inputs were given, as in get_qc_facets.
struct FacetSet {
  std::vector<std::unique_ptr<RecordBasedQualityControlFacet>> record_based;
  std::vector<std::unique_ptr<SequenceBasedQualityControlFacet>> sequence_based;
};
inline FacetSet get_qc_facets(const ReferenceGenome& genome, const std::optional<std::string>& only_facet,
                              std::unique_ptr<GenomicFeaturesFacet> features = nullptr, std::unique_ptr<EditsFacet> edits = nullptr) {
  FacetSet fs;
  fs.record_based.push_back(std::make_unique<GeneralMetricsFacet>());
  fs.record_based.push_back(std::make_unique<TemplateLengthFacet>(TemplateLengthFacet::with_capacity(1024)));
  fs.record_based.push_back(std::make_unique<GCContentFacet>());
  fs.record_based.push_back(std::make_unique<QualityScoreFacet>());
  if (features) fs.record_based.push_back(std::move(features));  // qc.rs:68-79: after the four defaults
  fs.sequence_based.push_back(std::make_unique<CoverageFacet>(genome, 50000));
  if (edits) fs.sequence_based.push_back(std::move(edits));      // qc.rs:92-94
  if (only_facet) {
    FacetSet filtered;
    for (auto& f : fs.record_based) if (eq_ignore_ascii_case(f->name(), *only_facet)) filtered.record_based.push_back(std::move(f));
    for (auto& f : fs.sequence_based) if (eq_ignore_ascii_case(f->name(), *only_facet)) filtered.sequence_based.push_back(std::move(f));
    size_t n = filtered.record_based.size() + filtered.sequence_based.size();
    if (n == 0) throw std::runtime_error("No facets matched the specified `--only` flag: " + *only_facet);
    if (n > 1) throw std::runtime_error("Too many facets matched the specified `--only` flag: " + *only_facet);
    return filtered;
  }
  return fs;
}

}  // namespace ngs
