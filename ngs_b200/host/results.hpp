// Result structs and the results JSON (reference: src/qc/results.rs:24-60 and the metrics
// structs of every facet).  Field names and order are the reference's struct declaration order,
// so the file deserialises back into the reference's `Results` unchanged.
#pragma once
#include <cstdint>
#include <fstream>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "histogram.hpp"
#include "json.hpp"

namespace ngs {

// general/metrics.rs:10-136
struct ReadDesignationMetrics { uint64_t primary = 0, secondary = 0, supplementary = 0; };
struct RecordMetrics {
  uint64_t total = 0, unmapped = 0, duplicate = 0;
  ReadDesignationMetrics designation;
  uint64_t primary_mapped = 0, primary_duplicate = 0, paired = 0, read_1 = 0, read_2 = 0, proper_pair = 0, singleton = 0,
           mate_mapped = 0, mate_reference_sequence_id_mismatch = 0, mate_reference_sequence_id_mismatch_hq = 0;
};
struct CigarMetrics { std::map<std::string, uint64_t> read_one_cigar_ops, read_two_cigar_ops; };
struct GeneralSummaryMetrics { double duplication_pct, mapped_pct, mate_reference_sequence_id_mismatch_pct, mate_reference_sequence_id_mismatch_hq_pct; };
struct GeneralMetrics { RecordMetrics records; CigarMetrics cigar; std::optional<GeneralSummaryMetrics> summary; };

// gc_content/metrics.rs:10-68
struct NucleobaseMetrics { uint64_t total_gc_count = 0, total_at_count = 0, total_other_count = 0; };
struct GCRecordMetrics { uint64_t processed = 0, ignored_flags = 0, ignored_too_short = 0; };
struct GCSummaryMetrics { double gc_content_pct, ignored_flags_pct, ignored_too_short_pct; };
struct GCContentMetrics {
  Histogram histogram = Histogram::zero_based_with_capacity(100);  // gc_content.rs:129-138
  NucleobaseMetrics nucleobases; GCRecordMetrics records; std::optional<GCSummaryMetrics> summary;
};

// template_length.rs:14-53
struct TlenRecordMetrics { uint64_t processed = 0, ignored = 0; };
struct TlenSummaryMetrics { double template_length_unknown_pct, template_length_out_of_range_pct; };
struct TemplateLengthMetrics {
  Histogram histogram; TlenRecordMetrics records; std::optional<TlenSummaryMetrics> summary;
  explicit TemplateLengthMetrics(uint64_t capacity) : histogram(Histogram::zero_based_with_capacity(capacity)) {}
};

// quality_scores.rs:16-19 (HashMap<usize, Histogram>, keys are 1-based positions)
struct QualityScoreMetrics { std::map<uint64_t, Histogram> scores; };

// coverage.rs:28-69
struct IgnoredMetrics { uint64_t nonsensical_records = 0; std::map<std::string, uint64_t> pileup_too_large_positions; };
struct CoverageMetrics {
  std::map<std::string, double> mean_coverage;
  std::map<std::string, std::vector<double>> mean_coverage_per_bin;
  std::map<std::string, double> median_coverage, median_over_mean_coverage;
  IgnoredMetrics ignored;
  Histogram coverage_distribution = Histogram::zero_based_with_capacity(2048);  // coverage.rs:76
  std::map<std::string, float> genome_covered_by;
  std::vector<std::string> covered_by_order;  // "10x".."60x" in insertion order
};

// features/metrics.rs:8-80
struct FeaturesMetrics {
  uint64_t utr_five_prime_count = 0, utr_three_prime_count = 0, coding_sequence_count = 0;   // exonic_translation_regions
  uint64_t intergenic_count = 0, exonic_count = 0, intronic_count = 0;                       // gene_regions
  uint64_t processed = 0, ignored_flags = 0, ignored_nonprimary_chromosome = 0;              // records
  std::optional<std::pair<double, double>> summary;  // ignored_flags_pct, ignored_nonprimary_chromosome_pct
};

// edits.rs:22-62
struct EditMetrics {
  Histogram read_one_edits, read_two_edits;                                  // Histogram::default() = 0..=512
  Histogram vaf_histogram = Histogram::zero_based_with_capacity(100);
  std::optional<std::pair<double, double>> summary;  // mean_edits_read_one, mean_edits_read_two
};

// results.rs:24-45
struct Results {
  std::optional<GeneralMetrics> general;
  std::optional<FeaturesMetrics> features;  // filled only by the Genomic Features facet
  std::optional<GCContentMetrics> gc_content;
  std::optional<TemplateLengthMetrics> template_length;
  std::optional<QualityScoreMetrics> quality_scores;
  std::optional<CoverageMetrics> coverage;
  std::optional<EditMetrics> edits;         // filled only by the Edits facet

  static void write_hist(JsonWriter& w, const Histogram& h) {  // histogram.rs:152-159 field order
    w.begin_object();
    w.key("values"); w.begin_array(); for (uint64_t v : h.values()) w.value_u64(v); w.end_array();
    w.key("range_start"); w.value_u64(h.range_start());
    w.key("range_stop"); w.value_u64(h.range_stop());
    w.end_object();
  }

  std::string to_json_pretty() const {
    JsonWriter w;
    w.begin_object();
    w.key("general");
    if (!general) w.value_null();
    else {
      const auto& r = general->records;
      w.begin_object();
      w.key("records"); w.begin_object();
      w.key("total"); w.value_u64(r.total); w.key("unmapped"); w.value_u64(r.unmapped); w.key("duplicate"); w.value_u64(r.duplicate);
      w.key("designation"); w.begin_object();
      w.key("primary"); w.value_u64(r.designation.primary); w.key("secondary"); w.value_u64(r.designation.secondary);
      w.key("supplementary"); w.value_u64(r.designation.supplementary); w.end_object();
      w.key("primary_mapped"); w.value_u64(r.primary_mapped); w.key("primary_duplicate"); w.value_u64(r.primary_duplicate);
      w.key("paired"); w.value_u64(r.paired); w.key("read_1"); w.value_u64(r.read_1); w.key("read_2"); w.value_u64(r.read_2);
      w.key("proper_pair"); w.value_u64(r.proper_pair); w.key("singleton"); w.value_u64(r.singleton); w.key("mate_mapped"); w.value_u64(r.mate_mapped);
      w.key("mate_reference_sequence_id_mismatch"); w.value_u64(r.mate_reference_sequence_id_mismatch);
      w.key("mate_reference_sequence_id_mismatch_hq"); w.value_u64(r.mate_reference_sequence_id_mismatch_hq);
      w.end_object();
      w.key("cigar"); w.begin_object();
      w.key("read_one_cigar_ops"); w.begin_object(); for (auto& kv : general->cigar.read_one_cigar_ops) { w.key(kv.first); w.value_u64(kv.second); } w.end_object();
      w.key("read_two_cigar_ops"); w.begin_object(); for (auto& kv : general->cigar.read_two_cigar_ops) { w.key(kv.first); w.value_u64(kv.second); } w.end_object();
      w.end_object();
      w.key("summary");
      if (!general->summary) w.value_null();
      else {
        w.begin_object();
        w.key("duplication_pct"); w.value_f64(general->summary->duplication_pct);
        w.key("mapped_pct"); w.value_f64(general->summary->mapped_pct);
        w.key("mate_reference_sequence_id_mismatch_pct"); w.value_f64(general->summary->mate_reference_sequence_id_mismatch_pct);
        w.key("mate_reference_sequence_id_mismatch_hq_pct"); w.value_f64(general->summary->mate_reference_sequence_id_mismatch_hq_pct);
        w.end_object();
      }
      w.end_object();
    }
    w.key("features");
    if (!features) w.value_null();
    else {
      const auto& f = *features;
      w.begin_object();
      w.key("exonic_translation_regions"); w.begin_object();
      w.key("utr_five_prime_count"); w.value_u64(f.utr_five_prime_count); w.key("utr_three_prime_count"); w.value_u64(f.utr_three_prime_count);
      w.key("coding_sequence_count"); w.value_u64(f.coding_sequence_count); w.end_object();
      w.key("gene_regions"); w.begin_object();
      w.key("intergenic_count"); w.value_u64(f.intergenic_count); w.key("exonic_count"); w.value_u64(f.exonic_count);
      w.key("intronic_count"); w.value_u64(f.intronic_count); w.end_object();
      w.key("records"); w.begin_object();
      w.key("processed"); w.value_u64(f.processed); w.key("ignored_flags"); w.value_u64(f.ignored_flags);
      w.key("ignored_nonprimary_chromosome"); w.value_u64(f.ignored_nonprimary_chromosome); w.end_object();
      w.key("summary");
      if (!f.summary) w.value_null();
      else {
        w.begin_object();
        w.key("ignored_flags_pct"); w.value_f64(f.summary->first);
        w.key("ignored_nonprimary_chromosome_pct"); w.value_f64(f.summary->second);
        w.end_object();
      }
      w.end_object();
    }
    w.key("gc_content");
    if (!gc_content) w.value_null();
    else {
      w.begin_object();
      w.key("histogram"); write_hist(w, gc_content->histogram);
      w.key("nucleobases"); w.begin_object();
      w.key("total_gc_count"); w.value_u64(gc_content->nucleobases.total_gc_count);
      w.key("total_at_count"); w.value_u64(gc_content->nucleobases.total_at_count);
      w.key("total_other_count"); w.value_u64(gc_content->nucleobases.total_other_count); w.end_object();
      w.key("records"); w.begin_object();
      w.key("processed"); w.value_u64(gc_content->records.processed);
      w.key("ignored_flags"); w.value_u64(gc_content->records.ignored_flags);
      w.key("ignored_too_short"); w.value_u64(gc_content->records.ignored_too_short); w.end_object();
      w.key("summary");
      if (!gc_content->summary) w.value_null();
      else {
        w.begin_object();
        w.key("gc_content_pct"); w.value_f64(gc_content->summary->gc_content_pct);
        w.key("ignored_flags_pct"); w.value_f64(gc_content->summary->ignored_flags_pct);
        w.key("ignored_too_short_pct"); w.value_f64(gc_content->summary->ignored_too_short_pct);
        w.end_object();
      }
      w.end_object();
    }
    w.key("template_length");
    if (!template_length) w.value_null();
    else {
      w.begin_object();
      w.key("histogram"); write_hist(w, template_length->histogram);
      w.key("records"); w.begin_object();
      w.key("processed"); w.value_u64(template_length->records.processed);
      w.key("ignored"); w.value_u64(template_length->records.ignored); w.end_object();
      w.key("summary");
      if (!template_length->summary) w.value_null();
      else {
        w.begin_object();
        w.key("template_length_unknown_pct"); w.value_f64(template_length->summary->template_length_unknown_pct);
        w.key("template_length_out_of_range_pct"); w.value_f64(template_length->summary->template_length_out_of_range_pct);
        w.end_object();
      }
      w.end_object();
    }
    w.key("quality_scores");
    if (!quality_scores) w.value_null();
    else {
      w.begin_object();
      w.key("scores"); w.begin_object();
      for (auto& kv : quality_scores->scores) { w.key(std::to_string(kv.first)); write_hist(w, kv.second); }
      w.end_object();
      w.end_object();
    }
    w.key("coverage");
    if (!coverage) w.value_null();
    else {
      const auto& c = *coverage;
      w.begin_object();
      w.key("mean_coverage"); w.begin_object(); for (auto& kv : c.mean_coverage) { w.key(kv.first); w.value_f64(kv.second); } w.end_object();
      w.key("mean_coverage_per_bin"); w.begin_object();
      for (auto& kv : c.mean_coverage_per_bin) { w.key(kv.first); w.begin_array(); for (double v : kv.second) w.value_f64(v); w.end_array(); }
      w.end_object();
      w.key("median_coverage"); w.begin_object(); for (auto& kv : c.median_coverage) { w.key(kv.first); w.value_f64(kv.second); } w.end_object();
      w.key("median_over_mean_coverage"); w.begin_object(); for (auto& kv : c.median_over_mean_coverage) { w.key(kv.first); w.value_f64(kv.second); } w.end_object();
      w.key("ignored"); w.begin_object();
      w.key("nonsensical_records"); w.value_u64(c.ignored.nonsensical_records);
      w.key("pileup_too_large_positions"); w.begin_object(); for (auto& kv : c.ignored.pileup_too_large_positions) { w.key(kv.first); w.value_u64(kv.second); } w.end_object();
      w.end_object();
      w.key("coverage_distribution"); write_hist(w, c.coverage_distribution);
      w.key("genome_covered_by"); w.begin_object(); for (auto& k : c.covered_by_order) { w.key(k); w.value_f32(c.genome_covered_by.at(k)); } w.end_object();
      w.end_object();
    }
    w.key("edits");
    if (!edits) w.value_null();
    else {
      w.begin_object();
      w.key("read_one_edits"); write_hist(w, edits->read_one_edits);
      w.key("read_two_edits"); write_hist(w, edits->read_two_edits);
      w.key("vaf_histogram"); write_hist(w, edits->vaf_histogram);
      w.key("summary");
      if (!edits->summary) w.value_null();
      else {
        w.begin_object();
        w.key("mean_edits_read_one"); w.value_f64(edits->summary->first);
        w.key("mean_edits_read_two"); w.value_f64(edits->summary->second);
        w.end_object();
      }
      w.end_object();
    }
    w.end_object();
    return w.out;
  }

  // results.rs:50-60: <directory>/<prefix>.results.json, silently overwritten
  void write(const std::string& output_prefix, const std::string& directory) const {
    std::string path = directory + "/" + output_prefix + ".results.json";
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot create " + path);
    f << to_json_pretty();
  }
};

}  // namespace ngs
