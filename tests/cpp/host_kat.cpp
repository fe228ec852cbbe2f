// Known-answer vectors of the reference's own unit tests, run against the PRODUCT's host-side C++
// (ngs_b200/host/histogram.hpp, facets.hpp): src/utils/histogram.rs:405-523, src/qc.rs:238-271,
// src/qc/record_based/gc_content.rs:145-151.  Prints one "name value" line per check; the Python test
// (tests/test_host_cpp.py) compares them with the expected values of SURVEY App. E.
#include <cstdio>
#include <optional>
#include <string>

#include "../../ngs_b200/host/facets.hpp"
#include "../../ngs_b200/host/genome.hpp"
#include "../../ngs_b200/host/histogram.hpp"

using namespace ngs;

static void show(const char* name, std::optional<double> v) {
  if (v) printf("%s %.17g\n", name, *v); else printf("%s None\n", name);
}

int main() {
  {  // histogram.rs:414-431
    Histogram h = Histogram::zero_based_with_capacity(100);
    h.increment(25); h.increment(50); h.increment_by(75, 3); h.increment_by(100, 5);
    show("mean", h.mean()); show("q1", h.first_quartile()); show("median", h.median()); show("q3", h.third_quartile());
    show("iqr", h.interquartile_range());
  }
  {  // histogram.rs:434-437
    Histogram h = Histogram::zero_based_with_capacity(5000);
    show("empty_median", h.median());
  }
  {  // histogram.rs:440-463
    Histogram h = Histogram::zero_based_with_capacity(5000);
    h.increment_by(0, 2500); h.increment_by(10, 2500); h.increment_by(100, 2500); h.increment_by(5000, 5000);
    show("tie_median_1", h.median());
    h.increment_by(200, 2500);
    show("tie_median_2", h.median());
    h.increment(200);
    show("tie_median_3", h.median());
  }
  {  // histogram.rs:466-469
    Histogram h = Histogram::zero_based_with_capacity(100);
    printf("out_of_bounds %d\n", h.increment(101) ? 0 : 1);
  }
  {  // histogram.rs:472-481
    Histogram h;
    printf("default_range %llu %llu %zu\n", (unsigned long long)h.range_start(), (unsigned long long)h.range_stop(), h.values().size());
  }
  {  // histogram.rs:484-523
    Histogram h = Histogram::zero_based_with_capacity(3);
    h.increment(1); h.increment(2); h.increment_by(3, 3);
    printf("values %llu %llu %llu %llu\n", (unsigned long long)h.get(0), (unsigned long long)h.get(1), (unsigned long long)h.get(2), (unsigned long long)h.get(3));
    auto n = h.values_normalized();
    printf("normalized %.17g %.17g %.17g %.17g\n", n[0], n[1], n[2], n[3]);
    Histogram c = Histogram::zero_based_with_capacity(3);
    c.increment_by(0, 5); c.increment_by(1, 3); c.increment_by(2, 6);
    printf("bottom %llu %llu %llu %llu\n", (unsigned long long)c.count_from_bottom_until(0), (unsigned long long)c.count_from_bottom_until(1),
           (unsigned long long)c.count_from_bottom_until(2), (unsigned long long)c.count_from_bottom_until(3));
    printf("top %llu %llu %llu %llu\n", (unsigned long long)c.count_from_top_until(3), (unsigned long long)c.count_from_top_until(2),
           (unsigned long long)c.count_from_top_until(1), (unsigned long long)c.count_from_top_until(0));
  }
  {  // qc.rs:238-271
    auto genome = get_reference_genome("GRCh38_no_alt_AnalysisSet");
    FacetSet all = get_qc_facets(*genome, std::nullopt);
    printf("default_facets %zu %zu\n", all.record_based.size(), all.sequence_based.size());
    FacetSet only = get_qc_facets(*genome, std::string("GC Content"));
    printf("only_gc %zu %zu %s\n", only.record_based.size(), only.sequence_based.size(), only.record_based[0]->name());
    FacetSet lower = get_qc_facets(*genome, std::string("coverage"));
    printf("only_coverage_case_insensitive %zu %zu\n", lower.record_based.size(), lower.sequence_based.size());
    try { get_qc_facets(*genome, std::string("Nope")); printf("only_unknown accepted\n"); }
    catch (const std::exception& e) { printf("only_unknown rejected\n"); }
  }
  {  // gc_content.rs:145-151
    GCContentFacet f;
    printf("gc_hist_range %llu %llu %zu\n", (unsigned long long)f.metrics.histogram.range_start(), (unsigned long long)f.metrics.histogram.range_stop(),
           f.metrics.histogram.values().size());
  }
  return 0;
}
