// Histogram — the numeric container of every facet (reference: src/utils/histogram.rs:152-392).
// Same operations and the same floating-point operation order, so derived values are
// bit-identical to the reference's when the integer bins are.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace ngs {

struct BinOutOfBoundsError {};

class Histogram {
 public:
  Histogram() : Histogram(512) {}  // histogram.rs:394-398 (Default = 0..=512)
  static Histogram zero_based_with_capacity(uint64_t capacity) { return Histogram(capacity); }

  // histogram.rs:185-197
  bool increment(uint64_t bin) { return increment_by(bin, 1); }
  bool increment_by(uint64_t bin, uint64_t value) {
    if (bin < range_start_ || bin > range_stop_) return false;  // Err(BinOutOfBoundsError)
    values_[bin] += value;
    return true;
  }
  // histogram.rs:200-207 (panics out of range)
  uint64_t get(uint64_t bin) const {
    if (bin >= values_.size()) throw std::out_of_range("Could not lookup value for template length histogram bin: " + std::to_string(bin) + ".");
    return values_[bin];
  }
  const std::vector<uint64_t>& values() const { return values_; }
  std::vector<double> values_normalized() const {
    double total = (double)sum();
    std::vector<double> out;
    for (uint64_t v : values_) out.push_back((double)v / total);
    return out;
  }
  uint64_t range_len() const { return range_stop_ - range_start_ + 1; }
  uint64_t range_start() const { return range_start_; }
  uint64_t range_stop() const { return range_stop_; }
  bool in_range(uint64_t v) const { return v >= range_start_ && v <= range_stop_; }

  // histogram.rs:258-269
  double mean() const {
    double sum = 0.0, denominator = 0.0;
    for (uint64_t i = range_start_; i <= range_stop_; ++i) {
      uint64_t bin_value = get(i);
      denominator += (double)bin_value;
      sum += (double)(bin_value * i);
    }
    return sum / denominator;
  }
  // histogram.rs:272-337
  std::optional<double> percentile(double percentile) const {
    if (!(percentile >= 0.0 && percentile <= 1.0)) throw std::invalid_argument("Provided percentile was not within a valid range.");
    uint64_t num_items = 0;
    for (uint64_t i = range_start_; i <= range_stop_; ++i) num_items += get(i);
    if (num_items == 0) return std::nullopt;
    double needed_items = percentile * (double)num_items;
    double collected_items = 0.0;
    uint64_t index = range_start_;
    for (;;) {
      if (index > range_stop_) throw std::runtime_error("Unknown error!");
      collected_items += (double)get(index);
      if (collected_items > needed_items) return (double)index;
      if (collected_items == needed_items) {
        uint64_t lowest = index;
        index += 1;
        while (get(index) == 0) index += 1;
        uint64_t highest = index;
        return (double)lowest + ((double)(highest - lowest) / 2.0);
      }
      index += 1;
    }
  }
  std::optional<double> first_quartile() const { return percentile(0.25); }
  std::optional<double> median() const { return percentile(0.5); }
  std::optional<double> third_quartile() const { return percentile(0.75); }
  std::optional<double> interquartile_range() const {
    auto a = first_quartile(), b = third_quartile();
    if (a && b) return *b - *a;
    return std::nullopt;
  }
  uint64_t sum() const { uint64_t s = 0; for (uint64_t v : values_) s += v; return s; }
  uint64_t count_from_bottom_until(uint64_t bin) const { uint64_t s = 0; for (uint64_t i = range_start_; i <= bin; ++i) s += get(i); return s; }
  uint64_t count_from_top_until(uint64_t bin) const { uint64_t s = 0; for (uint64_t i = bin; i <= range_stop_; ++i) s += get(i); return s; }

  // bulk fill from engine counters (equivalent to increment_by per bin)
  void fill_from(const uint64_t* v, size_t n) { for (size_t i = 0; i < n && i < values_.size(); ++i) values_[i] += v[i]; }

 private:
  explicit Histogram(uint64_t capacity) : values_(capacity + 1, 0), range_start_(0), range_stop_(capacity) {}
  std::vector<uint64_t> values_;
  uint64_t range_start_, range_stop_;
};

}  // namespace ngs
