// Pretty JSON writer reproducing serde_json::to_string_pretty (reference: src/qc/results.rs:50-60):
// 2-space indent, one array element per line, `null` for None / non-finite floats, floats in
// ryu's shortest round-trip notation.  Map-typed fields are emitted with sorted
// keys because the reference's HashMap order is not reproducible.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ngs {

// digits/exponent of the shortest decimal that round-trips, then ryu's "pretty" layout.
template <class F>
inline std::string format_float(F v, int max_prec, int kk_hi, int kk_lo) {
  if (std::isnan(v) || std::isinf(v)) return "null";
  if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= max_prec; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v);
    F back = sizeof(F) == 4 ? (F)strtof(buf, nullptr) : (F)strtod(buf, nullptr);
    if (back == v) break;
  }
  // buf = [-]d.ddddde[+-]XX
  std::string s(buf);
  bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  size_t epos = s.find('e');
  int exp10 = atoi(s.c_str() + epos + 1);
  std::string digits;
  for (size_t i = 0; i < epos; ++i) if (s[i] != '.') digits.push_back(s[i]);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  int length = (int)digits.size();
  int k = exp10 - (length - 1);  // value = digits * 10^k
  int kk = length + k;
  std::string out = neg ? "-" : "";
  if (0 <= k && kk <= kk_hi) {
    out += digits + std::string(k, '0') + ".0";
  } else if (0 < kk && kk <= kk_hi) {
    out += digits.substr(0, kk) + "." + digits.substr(kk);
  } else if (kk_lo < kk && kk <= 0) {
    out += "0." + std::string(-kk, '0') + digits;
  } else if (length == 1) {
    out += digits + "e" + std::to_string(kk - 1);
  } else {
    out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
  }
  return out;
}
inline std::string format_f64(double v) { return format_float<double>(v, 17, 16, -5); }
inline std::string format_f32(float v) { return format_float<float>(v, 9, 13, -6); }

class JsonWriter {
 public:
  std::string out;
  void begin_object() { open('{'); }
  void end_object() { close('}'); }
  void begin_array() { open('['); }
  void end_array() { close(']'); }
  void key(const std::string& k) { element(); out += '"'; escape(k); out += "\": "; pending_value_ = true; }
  void value_raw(const std::string& v) { element(); out += v; }
  void value_u64(uint64_t v) { value_raw(std::to_string(v)); }
  void value_f64(double v) { value_raw(format_f64(v)); }
  void value_f32(float v) { value_raw(format_f32(v)); }
  void value_null() { value_raw("null"); }
  void value_str(const std::string& s) { element(); out += '"'; escape(s); out += '"'; }

 private:
  struct Level { bool empty = true; };
  std::string indent_;
  std::vector<Level> stack_;
  bool pending_value_ = false;
  void element() {
    if (pending_value_) { pending_value_ = false; return; }
    if (!stack_.empty()) {
      out += stack_.back().empty ? "\n" : ",\n";
      stack_.back().empty = false;
      out += indent_;
    }
  }
  void open(char c) { element(); out += c; stack_.push_back({}); indent_ += "  "; }
  void close(char c) {
    indent_.resize(indent_.size() - 2);
    bool empty = stack_.back().empty;
    stack_.pop_back();
    if (!empty) { out += "\n"; out += indent_; }
    out += c;
  }
  void escape(const std::string& s) {
    for (char ch : s) {
      if (ch == '"' || ch == '\\') { out += '\\'; out += ch; }
      else if ((unsigned char)ch < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", ch); out += b; }
      else out += ch;
    }
  }
};

}  // namespace ngs
