// Host-side BAM/BAI handling for the CUDA qc path (reference: src/utils/formats/bam.rs:77-123,
// src/utils/pathbuf.rs:59-75).  BGZF framing (K1) and header parsing stay on the host as in the
// reference; the header blocks themselves are inflated on the GPU (ngsq_inflate_to_host), so no
// CPU inflate exists anywhere on this path.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ngs_cuda.h"
#include "facets.hpp"

namespace ngs {

class MappedFile {
 public:
  explicit MappedFile(const std::string& path) {
    fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (fstat(fd_, &st)) throw std::runtime_error("cannot stat " + path);
    size_ = (size_t)st.st_size;
    if (size_) {
      data_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (data_ == MAP_FAILED) throw std::runtime_error("cannot mmap " + path);
      madvise((void*)data_, size_, MADV_SEQUENTIAL);
    }
  }
  ~MappedFile() { if (data_ && size_) munmap((void*)data_, size_); if (fd_ >= 0) close(fd_); }
  const uint8_t* data() const { return data_; }
  size_t size() const { return size_; }

 private:
  int fd_ = -1;
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
};

struct BamHeader {
  std::string text;
  std::vector<ReferenceSequence> reference_sequences;  // binary reference list (bam.rs:109-111)
  uint64_t first_record_voffset = 0;
};

inline BamHeader read_bam_header(ngsq_engine* e, const uint8_t* bam, size_t size) {
  size_t take = 1 << 16;
  for (;;) {
    size_t n = std::min(take, size);
    uint32_t nb = 0;
    size_t used = 0;
    if (ngsq_bgzf_walk(bam, n, 0, nullptr, 0, &nb, &used)) throw std::runtime_error("malformed BGZF framing at the start of the file");
    std::vector<ngsq_block> blk(nb ? nb : 1);
    ngsq_bgzf_walk(bam, n, 0, blk.data(), nb, &nb, &used);
    std::vector<uint8_t> buf((size_t)nb * 65536 + 16);
    size_t got = 0;
    if (used) check(e, ngsq_inflate_to_host(e, bam, used, buf.data(), buf.size(), &got));
    auto need_more = [&]() {
      if (n == size) throw std::runtime_error("truncated BAM header");
      take *= 4;
    };
    auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, &buf[o], 4); return v; };
    if (got < 12) { need_more(); continue; }
    if (memcmp(buf.data(), "BAM\1", 4)) throw std::runtime_error("not a BAM file (bad magic)");
    size_t o = 8 + (size_t)u32(4);
    if (got < o + 4) { need_more(); continue; }
    BamHeader h;
    h.text.assign((const char*)&buf[8], o - 8);
    uint32_t n_ref = u32(o);
    o += 4;
    bool short_read = false;
    for (uint32_t i = 0; i < n_ref; ++i) {
      if (got < o + 4) { short_read = true; break; }
      uint32_t l_name = u32(o);
      o += 4;
      if (got < o + l_name + 4) { short_read = true; break; }
      ReferenceSequence rs;
      rs.name.assign((const char*)&buf[o], l_name ? l_name - 1 : 0);
      o += l_name;
      rs.length = u32(o);
      o += 4;
      h.reference_sequences.push_back(rs);
    }
    if (short_read) { need_more(); continue; }
    uint64_t acc = 0;
    bool found = false;
    for (uint32_t i = 0; i < nb; ++i) {
      if (o < acc + blk[i].isize) { h.first_record_voffset = (blk[i].coffset << 16) | (o - acc); found = true; break; }
      acc += blk[i].isize;
    }
    if (!found) h.first_record_voffset = (uint64_t)used << 16;
    return h;
  }
}

// BAI (SAM spec 5.2): only what sharding needs — per-reference file extents from pseudo-bin 37450.
struct BaiReference { bool has_extent = false; uint64_t ref_beg = 0, ref_end = 0, n_mapped = 0, n_unmapped = 0; uint32_t n_bins = 0, n_intv = 0; };
struct BaiIndex { std::vector<BaiReference> refs; bool has_n_no_coor = false; uint64_t n_no_coor = 0; };

// IndexCheck::Full (utils/formats/bam.rs:86-96): the index must exist next to the BAM and parse.
inline BaiIndex read_bai(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Could not find BAM index at " + path + ". Please index the BAM first (`ngs index`).");
  std::vector<uint8_t> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t o = 0;
  auto need = [&](size_t k) { if (o + k > b.size()) throw std::runtime_error("reading BAM index: truncated"); };
  need(8);
  if (memcmp(b.data(), "BAI\1", 4)) throw std::runtime_error("reading BAM index: bad magic");
  uint32_t n_ref;
  memcpy(&n_ref, &b[4], 4);
  o = 8;
  BaiIndex idx;
  for (uint32_t r = 0; r < n_ref; ++r) {
    BaiReference R;
    need(4);
    memcpy(&R.n_bins, &b[o], 4);
    o += 4;
    for (uint32_t k = 0; k < R.n_bins; ++k) {
      need(8);
      uint32_t bin, n_chunk;
      memcpy(&bin, &b[o], 4);
      memcpy(&n_chunk, &b[o + 4], 4);
      o += 8;
      need(16ull * n_chunk);
      if (bin == 37450 && n_chunk == 2) {
        memcpy(&R.ref_beg, &b[o], 8); memcpy(&R.ref_end, &b[o + 8], 8);
        memcpy(&R.n_mapped, &b[o + 16], 8); memcpy(&R.n_unmapped, &b[o + 24], 8);
        R.has_extent = R.ref_end > R.ref_beg;
      }
      o += 16ull * n_chunk;
    }
    need(4);
    memcpy(&R.n_intv, &b[o], 4);
    o += 4;
    need(8ull * R.n_intv);
    o += 8ull * R.n_intv;
    idx.refs.push_back(R);
  }
  if (o + 8 <= b.size()) { idx.has_n_no_coor = true; memcpy(&idx.n_no_coor, &b[o], 8); }
  return idx;
}

// One contiguous range of the file owned by a shard.
struct Shard {
  uint64_t first_voffset = 0, end_voffset = 0;  // end 0 = to EOF
  std::vector<uint32_t> contigs;                // references whose coverage this shard owns
  bool empty = false;
};

// Contig-aligned cuts: contiguous runs of contigs balanced by compressed bytes, so
// each coverage position has exactly one owner and the only exchange is the final sum-reduce.
inline std::vector<Shard> plan_shards(const BamHeader& h, const BaiIndex& bai, uint32_t n_shards, uint64_t file_size) {
  struct Span { uint32_t ref; uint64_t beg, end; };
  std::vector<Span> spans;
  for (uint32_t c = 0; c < bai.refs.size(); ++c)
    if (bai.refs[c].has_extent) spans.push_back({c, bai.refs[c].ref_beg, bai.refs[c].ref_end});
  std::sort(spans.begin(), spans.end(), [](const Span& a, const Span& b) { return a.beg < b.beg; });
  std::vector<Shard> out;
  auto all_refs = [&]() { std::vector<uint32_t> v; for (uint32_t c = 0; c < h.reference_sequences.size(); ++c) v.push_back(c); return v; };
  if (n_shards <= 1 || spans.size() < 2) {
    Shard s;
    s.first_voffset = h.first_record_voffset;
    s.contigs = all_refs();
    out.push_back(s);
  } else {
    const uint64_t base = h.first_record_voffset >> 16;
    const double total = (double)(file_size - base);
    Shard cur;
    cur.first_voffset = h.first_record_voffset;
    for (size_t i = 0; i < spans.size(); ++i) {
      cur.contigs.push_back(spans[i].ref);
      size_t contigs_left = spans.size() - i - 1;
      size_t shards_left = n_shards - out.size() - 1;
      double done = (double)((spans[i].end >> 16) - base);
      double target = total * (double)(out.size() + 1) / (double)n_shards;
      if (shards_left > 0 && contigs_left > 0 && (done >= target || contigs_left <= shards_left)) {
        cur.end_voffset = spans[i + 1].beg;
        out.push_back(cur);
        cur = Shard();
        cur.first_voffset = spans[i + 1].beg;
      }
    }
    out.push_back(cur);  // last shard runs to EOF: covers the unplaced-unmapped tail
    // references without records: coverage never touches them; give them to shard 0 for completeness
  }
  while (out.size() < n_shards) { Shard s; s.empty = true; out.push_back(s); }
  return out;
}

// Compressed byte range [lo, hi) a shard submits: through the block that holds end_voffset.
inline void shard_bytes(const Shard& s, const uint8_t* bam, uint64_t file_size, uint64_t* lo, uint64_t* hi) {
  *lo = s.first_voffset >> 16;
  if (s.end_voffset == 0) { *hi = file_size; return; }
  uint64_t co = s.end_voffset >> 16, uo = s.end_voffset & 0xFFFF;
  if (uo == 0) { *hi = co; return; }
  ngsq_block b;
  uint32_t n = 0;
  size_t used = 0;
  if (ngsq_bgzf_walk(bam + co, std::min<uint64_t>(file_size - co, 1 << 17), co, &b, 1, &n, &used) || n == 0)
    throw std::runtime_error("shard end does not address a BGZF block");
  *hi = co + b.csize;
}

}  // namespace ngs
