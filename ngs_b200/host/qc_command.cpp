// ngs-cuda-qc — host driver of the CUDA `ngs qc` path.  Mirrors the reference's subcommand
// (src/qc/command.rs): QcArgs (36-102), qc() (109-218), app() (226-421).  Same positional
// arguments and flags, same checks in the same order, same output file; the two hot loops of
// app() (305-316 and 350-397) are replaced by one pass of the CUDA engine per GPU.
//
//   ngs-cuda-qc qc <BAM> <REFERENCE_GENOME> [-n N] [-o DIR] [-p PREFIX] [--only FACET]
//               [--cuda-devices 0,1,..] [--cuda-gc-seed S] [--cuda-no-crc] [--cuda-chunk-mb M] [--cuda-buffered-io] [--cuda-perf]
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <optional>
#include <sstream>
#include <chrono>
#include <thread>

#include <fcntl.h>
#include <unistd.h>

#include "bam.hpp"
#include "facets.hpp"

using namespace ngs;

namespace {

// command.rs:36-102
struct QcArgs {
  std::string src;
  std::string reference_genome;
  std::optional<std::string> features_gff, reference_fasta, output_prefix, output_directory, only_facet, vaf_file_path;
  std::optional<uint64_t> num_records;
  FeatureNames feature_names;  // command.rs:78-101
  // engine-only knobs (not part of the reference CLI)
  std::vector<int> devices{0};
  uint64_t gc_seed = 0;
  bool verify_crc = true;
  size_t chunk_mb = 64;
  bool direct_io = true;
  bool perf = false;
};

void info(const std::string& s) { std::cerr << "INFO " << s << "\n"; }

// "  [*] Processed 1,000,000 records." (display.rs:43-52: RecordCounter::inc logs every millionth record; Locale::en)
std::string with_commas(uint64_t v) {
  std::string d = std::to_string(v), out;
  for (size_t i = 0; i < d.size(); ++i) {
    out += d[i];
    const size_t left = d.size() - 1 - i;
    if (left && left % 3 == 0) out += ',';
  }
  return out;
}

struct Progress {
  uint64_t logged = 0;  // millions already reported
  void report(ngsq_engine* e) {
    uint64_t n = 0;
    if (ngsq_progress(e, &n)) return;
    for (; (logged + 1) * 1000000ull <= n; ++logged) info("  [*] Processed " + with_commas((logged + 1) * 1000000ull) + " records.");
  }
};

// The shard's bytes, read with O_DIRECT into a ring of page-locked buffers and handed to the engine chunk by chunk:
// the file read of chunk k+2, the PCIe copy of chunk k+1 and the kernels of chunk k overlap, and the page cache is
// left alone (the reference reads through a BufReader, bam.rs:41-44).  A chunk ends on a BGZF block boundary; the
// partial block behind it moves in front of the next read.
class ChunkReader {
 public:
  static constexpr size_t kLead = 1 << 17;  // room in front of every buffer for the previous chunk's partial block
  ChunkReader(const std::string& path, size_t chunk_bytes, bool direct) : chunk_((chunk_bytes + 4095) & ~size_t(4095)) {
    fd_ = -1;
#ifdef O_DIRECT
    if (direct) fd_ = open(path.c_str(), O_RDONLY | O_DIRECT);
    direct_ = fd_ >= 0;
#endif
    if (fd_ < 0) fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("cannot open " + path);
    for (auto& b : buf_) {
      b = (uint8_t*)ngsq_host_alloc(kLead + chunk_ + 4096);
      if (!b) throw std::runtime_error("cannot allocate pinned staging buffers");
    }
  }
  ~ChunkReader() {
    for (auto& b : buf_) if (b) ngsq_host_free(b);
    if (fd_ >= 0) close(fd_);
  }
  bool direct() const { return direct_; }

  // streams file bytes [lo, hi) (lo on a block boundary) into e
  void run(ngsq_engine* e, uint64_t lo, uint64_t hi, Progress* progress) {
    uint64_t pos = lo & ~uint64_t(4095);  // aligned file position of the next read
    size_t skip = (size_t)(lo - pos);     // bytes of the first read that precede the shard
    size_t carry = 0;                     // bytes of a partial block kept from the previous chunk
    const uint8_t* carry_src = nullptr;
    uint32_t submits = 0;
    int32_t used_by[kBufs];
    for (auto& u : used_by) u = -1;
    uint64_t file_off = lo;               // file offset of the next byte to submit
    for (uint32_t k = 0; pos < hi; ++k) {
      const uint32_t bi = k % kBufs;
      if (used_by[bi] >= 0) check(e, ngsq_wait_copied(e, (uint32_t)used_by[bi]));  // its last copy must have left the buffer
      uint8_t* base = buf_[bi] + kLead;
      size_t want = (size_t)std::min<uint64_t>(chunk_, ((hi + 4095) & ~uint64_t(4095)) - pos);
      size_t got = 0;
      while (got < want) {
        ssize_t r = pread(fd_, base + got, want - got, (off_t)(pos + got));
        if (r < 0) {
          if (direct_ && got == 0) {  // a file system that refuses O_DIRECT reads: fall back to buffered reads once
            int fd2 = open_again_buffered();
            if (fd2 >= 0) { close(fd_); fd_ = fd2; direct_ = false; continue; }
          }
          throw std::runtime_error("read error on the BAM file");
        }
        if (r == 0) break;
        got += (size_t)r;
      }
      uint8_t* begin = base + skip;
      size_t n = got > skip ? got - skip : 0;
      if (pos + got > hi) n -= (size_t)std::min<uint64_t>(n, pos + got - hi);
      if (carry) {
        begin -= carry;
        memcpy(begin, carry_src, carry);
        n += carry;
      }
      pos += got;
      skip = 0;
      const bool last = pos >= hi || got < want;
      uint32_t nb = 0;
      size_t used = 0;
      if (ngsq_bgzf_walk(begin, n, file_off, nullptr, 0, &nb, &used)) throw std::runtime_error("malformed BGZF framing");
      if (last && used != n) throw std::runtime_error("truncated BGZF block at end of file");
      if (used) {
        check(e, ngsq_submit(e, begin, used, file_off));
        used_by[bi] = (int32_t)submits++;
        file_off += used;
      }
      carry = n - used;
      carry_src = begin + used;
      if (carry > kLead) throw std::runtime_error("BGZF block larger than 128 KiB");
      if (progress) progress->report(e);
      if (got < want) break;
    }
  }

 private:
  int open_again_buffered() {
    char link[64], path[4096];
    snprintf(link, sizeof link, "/proc/self/fd/%d", fd_);
    ssize_t n = readlink(link, path, sizeof path - 1);
    if (n <= 0) return -1;
    path[n] = 0;
    return open(path, O_RDONLY);
  }
  static constexpr int kBufs = 3;
  int fd_ = -1;
  bool direct_ = false;
  size_t chunk_;
  uint8_t* buf_[kBufs] = {nullptr, nullptr, nullptr};
};

void run_shard(ngsq_engine* e, const std::string& path, const MappedFile& file, const Shard& shard, size_t chunk_bytes, bool direct, bool log_progress,
               std::string* err) {
  try {
    if (shard.empty) { check(e, ngsq_set_range(e, 0, 0)); check(e, ngsq_finish(e)); return; }
    uint64_t lo, hi;
    shard_bytes(shard, file.data(), file.size(), &lo, &hi);
    ChunkReader reader(path, chunk_bytes, direct);
    for (int attempt = 0;; ++attempt) {
      check(e, ngsq_set_range(e, shard.first_voffset, shard.end_voffset));
      Progress progress;
      reader.run(e, lo, hi, log_progress ? &progress : nullptr);
      const int rc = ngsq_finish(e);
      if (rc == NGSQ_E_QUAL_CAP && attempt == 0) {
        // a read longer than the quality-by-position table (131072 positions by default): size the table for the
        // longest read the scan met and stream the shard again — the engine keeps nothing of a file it has streamed
        ngsq_stats st;
        check(e, ngsq_get_stats(e, &st));
        info("  [*] Reads of up to " + std::to_string(st.max_read_len) + " bases: enlarging the quality table and reading the file again.");
        check(e, ngsq_reset(e));
        check(e, ngsq_set_quality_positions(e, (uint32_t)st.max_read_len + 1));
        continue;
      }
      check(e, rc);
      if (log_progress) progress.report(e);
      break;
    }
  } catch (const std::exception& ex) {
    *err = ex.what();
  }
}

// command.rs:226-421
int app(const QcArgs& args, const ReferenceGenome& genome, const std::string& output_prefix, const std::string& output_directory) {
  // open_and_parse(src, IndexCheck::Full): bam.rs:77-123
  if (args.src.size() < 4 || args.src.substr(args.src.size() - 4) != ".bam")
    throw std::runtime_error("could not open " + args.src + ": expected a .bam file");
  MappedFile file(args.src);
  BaiIndex bai = read_bai(args.src + ".bai");  // pathbuf.rs:59-75: x.bam -> x.bam.bai
  const uint32_t n_dev = (uint32_t)args.devices.size();

  // facets first (cheap) so `--only` errors surface before any device work.  The two optional facets read their
  // inputs here, like GenomicFeaturesFacet::try_from / EditsFacet::try_from do inside get_qc_facets (qc.rs:68-94).
  std::unique_ptr<GenomicFeaturesFacet> features_facet;
  std::unique_ptr<EditsFacet> edits_facet;
  if (args.features_gff) features_facet = std::make_unique<GenomicFeaturesFacet>(GenomicFeaturesFacet::try_from(*args.features_gff, args.feature_names, genome));
  if (args.reference_fasta) edits_facet = std::make_unique<EditsFacet>(EditsFacet::try_from(*args.reference_fasta, args.vaf_file_path));
  FacetSet facets = get_qc_facets(genome, args.only_facet, std::move(features_facet), std::move(edits_facet));
  GenomicFeaturesFacet* features = nullptr;
  EditsFacet* edits = nullptr;
  bool coverage = false, record_defaults = false;
  for (auto& f : facets.record_based) { if (auto* g = dynamic_cast<GenomicFeaturesFacet*>(f.get())) features = g; else record_defaults = true; }
  for (auto& f : facets.sequence_based) { if (auto* g = dynamic_cast<EditsFacet*>(f.get())) edits = g; else coverage = true; }
  uint32_t flags = 0;
  if (record_defaults) flags |= NGSQ_F_RECORD_FACETS;
  if (coverage) flags |= NGSQ_F_COVERAGE;
  if (features) flags |= NGSQ_F_FEATURES;
  if (edits) flags |= NGSQ_F_EDITS;
  if (args.verify_crc) flags |= NGSQ_F_VERIFY_CRC;
  if (args.num_records && n_dev > 1) throw std::runtime_error("-n needs a single device (records are counted in file order)");

  std::vector<ngsq_engine*> engines(n_dev, nullptr);
  struct Cleanup { std::vector<ngsq_engine*>& v; ~Cleanup() { for (auto* e : v) if (e) ngsq_destroy(e); } } cleanup{engines};
  for (uint32_t i = 0; i < n_dev; ++i) {
    ngsq_config cfg{};
    cfg.struct_size = sizeof cfg;
    cfg.flags = flags;
    cfg.gc_seed = args.gc_seed;
    cfg.max_records = args.num_records.value_or(0);
    if (ngsq_create(args.devices[i], &cfg, &engines[i])) throw std::runtime_error(std::string("CUDA engine: ") + ngsq_last_error(nullptr));
  }
  BamHeader header = read_bam_header(engines[0], file.data(), file.size());

  if (!std::filesystem::exists(output_directory)) std::filesystem::create_directories(output_directory);

  // reference sequence concordance check: command.rs:258-272
  std::vector<Sequence> supported = get_all_sequences(genome);
  for (auto& rs : header.reference_sequences) {
    bool ok = false;
    for (auto& s : supported) if (s.name == rs.name) { ok = true; break; }
    if (!ok) throw std::runtime_error("Sequence \"" + rs.name + "\" not found in specified reference genome. Did you set the correct reference genome?");
  }

  for (auto& f : facets.record_based) info(std::string("  [*] ") + f->name() + ", " + to_string(f->computational_load()));
  for (auto& f : facets.sequence_based) info(std::string("  [*] ") + f->name() + ", " + to_string(f->computational_load()));

  // shards: contig-aligned file ranges from the BAI
  std::vector<Shard> shards = plan_shards(header, bai, n_dev, file.size());
  std::vector<uint32_t> ref_len;
  for (auto& rs : header.reference_sequences) ref_len.push_back(rs.length);
  for (uint32_t i = 0; i < n_dev; ++i) {
    // the same mask on every engine: the packed result buffer (the NCCL payload) is laid out from it, and a
    // contig outside an engine's shard simply stays untouched there (ngsq_reduce rejects differing layouts)
    std::vector<uint8_t> enabled(ref_len.size(), 0);
    for (auto& f : facets.sequence_based)
      for (uint32_t c = 0; c < ref_len.size(); ++c)
        if (f.get() != edits && f->supports_sequence_name(header.reference_sequences[c].name)) enabled[c] = 1;  // the Coverage mask
    check(engines[i], ngsq_set_references(engines[i], (uint32_t)ref_len.size(), ref_len.data(), enabled.data()));
    if (features) features->upload(engines[i], header.reference_sequences);
    if (edits) edits->upload(engines[i], header.reference_sequences);
  }
  if (n_dev > 1) {
    char id[128];
    if (ngsq_nccl_unique_id(id)) throw std::runtime_error(std::string("NCCL: ") + ngsq_last_error(nullptr));
    std::vector<std::thread> ts;
    std::vector<std::string> errs(n_dev);
    for (uint32_t i = 0; i < n_dev; ++i)
      ts.emplace_back([&, i] { if (ngsq_comm_init(engines[i], (int)n_dev, (int)i, id)) errs[i] = ngsq_last_error(engines[i]); });
    for (auto& t : ts) t.join();
    for (auto& s : errs) if (!s.empty()) throw std::runtime_error("NCCL: " + s);
  }

  // the hot path: one thread per GPU
  info("Starting CUDA pass for QC stats.");
  const auto t_hot = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> ts;
    std::vector<std::string> errs(n_dev);
    for (uint32_t i = 0; i < n_dev; ++i)
      ts.emplace_back(run_shard, engines[i], std::cref(args.src), std::cref(file), std::cref(shards[i]), args.chunk_mb << 20, args.direct_io, i == 0, &errs[i]);
    for (auto& t : ts) t.join();
    for (auto& s : errs) if (!s.empty()) throw std::runtime_error(s);
  }
  if (n_dev > 1) {
    std::vector<std::thread> ts;
    std::vector<std::string> errs(n_dev);
    for (uint32_t i = 0; i < n_dev; ++i) ts.emplace_back([&, i] { if (ngsq_reduce(engines[i], 0)) errs[i] = ngsq_last_error(engines[i]); });
    for (auto& t : ts) t.join();
    for (auto& s : errs) if (!s.empty()) throw std::runtime_error("NCCL: " + s);
  }
  const double hot_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_hot).count();
  ngsq_engine* root = engines[0];
  uint64_t n_records = 0;
  for (auto* e : engines) { ngsq_stats st; ngsq_get_stats(e, &st); n_records += st.records; }
  {
    std::ostringstream os;
    os << "Processed " << n_records << " records.";
    info(os.str());
  }

  // first pass: summarize (command.rs:328-330)
  for (auto& f : facets.record_based) { f->ingest(root); f->summarize(); }
  // second pass: per sequence in header order setup -> (process on the device) -> teardown (command.rs:356-396)
  if (edits) edits->set_engines(engines);  // the VAF file reads every engine's per-position counters
  for (auto& f : facets.sequence_based) f->ingest_global(root);
  for (uint32_t c = 0; c < header.reference_sequences.size(); ++c) {
    const ReferenceSequence& seq = header.reference_sequences[c];
    for (auto& f : facets.sequence_based) {
      if (!f->supports_sequence_name(seq.name)) continue;
      f->setup(seq);
      f->ingest(root, c, seq);
      f->teardown(seq);
    }
  }
  // finalize: command.rs:406-418
  info("Aggregating results.");
  Results results;
  for (auto& f : facets.record_based) f->aggregate(results);
  for (auto& f : facets.sequence_based) f->aggregate(results);
  info("Writing output.");
  results.write(output_prefix, output_directory);

  if (args.perf) {
    std::ofstream pf(output_directory + "/" + output_prefix + ".perf.json");
    pf << "[";
    for (uint32_t i = 0; i < n_dev; ++i) {
      ngsq_stats st;
      ngsq_get_stats(engines[i], &st);
      pf << (i ? "," : "") << "{\"device\":" << args.devices[i] << ",\"records\":" << st.records << ",\"blocks\":" << st.blocks
         << ",\"compressed_bytes\":" << st.compressed_bytes << ",\"inflated_bytes\":" << st.inflated_bytes << ",\"ms_inflate\":" << st.ms_inflate
         << ",\"ms_crc\":" << st.ms_crc << ",\"ms_scan\":" << st.ms_scan << ",\"ms_facets\":" << st.ms_facets << ",\"ms_coverage\":" << st.ms_coverage
         << ",\"ms_total\":" << st.ms_total << ",\"ms_tail\":" << st.ms_tail << ",\"waves\":" << st.waves << ",\"wall_ms_file_to_results\":" << hot_ms << "}";
    }
    pf << "]\n";
  }
  return 0;
}

// command.rs:109-218
int qc(const QcArgs& args) {
  info("Starting qc command...");
  auto genome = get_reference_genome(args.reference_genome);
  if (!genome)
    throw std::runtime_error("reference genome is not supported: " + args.reference_genome +
                             ". Did you set the correct reference genome?. Use the `list genomes` subcommand to see supported reference genomes.");
  std::string prefix = args.output_prefix.value_or(std::filesystem::path(args.src).filename().string());
  std::string outdir = args.output_directory.value_or(std::filesystem::current_path().string());
  return app(args, *genome, prefix, outdir);
}

void usage() {
  std::cerr << "usage: ngs-cuda-qc qc <BAM> <REFERENCE_GENOME> [-n N] [-o DIR] [-p PREFIX] [--only FACET]\n"
               "                  [--cuda-devices 0,1,..] [--cuda-gc-seed S] [--cuda-no-crc] [--cuda-chunk-mb M] [--cuda-buffered-io] [--cuda-perf]\n";
}

}  // namespace

extern "C" int ngs_cuda_qc_main(int argc, char** argv) {
  try {
    QcArgs a;
    std::vector<std::string> pos;
    int i = 1;
    if (i < argc && std::string(argv[i]) == "qc") ++i;
    for (; i < argc; ++i) {
      std::string s = argv[i];
      auto val = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value for " + s); return argv[++i]; };
      if (s == "-n" || s == "--num-records") a.num_records = std::stoull(val());
      else if (s == "-o" || s == "--output-directory") a.output_directory = val();
      else if (s == "-p" || s == "--output-prefix") a.output_prefix = val();
      else if (s == "-f" || s == "--features-gff") a.features_gff = val();
      else if (s == "-r" || s == "--reference-fasta") a.reference_fasta = val();
      else if (s == "--only") a.only_facet = val();
      else if (s == "--vaf-file" || s == "--vaf-file-path") a.vaf_file_path = val();
      else if (s == "--five-prime-utr-feature-name") a.feature_names.slot[0] = val();
      else if (s == "--three-prime-utr-feature-name") a.feature_names.slot[1] = val();
      else if (s == "--coding-sequence-feature-name") a.feature_names.slot[2] = val();
      else if (s == "--exon-feature-name") a.feature_names.slot[3] = val();
      else if (s == "--gene-feature-name") a.feature_names.slot[4] = val();
      else if (s == "--cuda-devices") { a.devices.clear(); std::stringstream ss(val()); std::string t; while (std::getline(ss, t, ',')) a.devices.push_back(std::stoi(t)); }
      else if (s == "--cuda-gc-seed") a.gc_seed = std::stoull(val(), nullptr, 0);
      else if (s == "--cuda-no-crc") a.verify_crc = false;
      else if (s == "--cuda-chunk-mb") a.chunk_mb = std::stoull(val());
      else if (s == "--cuda-perf") a.perf = true;
      else if (s == "--cuda-buffered-io") a.direct_io = false;
      else if (s == "-h" || s == "--help") { usage(); return 0; }
      else if (!s.empty() && s[0] == '-' && s != "-") throw std::runtime_error("unknown flag " + s);
      else pos.push_back(s);
    }
    if (pos.size() != 2 || a.devices.empty()) { usage(); return 2; }
    a.src = pos[0];
    a.reference_genome = pos[1];
    return qc(a);
  } catch (const std::exception& ex) {
    std::cerr << "Error: " << ex.what() << "\n";
    return 1;
  }
}

#ifndef NGS_CUDA_QC_NO_MAIN
int main(int argc, char** argv) { return ngs_cuda_qc_main(argc, argv); }
#endif
