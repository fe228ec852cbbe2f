// libngs_cuda.so — implementation of include/ngs_cuda.h.
// One engine = one GPU = one host thread.  Device-resident layout (sized for 180 GB HBM3e):
//   compressed BGZF bytes | inflated byte stream (contiguous, blocks at ISIZE prefix sums)
//   | block tables | record offset table (8 B/record) | per-contig int32 difference arrays
//   | packed u64 result buffer (the NCCL reduce payload).
// Streams: H2D copies on s_copy, kernels on s_comp, the per-block CRC32 check on s_aux (it depends only on the
// inflated bytes, so it runs beside the record scan and the facet kernel); each submitted chunk's inflate launch waits
// only for its own copy, so PCIe transfer of chunk k+1 overlaps inflate of chunk k.
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ngs_cuda.h"
#include "coverage.cuh"
#include "crc32.cuh"
#include "edits.cuh"
#include "facets.cuh"
#include "features.cuh"
#include "inflate2.cuh"
#include "recscan.cuh"

using namespace ngsq;

namespace {

thread_local std::string g_create_err;

struct DevFlags {
  ScanErr scan;
  uint32_t inflate_err;  // bit (1 << kBlk*) per failure kind
  uint32_t crc_bad;
  uint64_t n_records;
};

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  void* CommInitRank = nullptr;  // ncclCommInitRank(ncclComm_t*, int, ncclUniqueId by value, int)
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char b[128]; };

NcclApi g_nccl;
bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
  auto sym = [&](const char* s) { return dlsym(g_nccl.lib, s); };
  g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = sym("ncclCommInitRank");
  g_nccl.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))sym("ncclReduce");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Reduce || !g_nccl.AllReduce) {
    err = "libnccl is missing symbols";
    return false;
  }
  return true;
}

}  // namespace

struct ngsq_engine {
  int device = 0;
  int n_sm = 0;
  ngsq_config cfg{};
  std::string err;
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_aux = nullptr;
  std::vector<cudaEvent_t> copy_events;
  std::vector<uint32_t> copy_upto;  // h_blocks.size() once the chunk of copy_events[i] was appended
  uint32_t launched = 0;            // blocks [0, launched) have been handed to the inflate kernels
  struct InflateEvents { cudaEvent_t begin, decoded_from, decoded, end, crc_begin, crc_end; };
  std::vector<InflateEvents> inflate_events;
  cudaEvent_t ev_start = nullptr, ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_e = nullptr, ev_f = nullptr;
  bool run_started = false, finished = false;

  // references / coverage
  uint32_t n_ref = 0;
  std::vector<uint32_t> ref_len;
  std::vector<uint8_t> cov_enabled;
  std::vector<uint64_t> diff_base;
  std::vector<uint32_t> cov_slot, cov_nbins;
  uint32_t* d_ref_len = nullptr;
  uint8_t* d_cov_enabled = nullptr;
  uint64_t* d_diff_base = nullptr;
  uint32_t* d_cov_slot = nullptr;
  int32_t* d_diff = nullptr;
  uint64_t diff_elems = 0;
  int64_t* d_tile = nullptr;
  uint32_t tile_cap = 0;

  // results
  uint64_t* d_res = nullptr;
  size_t res_words = 0, res_cap_words = 0;
  uint32_t qual_off = R_FIXED_WORDS;
  uint32_t qpos_cap = 0;
  std::vector<uint64_t> h_res;
  uint32_t h_qpos = 0;

  // input / inflate
  struct CompSeg { uint8_t* ptr; size_t cap, used; };
  std::vector<CompSeg> comp_segs;  // compressed bytes live in segments: descriptors hold absolute addresses
  uint8_t* d_out = nullptr;
  size_t out_cap = 0;
  uint64_t out_used = 0;
  BlockDesc* d_blocks = nullptr;
  uint32_t blocks_cap = 0;
  std::vector<BlockDesc> h_blocks;
  std::vector<uint64_t> h_coff, h_out_off;
  std::vector<uint32_t> h_crc;
  uint64_t comp_bytes_total = 0;
  uint32_t* d_status = nullptr;
  uint32_t* d_crcx = nullptr;    // expected CRC32 of every block
  size_t crcx_cap = 0;
  uint32_t* d_bitmap = nullptr;  // inflate: one bit per inflated byte ("a match starts here"), kBitmapWords per block
  size_t bitmap_cap = 0;         // blocks
  uint32_t* d_queue = nullptr;
  uint64_t* d_agree = nullptr;  // 3 words exchanged before the reduce
  uint32_t n_launches = 0;
  static constexpr uint32_t kQueueSlots = 4096;
  uint64_t *d_out_off = nullptr, *d_coff = nullptr, *d_base = nullptr, *d_rec = nullptr;
  uint32_t *d_first = nullptr, *d_landed = nullptr, *d_count = nullptr;
  uint32_t aux_cap = 0;
  uint64_t rec_cap = 0;
  DevFlags* d_flags = nullptr;
  CrcTables* d_crc_tables = nullptr;

  uint64_t first_voff = 0, end_voff = 0;
  bool range_set = false;
  ngsq_stats stats{};
  uint32_t other_launches = 0;

  // Edits facet (NGSQ_F_EDITS): per-contig FASTA codes, per-position counters, result block
  std::vector<EditsContig> ed_contigs;
  std::vector<void*> ed_allocs;          // device allocations behind ed_contigs
  EditsContig* d_ed_contigs = nullptr;
  uint32_t *d_ed_refs = nullptr, *d_ed_alts = nullptr;
  uint64_t ed_pos_total = 0;
  unsigned long long* d_ed_res = nullptr;
  std::vector<uint64_t> h_ed_res;
  cudaEvent_t ev_g = nullptr;

  // Genomic Features facet (NGSQ_F_FEATURES): per contig and class, sorted starts / stops; nine counters
  std::vector<FeatureContig> ft_contigs;
  std::vector<void*> ft_allocs;
  FeatureContig* d_ft_contigs = nullptr;
  uint8_t ft_slot_class[8] = {0, 1, 2, 3, 4, 0, 0, 0};
  bool ft_model_set = false;
  unsigned long long* d_ft_res = nullptr;
  std::vector<uint64_t> h_ft_res;

  // nccl
  void* comm = nullptr;
  int n_ranks = 1, rank = 0;
};

namespace {

int fail(ngsq_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_err = buf;
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t _rc = (call);                                                                          \
    if (_rc != cudaSuccess) return fail(e, NGSQ_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_rc), __FILE__, __LINE__); \
  } while (0)

template <class T>
int grow(ngsq_engine* e, T*& ptr, size_t& cap, size_t need, size_t keep, cudaStream_t s, size_t pad = 0) {
  if (need <= cap) return NGSQ_OK;
  size_t ncap = std::max(need, cap + cap / 2);
  T* np = nullptr;
  cudaError_t rc = cudaMalloc(&np, (ncap + pad) * sizeof(T));
  if (rc != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc(%zu bytes): %s", (ncap + pad) * sizeof(T), cudaGetErrorString(rc));
  if (ptr) CU(cudaStreamSynchronize(e->s_aux));  // CRC kernels may still read the old buffer
  if (ptr && keep) {
    CU(cudaStreamSynchronize(e->s_comp));
    CU(cudaMemcpyAsync(np, ptr, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
    CU(cudaStreamSynchronize(s));
  } else if (ptr) {
    CU(cudaStreamSynchronize(e->s_comp));
  }
  if (ptr) cudaFree(ptr);
  ptr = np;
  cap = ncap;
  return NGSQ_OK;
}

// K2: lane-per-block Huffman decode, then warp-per-block LZ77 resolve (inflate2.cuh)
int launch_inflate(ngsq_engine* e, const BlockDesc* blocks, uint32_t n, uint8_t* out, uint32_t* queue, uint32_t* status,
                   uint32_t* bitmap, cudaStream_t s, cudaEvent_t ev_decode_from = nullptr, cudaEvent_t ev_decoded = nullptr) {
  if (!n) return NGSQ_OK;
  CU(cudaMemsetAsync(bitmap, 0, (size_t)n * kBitmapWords * 4, s));
  const uint32_t per_cta = kDecThreads;
  uint32_t grid = std::min<uint32_t>((n + per_cta - 1) / per_cta, (uint32_t)e->n_sm);
  if (ev_decode_from) CU(cudaEventRecord(ev_decode_from, s));
  inflate_decode_kernel<<<grid, kDecThreads, kDecSmem, s>>>(out, blocks, n, queue, status, bitmap);
  CU(cudaGetLastError());
  if (ev_decoded) CU(cudaEventRecord(ev_decoded, s));
  // full occupancy (8 CTAs x 8 warps per SM): measured 39 ms / 40 M records vs 47 / 56 / 79 ms with 4 / 3 / 2 CTAs per SM —
  // latency hiding beats keeping the blocks in flight L2-resident
  uint32_t rgrid = std::min<uint32_t>((n + kResWarps - 1) / kResWarps, (uint32_t)e->n_sm * 8);
  inflate_resolve_kernel<<<rgrid, kResThreads, 0, s>>>(out, blocks, n, bitmap, status);
  CU(cudaGetLastError());
  return NGSQ_OK;
}

// The inflate kernel takes absolute device addresses in BlockDesc.in_off (comp == nullptr).
int start_run(ngsq_engine* e) {
  if (e->run_started) return NGSQ_OK;
  CU(cudaEventRecord(e->ev_start, e->s_comp));
  if (e->res_words) CU(cudaMemsetAsync(e->d_res, 0, e->res_cap_words * 8, e->s_comp));
  if ((e->cfg.flags & NGSQ_F_COVERAGE) && e->diff_elems) CU(cudaMemsetAsync(e->d_diff, 0, e->diff_elems * 4, e->s_comp));
  CU(cudaMemsetAsync(e->d_queue, 0, ngsq_engine::kQueueSlots * 4, e->s_comp));
  CU(cudaMemsetAsync(e->d_flags, 0, sizeof(DevFlags), e->s_comp));
  if ((e->cfg.flags & NGSQ_F_FEATURES) && e->d_ft_res) CU(cudaMemsetAsync(e->d_ft_res, 0, F_WORDS * 8, e->s_comp));
  if ((e->cfg.flags & NGSQ_F_EDITS) && e->d_ed_res) {
    CU(cudaMemsetAsync(e->d_ed_res, 0, E_WORDS * 8, e->s_comp));
    if (e->ed_pos_total) {
      CU(cudaMemsetAsync(e->d_ed_refs, 0, e->ed_pos_total * 4, e->s_comp));
      CU(cudaMemsetAsync(e->d_ed_alts, 0, e->ed_pos_total * 4, e->s_comp));
    }
  }
  e->run_started = true;
  return NGSQ_OK;
}

int layout_results(ngsq_engine* e, uint32_t qpos) {
  // fixed | per-contig coverage slots | quality table (last, so its size may differ per run)
  uint32_t off = R_FIXED_WORDS;
  e->cov_slot.assign(e->n_ref, 0);
  e->cov_nbins.assign(e->n_ref, 0);
  for (uint32_t c = 0; c < e->n_ref; ++c) {
    e->cov_slot[c] = off;
    if (e->cov_enabled[c]) {
      uint32_t L = e->ref_len[c];
      uint32_t nb = 1 + L / kCovBin + (L % kCovBin ? 1 : 0);
      e->cov_nbins[c] = nb;
      off += COV_BINS + nb + 1;
    }
  }
  e->qual_off = off;
  e->qpos_cap = qpos;
  e->res_words = (size_t)off + (size_t)qpos * 94;
  return NGSQ_OK;
}

int ensure_res(ngsq_engine* e, uint32_t qpos, bool keep) {
  if (qpos < 256) qpos = 256;
  if (qpos <= e->qpos_cap && e->d_res) return NGSQ_OK;
  size_t old_words = e->res_words;
  layout_results(e, qpos);
  if (e->res_words > e->res_cap_words) {
    uint64_t* np = nullptr;
    size_t ncap = e->res_words;
    cudaError_t rc = cudaMalloc(&np, ncap * 8);
    if (rc != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc results: %s", cudaGetErrorString(rc));
    CU(cudaMemsetAsync(np, 0, ncap * 8, e->s_comp));
    if (keep && e->d_res && old_words) CU(cudaMemcpyAsync(np, e->d_res, old_words * 8, cudaMemcpyDeviceToDevice, e->s_comp));
    CU(cudaStreamSynchronize(e->s_comp));
    if (e->d_res) cudaFree(e->d_res);
    e->d_res = np;
    e->res_cap_words = ncap;
  }
  if (e->d_cov_slot) CU(cudaMemcpyAsync(e->d_cov_slot, e->cov_slot.data(), e->n_ref * 4, cudaMemcpyHostToDevice, e->s_comp));
  return NGSQ_OK;
}

int append_blocks(ngsq_engine* e, const ngsq_block* blk, uint32_t n, uint64_t dev_base_addr, uint64_t first_coffset,
                  uint32_t* first_new, uint32_t* n_new) {
  *first_new = (uint32_t)e->h_blocks.size();
  for (uint32_t i = 0; i < n; ++i) {
    const ngsq_block& b = blk[i];
    if (b.csize < b.hdr_len + 8 || b.isize > 65536) return fail(e, NGSQ_E_BAD_BLOCK, "malformed BGZF block at file offset %llu", (unsigned long long)b.coffset);
    if (b.isize == 0) continue;  // empty blocks (incl. the EOF marker) carry nothing
    BlockDesc d;
    d.in_off = dev_base_addr + (b.coffset - first_coffset) + b.hdr_len;
    d.out_off = e->out_used;
    d.clen = b.csize - b.hdr_len - 8;
    d.isize = b.isize;
    e->h_blocks.push_back(d);
    e->h_coff.push_back(b.coffset);
    e->h_out_off.push_back(e->out_used);
    e->h_crc.push_back(b.crc32);
    e->out_used += b.isize;
  }
  *n_new = (uint32_t)e->h_blocks.size() - *first_new;
  return NGSQ_OK;
}

int inflate_new_blocks(ngsq_engine* e, uint32_t first_new, uint32_t n_new) {
  if (!n_new) return NGSQ_OK;
  size_t bcap = e->blocks_cap, total = e->h_blocks.size();
  {
    size_t cap = bcap;
    int rc = grow(e, e->d_blocks, cap, total, first_new, e->s_comp);
    if (rc) return rc;
    if (cap != bcap) {
      uint32_t* ns = nullptr;
      CU(cudaMalloc(&ns, cap * 4));
      CU(cudaMemsetAsync(ns, 0, cap * 4, e->s_comp));
      if (e->d_status) {
        // the verdicts of the blocks decoded so far move along (same stream as the launches that wrote them):
        // ngsq_finish reports a failed block whichever launch it was in
        if (first_new) CU(cudaMemcpyAsync(ns, e->d_status, (size_t)first_new * 4, cudaMemcpyDeviceToDevice, e->s_comp));
        CU(cudaStreamSynchronize(e->s_comp));
        CU(cudaStreamSynchronize(e->s_aux));
        cudaFree(e->d_status);
      }
      e->d_status = ns;
    }
    if (cap > e->crcx_cap) {  // expected CRC32 per block (trailer values), checked right after each launch
      CU(cudaStreamSynchronize(e->s_comp));
      CU(cudaStreamSynchronize(e->s_aux));
      if (e->d_crcx) cudaFree(e->d_crcx);
      e->d_crcx = nullptr;
      CU(cudaMalloc(&e->d_crcx, cap * 4));
      e->crcx_cap = cap;
    }
    e->blocks_cap = (uint32_t)cap;
  }
  {
    size_t cap = e->out_cap;
    int rc = grow(e, e->d_out, cap, (size_t)e->out_used, (size_t)e->h_out_off[first_new], e->s_comp, 64);
    if (rc) return rc;
    e->out_cap = cap;
  }
  CU(cudaMemcpyAsync(e->d_blocks + first_new, e->h_blocks.data() + first_new, n_new * sizeof(BlockDesc), cudaMemcpyHostToDevice, e->s_comp));
  if (e->n_launches >= ngsq_engine::kQueueSlots) return fail(e, NGSQ_E_ARG, "too many submits in one run (max %u)", ngsq_engine::kQueueSlots);
  ngsq_engine::InflateEvents ev{};
  for (cudaEvent_t* x : {&ev.begin, &ev.decoded_from, &ev.decoded, &ev.end, &ev.crc_begin, &ev.crc_end}) CU(cudaEventCreate(x));
  CU(cudaEventRecord(ev.begin, e->s_comp));
  int rc;
  {
    if (total > e->bitmap_cap) {
      // earlier submits' bitmaps are dead once their resolve kernels ran: no need to keep them
      CU(cudaStreamSynchronize(e->s_comp));
      if (e->d_bitmap) cudaFree(e->d_bitmap);
      e->d_bitmap = nullptr;
      size_t cap = std::max<size_t>(total, e->blocks_cap);
      cudaError_t r2 = cudaMalloc(&e->d_bitmap, cap * kBitmapWords * 4);
      if (r2 != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc match bitmap (%zu bytes): %s", cap * kBitmapWords * 4, cudaGetErrorString(r2));
      e->bitmap_cap = cap;
    }
    rc = launch_inflate(e, e->d_blocks + first_new, n_new, e->d_out, e->d_queue + e->n_launches, e->d_status + first_new,
                        e->d_bitmap + (size_t)first_new * kBitmapWords, e->s_comp, ev.decoded_from, ev.decoded);
    e->other_launches += 1;  // resolve kernel (the decode kernel is counted as the inflate launch)
  }
  if (rc) return rc;
  CU(cudaEventRecord(ev.end, e->s_comp));
  // CRC32 of the new blocks against their trailers (what noodles-bgzf checks per block), launched per
  // wave on its own stream: it overlaps the host-to-device copy of the next chunks, the next wave and,
  // for the last wave, the record scan and the facet kernel
  CU(cudaStreamWaitEvent(e->s_aux, ev.end, 0));
  CU(cudaEventRecord(ev.crc_begin, e->s_aux));
  if (e->cfg.flags & NGSQ_F_VERIFY_CRC) {
    CU(cudaMemcpyAsync(e->d_crcx + first_new, e->h_crc.data() + first_new, (size_t)n_new * 4, cudaMemcpyHostToDevice, e->s_aux));
    const uint32_t grid = std::min<uint32_t>((n_new + kCrcThreads / 32 - 1) / (kCrcThreads / 32), (uint32_t)e->n_sm * 6);
    crc32_kernel<<<grid, kCrcThreads, kCrcSmem, e->s_aux>>>(e->d_out, e->d_blocks + first_new, e->d_crcx + first_new, n_new, e->d_crc_tables,
                                                            &e->d_flags->crc_bad);
    CU(cudaGetLastError());
    e->other_launches++;
  }
  CU(cudaEventRecord(ev.crc_end, e->s_aux));
  e->inflate_events.push_back(ev);
  e->n_launches++;
  return NGSQ_OK;
}

// Hands blocks [launched, upto) to the inflate kernels.  Launches are sized in whole waves of the
// decode kernel (one BGZF block per lane, n_sm x kDecThreads lanes): a launch takes the time of its
// slowest lane, so many small launches would each pay a full block-decode latency.
int launch_pending(ngsq_engine* e, uint32_t upto) {
  if (upto <= e->launched) return NGSQ_OK;
  for (size_t i = 0; i < e->copy_upto.size(); ++i)
    if (e->copy_upto[i] >= upto) { CU(cudaStreamWaitEvent(e->s_comp, e->copy_events[i], 0)); break; }
  int rc = inflate_new_blocks(e, e->launched, upto - e->launched);
  if (rc) return rc;
  e->launched = upto;
  return NGSQ_OK;
}

// waits for the CRC kernels (own stream) and returns the number of blocks whose CRC32 did not match
int crc_verdict(ngsq_engine* e, uint32_t* n_bad) {
  CU(cudaStreamSynchronize(e->s_aux));
  CU(cudaMemcpy(n_bad, &e->d_flags->crc_bad, 4, cudaMemcpyDeviceToHost));
  return NGSQ_OK;
}

uint32_t launch_quantum(const ngsq_engine* e) {
  // One wave: a launch takes about one block-decode latency whatever its size (every lane decodes its
  // block serially), so smaller launches only add latencies; measured: half waves made e2e 40 % slower.
  // When the caller announced the number of blocks (reserve_blocks), the waves are made equal so that no
  // small remainder launch (a full latency for a few blocks) is left for ngsq_finish.
  if (e->cfg.launch_blocks) return e->cfg.launch_blocks;
  const uint32_t wave = (uint32_t)e->n_sm * kDecThreads;
  if (e->cfg.reserve_blocks > wave) {
    const uint32_t n_waves = (e->cfg.reserve_blocks + wave - 1) / wave;
    return (e->cfg.reserve_blocks + n_waves - 1) / n_waves;
  }
  return wave;
}

}  // namespace

extern "C" {

int ngsq_version(void) { return NGSQ_VERSION; }

const char* ngsq_last_error(ngsq_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int ngsq_create(int device, const ngsq_config* cfg, ngsq_engine** out) {
  ngsq_engine* e = nullptr;
  if (!out) return fail(e, NGSQ_E_ARG, "out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t rc = cudaGetDeviceCount(&n_dev);
  if (rc != cudaSuccess || n_dev == 0)
    return fail(e, NGSQ_E_CUDA, "no CUDA device available (%s); the ngs-cuda engine has no CPU fallback", cudaGetErrorString(rc));
  if (device < 0 || device >= n_dev) return fail(e, NGSQ_E_ARG, "device %d out of range (%d devices)", device, n_dev);
  ngsq_engine* ne = new ngsq_engine();
  ne->device = device;
  if (cfg) memcpy(&ne->cfg, cfg, std::min<size_t>(cfg->struct_size ? cfg->struct_size : sizeof(ngsq_config), sizeof(ngsq_config)));
  if (!ne->cfg.flags) ne->cfg.flags = NGSQ_F_RECORD_FACETS | NGSQ_F_COVERAGE;
  e = ne;
  auto bail = [&](int code) { std::string m = e->err; ngsq_destroy(e); g_create_err = m; return code; };
#define CUC(call) do { cudaError_t _r = (call); if (_r != cudaSuccess) { fail(e, NGSQ_E_CUDA, "%s: %s", #call, cudaGetErrorString(_r)); return bail(NGSQ_E_CUDA); } } while (0)
  CUC(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, device));
  e->n_sm = prop.multiProcessorCount;
  CUC(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&e->s_aux, cudaStreamNonBlocking));
  for (cudaEvent_t* ev : {&e->ev_start, &e->ev_a, &e->ev_b, &e->ev_c, &e->ev_d, &e->ev_e, &e->ev_f, &e->ev_g}) CUC(cudaEventCreate(ev));
  CUC(cudaMalloc(&e->d_queue, ngsq_engine::kQueueSlots * 4));
  CUC(cudaMalloc(&e->d_flags, sizeof(DevFlags)));
  CUC(cudaMalloc(&e->d_crc_tables, sizeof(CrcTables)));
  // opt-in shared memory sizes are per device: set them for this engine's device (several engines of
  // one process may sit on different GPUs)
  CUC(cudaFuncSetAttribute(crc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrcSmem));
  CUC(cudaFuncSetAttribute(inflate_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecSmem));
  {
    CrcTables t;
    crc_make_tables(t);
    CUC(cudaMemcpy(e->d_crc_tables, &t, sizeof t, cudaMemcpyHostToDevice));
  }
  if (e->cfg.reserve_compressed) {
    uint8_t* p = nullptr;
    size_t cap = e->cfg.reserve_compressed + (e->cfg.reserve_compressed >> 6) + 4096;  // room for 16-byte chunk padding
    CUC(cudaMalloc(&p, cap + 512));
    e->comp_segs.push_back({p, cap, 0});
  }
  if (e->cfg.reserve_inflated) {
    CUC(cudaMalloc(&e->d_out, e->cfg.reserve_inflated + 64));
    e->out_cap = e->cfg.reserve_inflated;
  }
  if (e->cfg.reserve_blocks) {
    CUC(cudaMalloc(&e->d_blocks, (size_t)e->cfg.reserve_blocks * sizeof(BlockDesc)));
    CUC(cudaMalloc(&e->d_status, (size_t)e->cfg.reserve_blocks * 4));
    CUC(cudaMemset(e->d_status, 0, (size_t)e->cfg.reserve_blocks * 4));
    e->blocks_cap = e->cfg.reserve_blocks;
  }
#undef CUC
  *out = e;
  return NGSQ_OK;
}

void ngsq_destroy(ngsq_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (auto ev : e->copy_events) cudaEventDestroy(ev);
  for (auto& p : e->inflate_events) for (cudaEvent_t x : {p.begin, p.decoded_from, p.decoded, p.end, p.crc_begin, p.crc_end}) cudaEventDestroy(x);
  for (cudaEvent_t ev : {e->ev_start, e->ev_a, e->ev_b, e->ev_c, e->ev_d, e->ev_e, e->ev_f, e->ev_g}) if (ev) cudaEventDestroy(ev);
  void* ptrs[] = {e->d_ref_len, e->d_cov_enabled, e->d_diff_base, e->d_cov_slot, e->d_diff, e->d_tile, e->d_res, e->d_out,
                  e->d_blocks, e->d_status, e->d_crcx, e->d_bitmap, e->d_queue, e->d_agree, e->d_out_off, e->d_coff, e->d_base, e->d_rec, e->d_first, e->d_landed,
                  e->d_count, e->d_flags, e->d_crc_tables};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (void* p : e->ed_allocs) cudaFree(p);
  for (void* p : {(void*)e->d_ed_contigs, (void*)e->d_ed_refs, (void*)e->d_ed_alts, (void*)e->d_ed_res}) if (p) cudaFree(p);
  for (void* p : e->ft_allocs) cudaFree(p);
  for (void* p : {(void*)e->d_ft_contigs, (void*)e->d_ft_res}) if (p) cudaFree(p);
  for (auto& sg : e->comp_segs) cudaFree(sg.ptr);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  if (e->s_aux) cudaStreamDestroy(e->s_aux);
  delete e;
}

int ngsq_reset(ngsq_engine* e) {
  if (!e) return NGSQ_E_ARG;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->s_copy));
  CU(cudaStreamSynchronize(e->s_comp));
  CU(cudaStreamSynchronize(e->s_aux));
  for (auto ev : e->copy_events) cudaEventDestroy(ev);
  e->copy_events.clear();
  e->copy_upto.clear();
  e->launched = 0;
  for (auto& p : e->inflate_events) for (cudaEvent_t x : {p.begin, p.decoded_from, p.decoded, p.end, p.crc_begin, p.crc_end}) cudaEventDestroy(x);
  e->inflate_events.clear();
  e->h_blocks.clear(); e->h_coff.clear(); e->h_out_off.clear(); e->h_crc.clear();
  // keep the largest compressed segment, drop the rest
  if (e->comp_segs.size() > 1) {
    size_t best = 0;
    for (size_t i = 1; i < e->comp_segs.size(); ++i) if (e->comp_segs[i].cap > e->comp_segs[best].cap) best = i;
    for (size_t i = 0; i < e->comp_segs.size(); ++i) if (i != best) cudaFree(e->comp_segs[i].ptr);
    ngsq_engine::CompSeg keep = e->comp_segs[best];
    e->comp_segs.assign(1, keep);
  }
  for (auto& sg : e->comp_segs) sg.used = 0;
  e->out_used = 0; e->comp_bytes_total = 0; e->n_launches = 0; e->other_launches = 0;
  if (e->d_status && e->blocks_cap) CU(cudaMemsetAsync(e->d_status, 0, (size_t)e->blocks_cap * 4, e->s_comp));
  e->run_started = false; e->finished = false;
  e->h_res.clear(); e->h_qpos = 0;
  e->h_ed_res.clear();
  e->h_ft_res.clear();
  memset(&e->stats, 0, sizeof e->stats);
  return NGSQ_OK;
}

int ngsq_set_references(ngsq_engine* e, uint32_t n_ref, const uint32_t* ref_len, const uint8_t* coverage_enabled) {
  if (!e || (n_ref && (!ref_len || !coverage_enabled))) return fail(e, NGSQ_E_ARG, "bad references");
  if (e->run_started) return fail(e, NGSQ_E_ARG, "ngsq_set_references must precede the first submit");
  CU(cudaSetDevice(e->device));
  e->n_ref = n_ref;
  e->ref_len.assign(ref_len, ref_len + n_ref);
  e->cov_enabled.assign(coverage_enabled, coverage_enabled + n_ref);
  if (!(e->cfg.flags & NGSQ_F_COVERAGE)) std::fill(e->cov_enabled.begin(), e->cov_enabled.end(), 0);
  e->diff_base.assign(n_ref, 0);
  uint64_t elems = 0;
  uint32_t max_tiles = 1;
  for (uint32_t c = 0; c < n_ref; ++c) {
    e->diff_base[c] = elems;
    if (e->cov_enabled[c]) {
      elems += ((uint64_t)ref_len[c] + 2 + 3) & ~3ull;  // keep every contig 16-byte aligned
      max_tiles = std::max<uint32_t>(max_tiles, (uint32_t)(((uint64_t)ref_len[c] + 1 + kCovTile - 1) / kCovTile));
    }
  }
  for (void* p : {(void*)e->d_ref_len, (void*)e->d_cov_enabled, (void*)e->d_diff_base, (void*)e->d_cov_slot, (void*)e->d_diff, (void*)e->d_tile}) if (p) cudaFree(p);
  e->d_ref_len = nullptr; e->d_cov_enabled = nullptr; e->d_diff_base = nullptr; e->d_cov_slot = nullptr; e->d_diff = nullptr; e->d_tile = nullptr;
  size_t nr = n_ref ? n_ref : 1;
  CU(cudaMalloc(&e->d_ref_len, nr * 4));
  CU(cudaMalloc(&e->d_cov_enabled, nr));
  CU(cudaMalloc(&e->d_diff_base, nr * 8));
  CU(cudaMalloc(&e->d_cov_slot, nr * 4));
  if (n_ref) {
    CU(cudaMemcpy(e->d_ref_len, e->ref_len.data(), n_ref * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_cov_enabled, e->cov_enabled.data(), n_ref, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_diff_base, e->diff_base.data(), n_ref * 8, cudaMemcpyHostToDevice));
  }
  e->diff_elems = elems;
  if (elems) {
    cudaError_t rc = cudaMalloc(&e->d_diff, elems * 4 + 64);
    if (rc != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc difference arrays (%llu bytes): %s", (unsigned long long)elems * 4, cudaGetErrorString(rc));
  }
  CU(cudaMalloc(&e->d_tile, (size_t)max_tiles * 8));
  e->tile_cap = max_tiles;
  e->qpos_cap = 0;
  e->res_words = 0;
  int rc = ensure_res(e, 256, false);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->s_comp));
  if (e->cfg.flags & NGSQ_F_EDITS) {
    // sequences loaded for an earlier header are dropped; per-position counters cover 0..=L of every contig
    for (void* p : e->ed_allocs) cudaFree(p);
    e->ed_allocs.clear();
    for (void* p : {(void*)e->d_ed_contigs, (void*)e->d_ed_refs, (void*)e->d_ed_alts}) if (p) cudaFree(p);
    e->d_ed_contigs = nullptr; e->d_ed_refs = nullptr; e->d_ed_alts = nullptr;
    e->ed_contigs.assign(n_ref, EditsContig{});
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_ref; ++c) {
      e->ed_contigs[c].hdr_len = ref_len[c];
      e->ed_contigs[c].pos_off = total;
      total += (uint64_t)ref_len[c] + 1;
    }
    e->ed_pos_total = total;
    CU(cudaMalloc(&e->d_ed_contigs, nr * sizeof(EditsContig)));
    if (n_ref) CU(cudaMemcpy(e->d_ed_contigs, e->ed_contigs.data(), n_ref * sizeof(EditsContig), cudaMemcpyHostToDevice));
    if (total) {
      cudaError_t r1 = cudaMalloc(&e->d_ed_refs, total * 4), r2 = r1 == cudaSuccess ? cudaMalloc(&e->d_ed_alts, total * 4) : r1;
      if (r2 != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc per-position edit counters (%llu bytes): %s", (unsigned long long)total * 8, cudaGetErrorString(r2));
    }
    if (!e->d_ed_res) CU(cudaMalloc(&e->d_ed_res, E_WORDS * 8));
  }
  if (e->cfg.flags & NGSQ_F_FEATURES) {
    for (void* p : e->ft_allocs) cudaFree(p);
    e->ft_allocs.clear();
    if (e->d_ft_contigs) cudaFree(e->d_ft_contigs);
    e->d_ft_contigs = nullptr;
    e->ft_contigs.assign(n_ref, FeatureContig{});
    e->ft_model_set = false;
    CU(cudaMalloc(&e->d_ft_contigs, nr * sizeof(FeatureContig)));
    if (n_ref) CU(cudaMemcpy(e->d_ft_contigs, e->ft_contigs.data(), n_ref * sizeof(FeatureContig), cudaMemcpyHostToDevice));
    if (!e->d_ft_res) CU(cudaMalloc(&e->d_ft_res, F_WORDS * 8));
  }
  return NGSQ_OK;
}

int ngsq_set_feature_model(ngsq_engine* e, const uint8_t slot_class[5], const uint8_t* primary) {
  if (!e || !slot_class || (e->n_ref && !primary)) return fail(e, NGSQ_E_ARG, "bad feature model");
  if (!(e->cfg.flags & NGSQ_F_FEATURES)) return fail(e, NGSQ_E_ARG, "the engine was created without NGSQ_F_FEATURES");
  if (e->ft_contigs.size() != e->n_ref) return fail(e, NGSQ_E_ARG, "ngsq_set_feature_model must follow ngsq_set_references");
  if (e->run_started) return fail(e, NGSQ_E_ARG, "ngsq_set_feature_model must precede the first submit");
  for (int j = 0; j < 5; ++j)
    if (slot_class[j] > j || slot_class[slot_class[j]] != slot_class[j]) return fail(e, NGSQ_E_ARG, "slot_class[%d] must name the first slot with the same feature name", j);
  CU(cudaSetDevice(e->device));
  memcpy(e->ft_slot_class, slot_class, 5);
  for (uint32_t c = 0; c < e->n_ref; ++c) e->ft_contigs[c].primary = primary[c] ? 1u : 0u;
  if (e->n_ref) CU(cudaMemcpy(e->d_ft_contigs, e->ft_contigs.data(), e->n_ref * sizeof(FeatureContig), cudaMemcpyHostToDevice));
  e->ft_model_set = true;
  return NGSQ_OK;
}

int ngsq_set_features(ngsq_engine* e, uint32_t ref, uint32_t n, const uint32_t* start, const uint32_t* stop, const uint8_t* cls) {
  if (!e || (n && (!start || !stop || !cls))) return fail(e, NGSQ_E_ARG, "bad features");
  if (!(e->cfg.flags & NGSQ_F_FEATURES)) return fail(e, NGSQ_E_ARG, "the engine was created without NGSQ_F_FEATURES");
  if (ref >= e->n_ref || e->ft_contigs.size() != e->n_ref) return fail(e, NGSQ_E_ARG, "ngsq_set_features: reference %u is not in the header given to ngsq_set_references", ref);
  if (e->run_started) return fail(e, NGSQ_E_ARG, "ngsq_set_features must precede the first submit");
  CU(cudaSetDevice(e->device));
  FeatureContig& C = e->ft_contigs[ref];
  for (int k = 0; k < 5; ++k) if (C.n[k]) return fail(e, NGSQ_E_ARG, "reference %u already has features", ref);
  std::vector<uint32_t> a[5], b[5];
  for (uint32_t i = 0; i < n; ++i) {
    if (cls[i] > 4) return fail(e, NGSQ_E_ARG, "feature class %u out of range", cls[i]);
    if (start[i] > stop[i]) return fail(e, NGSQ_E_ARG, "feature %u has start > end (%u > %u): not supported by the counting lookup", i, start[i], stop[i]);
    a[cls[i]].push_back(start[i]);
    b[cls[i]].push_back(stop[i]);
  }
  for (int k = 0; k < 5; ++k) {
    if (a[k].empty()) continue;
    std::sort(a[k].begin(), a[k].end());
    std::sort(b[k].begin(), b[k].end());
    uint32_t *da = nullptr, *db = nullptr;
    const size_t bytes = a[k].size() * 4;
    cudaError_t rc = cudaMalloc(&da, bytes);
    if (rc == cudaSuccess) { e->ft_allocs.push_back(da); rc = cudaMalloc(&db, bytes); }
    if (rc != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc features (%zu bytes): %s", bytes, cudaGetErrorString(rc));
    e->ft_allocs.push_back(db);
    CU(cudaMemcpy(da, a[k].data(), bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, b[k].data(), bytes, cudaMemcpyHostToDevice));
    C.starts[k] = da; C.stops[k] = db; C.n[k] = (uint32_t)a[k].size();
  }
  CU(cudaMemcpy(e->d_ft_contigs + ref, &C, sizeof C, cudaMemcpyHostToDevice));
  return NGSQ_OK;
}

int ngsq_set_reference_bases(ngsq_engine* e, uint32_t ref, const uint8_t* letters, uint64_t n) {
  if (!e || (!letters && n)) return fail(e, NGSQ_E_ARG, "bad reference bases");
  if (!(e->cfg.flags & NGSQ_F_EDITS)) return fail(e, NGSQ_E_ARG, "the engine was created without NGSQ_F_EDITS");
  if (ref >= e->n_ref || e->ed_contigs.size() != e->n_ref) return fail(e, NGSQ_E_ARG, "ngsq_set_reference_bases: reference %u is not in the header given to ngsq_set_references", ref);
  if (e->run_started) return fail(e, NGSQ_E_ARG, "ngsq_set_reference_bases must precede the first submit");
  if (e->ed_contigs[ref].codes) return fail(e, NGSQ_E_ARG, "reference %u already has a sequence", ref);
  CU(cudaSetDevice(e->device));
  const uint64_t words = n / 32 + 1;
  std::vector<uint8_t> codes(n + 1);
  std::vector<uint32_t> bits(words), prefix(words);
  edits_encode(letters, n, codes.data(), bits.data(), prefix.data());
  uint8_t* d_codes = nullptr;
  uint32_t *d_bits = nullptr, *d_prefix = nullptr;
  cudaError_t rc = cudaMalloc(&d_codes, n + 64);
  if (rc == cudaSuccess) { e->ed_allocs.push_back(d_codes); rc = cudaMalloc(&d_bits, words * 4); }
  if (rc == cudaSuccess) { e->ed_allocs.push_back(d_bits); rc = cudaMalloc(&d_prefix, words * 4); }
  if (rc != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc reference sequence (%llu bases): %s", (unsigned long long)n, cudaGetErrorString(rc));
  e->ed_allocs.push_back(d_prefix);
  CU(cudaMemcpy(d_codes, codes.data(), n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_bits, bits.data(), words * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_prefix, prefix.data(), words * 4, cudaMemcpyHostToDevice));
  EditsContig& C = e->ed_contigs[ref];
  C.codes = d_codes; C.bad_bits = d_bits; C.bad_prefix = d_prefix; C.code_len = n;
  CU(cudaMemcpy(e->d_ed_contigs + ref, &C, sizeof C, cudaMemcpyHostToDevice));
  return NGSQ_OK;
}

int ngsq_set_range(ngsq_engine* e, uint64_t first_rec_voffset, uint64_t end_voffset) {
  if (!e) return NGSQ_E_ARG;
  e->first_voff = first_rec_voffset;
  e->end_voff = end_voffset;
  e->range_set = true;
  return NGSQ_OK;
}

int ngsq_bgzf_walk(const uint8_t* p, size_t n, uint64_t file_off, ngsq_block* out, uint32_t cap, uint32_t* n_blocks, size_t* consumed) {
  size_t o = 0;
  uint32_t k = 0;
  while (n - o >= 18) {
    const uint8_t* h = p + o;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return NGSQ_E_BAD_BLOCK;
    uint32_t xlen = h[10] | (h[11] << 8);
    if (n - o < 12 + xlen) break;
    uint32_t bsize = 0;
    bool found = false;
    for (uint32_t q = 0; q + 4 <= xlen;) {
      const uint8_t* s = h + 12 + q;
      uint32_t slen = s[2] | (s[3] << 8);
      if (s[0] == 'B' && s[1] == 'C' && slen == 2 && q + 6 <= xlen) { bsize = s[4] | (s[5] << 8); found = true; }
      q += 4 + slen;
    }
    if (!found) return NGSQ_E_BAD_BLOCK;
    size_t total = (size_t)bsize + 1;
    if (total < 12 + xlen + 8) return NGSQ_E_BAD_BLOCK;
    if (n - o < total) break;
    if (out) {
      if (k >= cap) break;
      ngsq_block& b = out[k];
      b.coffset = file_off + o;
      b.hdr_len = 12 + xlen;
      b.csize = (uint32_t)total;
      memcpy(&b.crc32, h + total - 8, 4);
      memcpy(&b.isize, h + total - 4, 4);
    }
    ++k;
    o += total;
  }
  if (n_blocks) *n_blocks = k;
  if (consumed) *consumed = o;
  return NGSQ_OK;
}

int ngsq_submit(ngsq_engine* e, const uint8_t* bgzf, size_t nbytes, uint64_t file_off) {
  if (!e || !bgzf) return fail(e, NGSQ_E_ARG, "bad submit arguments");
  if (e->finished) return fail(e, NGSQ_E_ARG, "submit after finish; call ngsq_reset");
  CU(cudaSetDevice(e->device));
  uint32_t n = 0;
  size_t used = 0;
  int rc = ngsq_bgzf_walk(bgzf, nbytes, file_off, nullptr, 0, &n, &used);
  if (rc) return fail(e, rc, "malformed BGZF framing in chunk at file offset %llu", (unsigned long long)file_off);
  if (used != nbytes) return fail(e, NGSQ_E_TRUNCATED, "chunk at file offset %llu ends inside a BGZF block (%zu of %zu bytes are whole blocks)", (unsigned long long)file_off, used, nbytes);
  std::vector<ngsq_block> blk(n);
  ngsq_bgzf_walk(bgzf, nbytes, file_off, blk.data(), n, &n, &used);
  rc = start_run(e);
  if (rc) return rc;
  if (e->comp_segs.empty() || e->comp_segs.back().used + nbytes > e->comp_segs.back().cap) {
    // a segment that is full stays where it is (in-flight descriptors point into it); open a new one
    size_t cap = std::max<size_t>(nbytes, e->comp_segs.empty() ? nbytes : (size_t)256 << 20);
    uint8_t* p = nullptr;
    cudaError_t r2 = cudaMalloc(&p, cap + 512);
    if (r2 != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc compressed segment (%zu bytes): %s", cap, cudaGetErrorString(r2));
    e->comp_segs.push_back({p, cap, 0});
  }
  ngsq_engine::CompSeg& seg = e->comp_segs.back();
  uint8_t* dst = seg.ptr + seg.used;
  CU(cudaMemcpyAsync(dst, bgzf, nbytes, cudaMemcpyHostToDevice, e->s_copy));
  cudaEvent_t ev;
  CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU(cudaEventRecord(ev, e->s_copy));
  e->copy_events.push_back(ev);
  seg.used += (nbytes + 15) & ~size_t(15);
  e->comp_bytes_total += nbytes;
  uint32_t first_new, n_new;
  rc = append_blocks(e, blk.data(), n, (uint64_t)(uintptr_t)dst, file_off, &first_new, &n_new);
  if (rc) return rc;
  e->copy_upto.push_back((uint32_t)e->h_blocks.size());
  // inflate in whole waves; the copy of the next chunk overlaps the kernels of this one
  const uint32_t quantum = launch_quantum(e), pending = (uint32_t)e->h_blocks.size() - e->launched;
  if (pending >= quantum) return launch_pending(e, e->launched + pending / quantum * quantum);
  return NGSQ_OK;
}

int ngsq_submit_device(ngsq_engine* e, const void* dev_bgzf, size_t nbytes, const ngsq_block* blocks, uint32_t n_blocks) {
  if (!e || !dev_bgzf || (!blocks && n_blocks)) return fail(e, NGSQ_E_ARG, "bad submit arguments");
  if (e->finished) return fail(e, NGSQ_E_ARG, "submit after finish; call ngsq_reset");
  CU(cudaSetDevice(e->device));
  int rc = start_run(e);
  if (rc) return rc;
  if (!n_blocks) return NGSQ_OK;
  const ngsq_block& last = blocks[n_blocks - 1];
  if (last.coffset + last.csize - blocks[0].coffset > nbytes) return fail(e, NGSQ_E_TRUNCATED, "block table runs past the device buffer");
  e->comp_bytes_total += nbytes;
  uint32_t first_new, n_new;
  rc = append_blocks(e, blocks, n_blocks, (uint64_t)(uintptr_t)dev_bgzf, blocks[0].coffset, &first_new, &n_new);
  if (rc) return rc;
  // resident data: one persistent launch over everything submitted so far (its block queue balances the lanes)
  return launch_pending(e, (uint32_t)e->h_blocks.size());
}

static int voff_to_off(ngsq_engine* e, uint64_t voff, uint64_t* off) {
  uint64_t co = voff >> 16, uo = voff & 0xFFFF;
  auto it = std::lower_bound(e->h_coff.begin(), e->h_coff.end(), co);
  if (it == e->h_coff.end()) {  // at or past the last block: only "end of data" is acceptable
    if (uo == 0) { *off = e->out_used; return NGSQ_OK; }
    return fail(e, NGSQ_E_ARG, "virtual offset %llu is beyond the submitted data", (unsigned long long)voff);
  }
  size_t b = it - e->h_coff.begin();
  if (*it != co) {
    // coffset of an empty (skipped) block: resolves to the start of the next real block
    if (uo != 0) return fail(e, NGSQ_E_ARG, "virtual offset %llu does not address a submitted block", (unsigned long long)voff);
    *off = e->h_out_off[b];
    return NGSQ_OK;
  }
  if (uo > e->h_blocks[b].isize) return fail(e, NGSQ_E_ARG, "virtual offset %llu: uoffset beyond block", (unsigned long long)voff);
  *off = e->h_out_off[b] + uo;
  return NGSQ_OK;
}

int ngsq_finish(ngsq_engine* e) {
  if (!e) return NGSQ_E_ARG;
  if (e->finished) return NGSQ_OK;
  CU(cudaSetDevice(e->device));
  int rc = start_run(e);
  if (rc) return rc;
  rc = launch_pending(e, (uint32_t)e->h_blocks.size());
  if (rc) return rc;
  cudaStream_t s = e->s_comp;
  const uint32_t nb = (uint32_t)e->h_blocks.size();
  const uint64_t d_end = e->out_used;
  DevFlags hf{};
  uint64_t n_rec = 0;
  uint64_t start_off = 0, end_off = d_end;
  CU(cudaEventRecord(e->ev_a, s));
  if (nb) {
    if (!e->range_set) return fail(e, NGSQ_E_ARG, "ngsq_set_range was not called");
    rc = voff_to_off(e, e->first_voff, &start_off);
    if (rc) return rc;
    if (e->end_voff) { rc = voff_to_off(e, e->end_voff, &end_off); if (rc) return rc; }
    if (start_off > end_off) return fail(e, NGSQ_E_ARG, "shard range is empty or inverted");
    // aux tables
    if (nb + 1 > e->aux_cap) {
      for (void* p : {(void*)e->d_out_off, (void*)e->d_coff, (void*)e->d_base, (void*)e->d_first, (void*)e->d_landed, (void*)e->d_count}) if (p) cudaFree(p);
      uint32_t cap = nb + 1 + nb / 4;
      CU(cudaMalloc(&e->d_out_off, (size_t)cap * 8));
      CU(cudaMalloc(&e->d_coff, (size_t)cap * 8));
      CU(cudaMalloc(&e->d_base, (size_t)cap * 8));
      CU(cudaMalloc(&e->d_first, (size_t)cap * 4));
      CU(cudaMalloc(&e->d_landed, (size_t)cap * 4));
      CU(cudaMalloc(&e->d_count, (size_t)cap * 4));
      e->aux_cap = cap;
    }
    e->h_out_off.push_back(d_end);
    CU(cudaMemcpyAsync(e->d_out_off, e->h_out_off.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    e->h_out_off.pop_back();
    CU(cudaMemcpyAsync(e->d_coff, e->h_coff.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(e->d_landed, 0, (size_t)nb * 4, s));
    CU(cudaEventRecord(e->ev_b, s));
    // K3
    size_t first_block = std::upper_bound(e->h_out_off.begin(), e->h_out_off.end(), start_off) - e->h_out_off.begin() - 1;
    find_first_kernel<<<(nb * 32 + 255) / 256, 256, 0, s>>>(e->d_out, e->d_out_off, nb, std::min(d_end, end_off), (int32_t)e->n_ref, start_off, e->d_first);
    walk_kernel<false><<<(nb + 127) / 128, 128, 0, s>>>(e->d_out, e->d_out_off, nb, d_end, end_off, e->d_first, e->d_landed, e->d_count, nullptr, nullptr, &e->d_flags->scan);
    check_landed_kernel<<<(nb + 255) / 256, 256, 0, s>>>(e->d_first, e->d_landed, nb, (uint32_t)first_block, &e->d_flags->scan);
    CU(cudaGetLastError());
    e->other_launches += 3;
    CU(cudaMemcpyAsync(&hf, e->d_flags, sizeof hf, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hf.inflate_err == 0 && hf.scan.chain) {
      // speculative boundaries failed closure: rebuild them serially (correct by construction)
      CU(cudaMemsetAsync(&e->d_flags->scan, 0, sizeof(ScanErr), s));
      CU(cudaMemsetAsync(e->d_landed, 0, (size_t)nb * 4, s));
      chain_serial_kernel<<<1, 32, 0, s>>>(e->d_out, e->d_out_off, nb, d_end, start_off, end_off, e->d_first, &e->d_flags->scan);
      walk_kernel<false><<<(nb + 127) / 128, 128, 0, s>>>(e->d_out, e->d_out_off, nb, d_end, end_off, e->d_first, e->d_landed, e->d_count, nullptr, nullptr, &e->d_flags->scan);
      CU(cudaGetLastError());
      e->other_launches += 2;
      CU(cudaMemcpyAsync(&hf, e->d_flags, sizeof hf, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      if (hf.scan.chain) {
        uint32_t nbad = 0;
        if (hf.inflate_err == 0 && crc_verdict(e, &nbad) == NGSQ_OK && nbad) { e->finished = true; return fail(e, NGSQ_E_CRC, "%u BGZF block(s) failed the CRC32 check", nbad); }
        return fail(e, NGSQ_E_CHAIN, "record chain does not close on the shard's end offset");
      }
    }
  } else {
    CU(cudaEventRecord(e->ev_b, s));
  }
  // inflate status
  if (nb) {
    std::vector<uint32_t> st;
    // cheap summary first: any non-zero status?
    st.resize(nb);
    CU(cudaMemcpyAsync(st.data(), e->d_status, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (uint32_t b = 0; b < nb; ++b)
      if (st[b]) {
        e->finished = true;
        return fail(e, NGSQ_E_BAD_BLOCK, "BGZF block at file offset %llu failed to inflate (%s)", (unsigned long long)e->h_coff[b],
                    st[b] == kBlkIsize ? "ISIZE mismatch" : st[b] == kBlkOverrun ? "output overrun / bad distance" : "invalid DEFLATE stream");
      }
    if (hf.scan.truncated || hf.scan.bad_record) {  // garbage bytes of a block that fails its CRC break the chain too: CRC first
      uint32_t nbad = 0;
      rc = crc_verdict(e, &nbad);
      if (rc) return rc;
      if (nbad) { e->finished = true; return fail(e, NGSQ_E_CRC, "%u BGZF block(s) failed the CRC32 check", nbad); }
    }
    if (hf.scan.truncated) { e->finished = true; return fail(e, NGSQ_E_TRUNCATED, "record chain runs past the end of the submitted data"); }
    if (hf.scan.bad_record) { e->finished = true; return fail(e, NGSQ_E_BAD_RECORD, "malformed record length on the record chain"); }
    scan_counts_kernel<<<1, 1024, 0, s>>>(e->d_count, nb, e->d_base, &e->d_flags->n_records);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&n_rec, &e->d_flags->n_records, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    e->other_launches++;
    if (n_rec > e->rec_cap) {
      if (e->d_rec) cudaFree(e->d_rec);
      e->d_rec = nullptr;
      uint64_t cap = n_rec + n_rec / 8 + 1024;
      cudaError_t r2 = cudaMalloc(&e->d_rec, cap * 8);
      if (r2 != cudaSuccess) return fail(e, NGSQ_E_NOMEM, "cudaMalloc record table (%llu records): %s", (unsigned long long)cap, cudaGetErrorString(r2));
      e->rec_cap = cap;
    }
    if (n_rec) {
      walk_kernel<true><<<(nb + 127) / 128, 128, 0, s>>>(e->d_out, e->d_out_off, nb, d_end, end_off, e->d_first, nullptr, nullptr, e->d_base, e->d_rec, &e->d_flags->scan);
      e->other_launches++;
    }
  }
  CU(cudaEventRecord(e->ev_c, s));
  // K4-K8
  uint32_t max_lseq = hf.scan.max_lseq;
  rc = ensure_res(e, max_lseq, true);
  if (rc) return rc;
  if (n_rec) {
    FacetParams P{};
    P.d = e->d_out; P.rec = e->d_rec; P.n_rec = n_rec; P.out_off = e->d_out_off; P.coff = e->d_coff; P.d_end = d_end;
    P.max_records = e->cfg.max_records; P.gc_seed = e->cfg.gc_seed; P.n_ref = (int32_t)e->n_ref; P.flags = e->cfg.flags;
    P.ref_len = e->d_ref_len; P.cov_enabled = e->d_cov_enabled; P.diff_base = e->d_diff_base; P.diff = e->d_diff; P.cov_slot = e->d_cov_slot;
    P.res = e->d_res; P.qual = e->d_res + e->qual_off;
    P.qpos_smem = std::min<uint32_t>(std::max<uint32_t>(max_lseq, 1), 256);
    P.qpos_cap = e->qpos_cap;
    size_t smem = (size_t)qual_table_bytes(P.qpos_smem) * (kFacetThreads / 32) + (size_t)(kTlenPad + kGcPad + kCigWords) * 4;
    CU(cudaFuncSetAttribute(facets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, facets_kernel, kFacetThreads, smem));
    if (occ < 1) occ = 1;
    uint64_t want = (n_rec + kFacetThreads - 1) / kFacetThreads;  // a warp takes 32 records per step
    uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)e->n_sm * occ);
    facets_kernel<<<grid, kFacetThreads, smem, s>>>(P);
    CU(cudaGetLastError());
    e->other_launches++;
  }
  CU(cudaEventRecord(e->ev_d, s));
  // K9
  if (e->cfg.flags & NGSQ_F_COVERAGE) {
    for (uint32_t c = 0; c < e->n_ref; ++c) {
      if (!e->cov_enabled[c]) continue;
      uint32_t n = e->ref_len[c] + 1;
      uint32_t n_tiles = (n + kCovTile - 1) / kCovTile;
      const int32_t* df = e->d_diff + e->diff_base[c];
      uint64_t* slot = e->d_res + e->cov_slot[c];
      cov_tile_sums_kernel<<<n_tiles, kCovThreads, 0, s>>>(df, n, slot + COV_TOUCHED, e->d_tile);
      cov_scan_tiles_kernel<<<1, 1024, 0, s>>>(e->d_tile, n_tiles, slot + COV_TOUCHED);
      cov_resolve_kernel<<<std::min<uint32_t>(n_tiles, e->n_sm * 4), kCovThreads, 0, s>>>(df, n, e->d_tile, n_tiles, slot);
      e->other_launches += 3;
    }
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(e->ev_e, s));
  // K11 (NGSQ_F_EDITS): per-record edit counts and per-position ref / alt counters, then the VAF histogram
  const bool do_edits = (e->cfg.flags & NGSQ_F_EDITS) && e->d_ed_res && e->ed_contigs.size() == e->n_ref;
  if (do_edits && n_rec) {
    EditsParams EP{};
    EP.d = e->d_out; EP.rec = e->d_rec; EP.n_rec = n_rec; EP.out_off = e->d_out_off; EP.n_ref = (int32_t)e->n_ref;
    EP.contigs = e->d_ed_contigs; EP.refs = e->d_ed_refs; EP.alts = e->d_ed_alts; EP.res = e->d_ed_res;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n_rec + 255) / 256, (uint64_t)e->n_sm * 8);
    edits_kernel<<<grid, 256, 0, s>>>(EP);
    if (e->ed_pos_total) {
      const uint32_t vgrid = (uint32_t)std::min<uint64_t>((e->ed_pos_total + 255) / 256, (uint64_t)e->n_sm * 8);
      edits_vaf_kernel<<<vgrid, 256, 0, s>>>(e->d_ed_refs, e->d_ed_alts, e->ed_pos_total, e->d_ed_res);
    }
    CU(cudaGetLastError());
    e->other_launches += 2;
  }
  if (do_edits) {
    e->h_ed_res.resize(E_WORDS);
    CU(cudaMemcpyAsync(e->h_ed_res.data(), e->d_ed_res, E_WORDS * 8, cudaMemcpyDeviceToHost, s));
  }
  // K12 (NGSQ_F_FEATURES): per-record overlap counts against the gene model
  const bool do_features = (e->cfg.flags & NGSQ_F_FEATURES) && e->d_ft_res && e->ft_model_set && e->ft_contigs.size() == e->n_ref;
  if ((e->cfg.flags & NGSQ_F_FEATURES) && !do_features) return fail(e, NGSQ_E_ARG, "NGSQ_F_FEATURES needs ngsq_set_feature_model before the first submit");
  if (do_features && n_rec) {
    FeatureParams FP{};
    FP.d = e->d_out; FP.rec = e->d_rec; FP.n_rec = n_rec; FP.max_records = e->cfg.max_records; FP.out_off = e->d_out_off; FP.n_ref = (int32_t)e->n_ref;
    FP.contigs = e->d_ft_contigs; FP.res = e->d_ft_res;
    memcpy(FP.slot_class, e->ft_slot_class, 8);
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n_rec + 255) / 256, (uint64_t)e->n_sm * 8);
    features_kernel<<<grid, 256, 0, s>>>(FP);
    CU(cudaGetLastError());
    e->other_launches += 1;
  }
  if (do_features) {
    e->h_ft_res.resize(F_WORDS);
    CU(cudaMemcpyAsync(e->h_ft_res.data(), e->d_ft_res, F_WORDS * 8, cudaMemcpyDeviceToHost, s));
  }
  CU(cudaEventRecord(e->ev_g, s));
  // the step ends when the CRC stream is done too (its last event closes the device-timed region)
  if (!e->inflate_events.empty()) CU(cudaStreamWaitEvent(s, e->inflate_events.back().crc_end, 0));
  e->h_res.resize(e->res_words);
  CU(cudaMemcpyAsync(e->h_res.data(), e->d_res, e->res_words * 8, cudaMemcpyDeviceToHost, s));
  uint32_t crc_bad = 0;
  CU(cudaMemcpyAsync(&crc_bad, &e->d_flags->crc_bad, 4, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(e->ev_f, s));
  CU(cudaStreamSynchronize(s));
  e->finished = true;
  if (crc_bad) return fail(e, NGSQ_E_CRC, "%u BGZF block(s) failed the CRC32 check", crc_bad);
  e->h_qpos = (uint32_t)e->h_res[R_QUAL_POSITIONS];
  // stats
  ngsq_stats& st = e->stats;
  st.records = n_rec; st.blocks = nb; st.compressed_bytes = e->comp_bytes_total; st.inflated_bytes = d_end; st.max_read_len = max_lseq;
  st.ms_inflate = 0;
  st.ms_inflate_decode = 0;
  st.ms_inflate_resolve = 0;
  float crc_ms = 0;
  for (auto& p : e->inflate_events) {
    float ms = 0;
    cudaEventElapsedTime(&ms, p.begin, p.end); st.ms_inflate += ms;
    cudaEventElapsedTime(&ms, p.decoded_from, p.decoded); st.ms_inflate_decode += ms;
    cudaEventElapsedTime(&ms, p.decoded, p.end); st.ms_inflate_resolve += ms;
    cudaEventElapsedTime(&ms, p.crc_begin, p.crc_end); crc_ms += ms;
  }
  st.ms_crc = crc_ms;
  cudaEventElapsedTime(&st.ms_scan, e->ev_b, e->ev_c);
  cudaEventElapsedTime(&st.ms_facets, e->ev_c, e->ev_d);
  cudaEventElapsedTime(&st.ms_coverage, e->ev_d, e->ev_e);
  cudaEventElapsedTime(&st.ms_edits, e->ev_e, e->ev_g);
  cudaEventElapsedTime(&st.ms_total, e->ev_start, e->ev_f);
  st.inflate_launches = e->n_launches;
  st.other_launches = e->other_launches;
  if (e->h_res[R_ERR_QUAL]) return fail(e, NGSQ_E_QUAL_RANGE, "a record holds a quality score above 93 (the reference's decoder rejects it)");
  if (e->h_res[R_ERR_RECORD]) return fail(e, NGSQ_E_BAD_RECORD, "malformed BAM record (field overrun, CIGAR op > 8, reference id out of range, or a mapped pair without reference ids)");
  if (do_edits && e->h_ed_res[E_ERR]) {
    static const char* const kinds[] = {"", "", "Could not parse read name", "sequence not found in reference FASTA", "record reaches past the end of the reference sequence",
                                        "invalid base in the reference sequence", "step-through: no reference base left", "step-through: no record base left",
                                        "step-through: reference sequence was not fully consumed", "step-through: record sequence was not fully consumed",
                                        "more than 512 edits in one read", "matched position beyond the sequence length of the header", "invalid CIGAR operation"};
    const uint64_t k = e->h_ed_res[E_ERR];
    return fail(e, NGSQ_E_EDITS, "Edits: %s (the reference aborts the run here)", k < sizeof kinds / sizeof *kinds ? kinds[k] : "unknown failure");
  }
  if (do_features && e->h_ft_res[F_ERR]) {
    static const char* const kinds[] = {"", "Could not parse read name", "Could not parse reference sequence id for read", "Could not parse record's start position.",
                                        "invalid CIGAR operation"};
    const uint64_t k = e->h_ft_res[F_ERR];
    return fail(e, NGSQ_E_FEATURES, "Genomic Features: %s (the reference aborts the run here)", k < sizeof kinds / sizeof *kinds ? kinds[k] : "unknown failure");
  }
  return NGSQ_OK;
}

int ngsq_get_features(ngsq_engine* e, uint64_t counts[9]) {
  if (!e || !counts) return NGSQ_E_ARG;
  if (!e->finished || e->h_ft_res.size() != F_WORDS) return fail(e, NGSQ_E_ARG, "Genomic Features results requested before ngsq_finish or without NGSQ_F_FEATURES");
  if (e->h_ft_res[F_ERR]) return fail(e, NGSQ_E_FEATURES, "the Genomic Features facet failed; no results");
  memcpy(counts, e->h_ft_res.data(), 9 * 8);
  return NGSQ_OK;
}

int ngsq_get_edits(ngsq_engine* e, uint64_t read_one[513], uint64_t read_two[513], uint64_t vaf[101], uint64_t* records) {
  if (!e) return NGSQ_E_ARG;
  if (!e->finished || e->h_ed_res.size() != E_WORDS) return fail(e, NGSQ_E_ARG, "Edits results requested before ngsq_finish or without NGSQ_F_EDITS");
  if (e->h_ed_res[E_ERR]) return fail(e, NGSQ_E_EDITS, "the Edits facet failed; no results");
  if (read_one) memcpy(read_one, &e->h_ed_res[E_READ_ONE], 513 * 8);
  if (read_two) memcpy(read_two, &e->h_ed_res[E_READ_TWO], 513 * 8);
  if (vaf) memcpy(vaf, &e->h_ed_res[E_VAF], 101 * 8);
  if (records) *records = e->h_ed_res[E_RECORDS];
  return NGSQ_OK;
}

#define NEED_RESULTS()                                                                   \
  if (!e) return NGSQ_E_ARG;                                                             \
  if (!e->finished || e->h_res.empty()) return fail(e, NGSQ_E_ARG, "results requested before ngsq_finish")

int ngsq_get_general(ngsq_engine* e, uint64_t out[34]) {
  NEED_RESULTS();
  memcpy(out, &e->h_res[R_GENERAL], 34 * 8);
  return NGSQ_OK;
}

int ngsq_get_tlen(ngsq_engine* e, uint64_t hist[1025], uint64_t* processed, uint64_t* ignored) {
  NEED_RESULTS();
  memcpy(hist, &e->h_res[R_TLEN_HIST], 1025 * 8);
  *processed = e->h_res[R_TLEN_PROCESSED];
  *ignored = e->h_res[R_TLEN_IGNORED];
  return NGSQ_OK;
}

int ngsq_get_gc(ngsq_engine* e, uint64_t hist[101], uint64_t nuc[3], uint64_t rec[3]) {
  NEED_RESULTS();
  memcpy(hist, &e->h_res[R_GC_HIST], 101 * 8);
  memcpy(nuc, &e->h_res[R_GC_NUC], 24);
  memcpy(rec, &e->h_res[R_GC_REC], 24);
  return NGSQ_OK;
}

int ngsq_get_quality(ngsq_engine* e, uint64_t* out, size_t cap_positions, uint32_t* n_positions) {
  NEED_RESULTS();
  uint32_t n = e->h_qpos;
  if (n_positions) *n_positions = n;
  if (!out) return NGSQ_OK;
  if (cap_positions < n) return fail(e, NGSQ_E_ARG, "quality buffer holds %zu positions, %u needed", cap_positions, n);
  memcpy(out, &e->h_res[e->qual_off], (size_t)n * 94 * 8);
  return NGSQ_OK;
}

int ngsq_get_coverage_contig(ngsq_engine* e, uint32_t ref, ngsq_cov_ints* out, uint64_t* bin_sums, size_t cap) {
  NEED_RESULTS();
  if (ref >= e->n_ref || !out) return fail(e, NGSQ_E_ARG, "bad reference index");
  memset(out, 0, sizeof *out);
  if (!e->cov_enabled[ref]) return NGSQ_OK;
  const uint64_t* slot = &e->h_res[e->cov_slot[ref]];
  out->touched = slot[COV_TOUCHED] ? 1 : 0;
  if (!out->touched) return NGSQ_OK;
  out->n_bins = e->cov_nbins[ref];
  out->pileup_too_large = slot[COV_TOO_LARGE];
  memcpy(out->hist, slot + COV_HIST, 2049 * 8);
  if (bin_sums) {
    if (cap < out->n_bins) return fail(e, NGSQ_E_ARG, "bin buffer holds %zu entries, %u needed", cap, out->n_bins);
    memcpy(bin_sums, slot + COV_BINS, (size_t)out->n_bins * 8);
  }
  return NGSQ_OK;
}

int ngsq_get_coverage_global(ngsq_engine* e, uint64_t* nonsensical_records) {
  NEED_RESULTS();
  *nonsensical_records = e->h_res[R_NONSENSICAL];
  return NGSQ_OK;
}

int ngsq_get_stats(ngsq_engine* e, ngsq_stats* out) {
  if (!e || !out) return NGSQ_E_ARG;
  *out = e->stats;
  return NGSQ_OK;
}

int ngsq_nccl_unique_id(char out[128]) {
  std::string err;
  if (!load_nccl(err)) { g_create_err = err; return NGSQ_E_NCCL; }
  NcclId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc) { g_create_err = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return NGSQ_E_NCCL; }
  memcpy(out, id.b, 128);
  return NGSQ_OK;
}

int ngsq_comm_init(ngsq_engine* e, int n_ranks, int rank, const char id[128]) {
  if (!e || !id) return NGSQ_E_ARG;
  std::string err;
  if (!load_nccl(err)) return fail(e, NGSQ_E_NCCL, "%s", err.c_str());
  CU(cudaSetDevice(e->device));
  // ncclCommInitRank(ncclComm_t*, int, ncclUniqueId /*by value*/, int)
  typedef int (*init_fn)(void**, int, NcclId, int);
  NcclId nid;
  memcpy(nid.b, id, 128);
  int rc = ((init_fn)g_nccl.CommInitRank)(&e->comm, n_ranks, nid, rank);
  if (rc) return fail(e, NGSQ_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  e->n_ranks = n_ranks;
  e->rank = rank;
  return NGSQ_OK;
}

int ngsq_set_quality_positions(ngsq_engine* e, uint32_t n_positions) {
  if (!e) return NGSQ_E_ARG;
  CU(cudaSetDevice(e->device));
  int rc = ensure_res(e, n_positions, true);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->s_comp));
  return NGSQ_OK;
}

int ngsq_result_buffer(ngsq_engine* e, void** dev_ptr, size_t* n_words) {
  if (!e || !dev_ptr || !n_words) return NGSQ_E_ARG;
  *dev_ptr = e->d_res;
  *n_words = e->res_words;
  return NGSQ_OK;
}

// After an external reduction of the device buffer, refresh the host copy.
int ngsq_refresh_results(ngsq_engine* e) {
  if (!e || !e->finished) return fail(e, NGSQ_E_ARG, "refresh before finish");
  CU(cudaSetDevice(e->device));
  e->h_res.resize(e->res_words);
  CU(cudaMemcpyAsync(e->h_res.data(), e->d_res, e->res_words * 8, cudaMemcpyDeviceToHost, e->s_comp));
  CU(cudaStreamSynchronize(e->s_comp));
  return NGSQ_OK;
}

int ngsq_reduce(ngsq_engine* e, int root) {
  if (!e || !e->finished) return fail(e, NGSQ_E_ARG, "reduce before finish");
  if (!e->comm) return fail(e, NGSQ_E_NCCL, "ngsq_comm_init was not called");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = e->s_comp;
  CU(cudaEventRecord(e->ev_a, s));
  // agree on the quality table size (max over ranks of the longest read with qualities) and check
  // that every rank packs the same layout: a count mismatch would hang the reduce
  if (!e->d_agree) CU(cudaMalloc(&e->d_agree, 3 * 8));
  const int ncclUint64 = 5, ncclSum = 0, ncclMax = 2;
  const uint64_t layout = e->qual_off;
  uint64_t agree[3] = {e->h_res[R_QUAL_POSITIONS], layout, ~layout};
  CU(cudaMemcpyAsync(e->d_agree, agree, sizeof agree, cudaMemcpyHostToDevice, s));
  int rc = g_nccl.AllReduce(e->d_agree, e->d_agree, 3, ncclUint64, ncclMax, e->comm, s);
  if (rc) return fail(e, NGSQ_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  CU(cudaMemcpyAsync(agree, e->d_agree, sizeof agree, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (agree[1] != layout || ~agree[2] != layout)
    return fail(e, NGSQ_E_ARG, "result layouts differ across ranks: ngsq_set_references must get the same lengths and coverage mask on every rank");
  const uint64_t qmax = agree[0];
  rc = ensure_res(e, (uint32_t)qmax, true);
  if (rc) return rc;
  // the max word must not be summed: park it, reduce, restore
  CU(cudaMemsetAsync(e->d_res + R_QUAL_POSITIONS, 0, 8, s));
  // the same count on every rank (a rank may hold a larger table from an earlier run)
  const size_t n_words = (size_t)e->qual_off + (size_t)std::max<uint64_t>(qmax, 256) * 94;
  rc = g_nccl.Reduce(e->d_res, e->d_res, n_words, ncclUint64, ncclSum, root, e->comm, s);
  if (rc) return fail(e, NGSQ_E_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  CU(cudaMemcpyAsync(e->d_res + R_QUAL_POSITIONS, &qmax, 8, cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));
  e->other_launches += 2;
  if ((e->cfg.flags & NGSQ_F_EDITS) && e->d_ed_res && e->h_ed_res.size() == E_WORDS) {
    // additive like everything else: every shard owns whole contigs, so per-position counters never meet
    rc = g_nccl.Reduce(e->d_ed_res, e->d_ed_res, E_WORDS, ncclUint64, ncclSum, root, e->comm, s);
    if (rc) return fail(e, NGSQ_E_NCCL, "ncclReduce (edits): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    CU(cudaMemcpyAsync(e->h_ed_res.data(), e->d_ed_res, E_WORDS * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    e->other_launches += 1;
  }
  if ((e->cfg.flags & NGSQ_F_FEATURES) && e->d_ft_res && e->h_ft_res.size() == F_WORDS) {
    rc = g_nccl.Reduce(e->d_ft_res, e->d_ft_res, F_WORDS, ncclUint64, ncclSum, root, e->comm, s);
    if (rc) return fail(e, NGSQ_E_NCCL, "ncclReduce (features): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    CU(cudaMemcpyAsync(e->h_ft_res.data(), e->d_ft_res, F_WORDS * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    e->other_launches += 1;
  }
  rc = ngsq_refresh_results(e);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_b, s));
  CU(cudaEventSynchronize(e->ev_b));
  cudaEventElapsedTime(&e->stats.ms_reduce, e->ev_a, e->ev_b);
  e->h_qpos = (uint32_t)qmax;
  // touched flags were summed: any non-zero means touched (getters test != 0)
  return NGSQ_OK;
}

void* ngsq_host_alloc(size_t nbytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void ngsq_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ngsq_inflate_to_host(ngsq_engine* e, const uint8_t* bgzf, size_t nbytes, uint8_t* out, size_t cap, size_t* n_out) {
  if (!e || !bgzf || !out) return fail(e, NGSQ_E_ARG, "bad arguments");
  CU(cudaSetDevice(e->device));
  uint32_t n = 0;
  size_t used = 0;
  int rc = ngsq_bgzf_walk(bgzf, nbytes, 0, nullptr, 0, &n, &used);
  if (rc) return fail(e, rc, "malformed BGZF framing");
  std::vector<ngsq_block> blk(n);
  ngsq_bgzf_walk(bgzf, nbytes, 0, blk.data(), n, &n, &used);
  uint8_t* d_in = nullptr;
  uint8_t* d_o = nullptr;
  BlockDesc* d_b = nullptr;
  uint32_t *d_st = nullptr, *d_q = nullptr;
  std::vector<BlockDesc> hb;
  uint64_t total = 0;
  CU(cudaMalloc(&d_in, used + 512));
  for (auto& b : blk) {
    if (b.csize < b.hdr_len + 8 || b.isize > 65536) { cudaFree(d_in); return fail(e, NGSQ_E_BAD_BLOCK, "malformed BGZF block"); }
    if (!b.isize) continue;
    hb.push_back({(uint64_t)(uintptr_t)d_in + b.coffset + b.hdr_len, total, b.csize - b.hdr_len - 8, b.isize});
    total += b.isize;
  }
  if (n_out) *n_out = (size_t)total;
  if (total > cap) { cudaFree(d_in); return fail(e, NGSQ_E_ARG, "output buffer too small (%llu needed)", (unsigned long long)total); }
  int ret = NGSQ_OK;
  if (!hb.empty()) {
    cudaStream_t s = e->s_copy;
    CU(cudaMalloc(&d_o, total + 64));
    CU(cudaMalloc(&d_b, hb.size() * sizeof(BlockDesc)));
    CU(cudaMalloc(&d_st, hb.size() * 4 + 4));
    d_q = d_st + hb.size();
    CU(cudaMemsetAsync(d_st, 0, hb.size() * 4 + 4, s));
    CU(cudaMemcpyAsync(d_in, bgzf, used, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_b, hb.data(), hb.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice, s));
    uint32_t* d_bm = nullptr;
    CU(cudaMalloc(&d_bm, hb.size() * kBitmapWords * 4));
    ret = launch_inflate(e, d_b, (uint32_t)hb.size(), d_o, d_q, d_st, d_bm, s);
    std::vector<uint32_t> st(hb.size());
    if (!ret) {
      CU(cudaMemcpyAsync(out, d_o, total, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(st.data(), d_st, hb.size() * 4, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      for (size_t i = 0; i < st.size(); ++i)
        if (st[i]) { ret = fail(e, NGSQ_E_BAD_BLOCK, "block %zu failed to inflate (status %u)", i, st[i]); break; }
    }
    cudaStreamSynchronize(s);
    cudaFree(d_o); cudaFree(d_b); cudaFree(d_st);
    if (d_bm) cudaFree(d_bm);
  }
  cudaFree(d_in);
  return ret;
}

}  // extern "C"
