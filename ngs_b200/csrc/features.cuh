// K12 — the Genomic Features facet (reference: src/qc/record_based/features.rs:115-242 over the
// rust-lapper interval lookups built by try_from, 270-355): per record, which kinds of gene-model features its
// alignment interval overlaps — 5' UTR / 3' UTR / CDS tallies and exonic / intronic / intergenic.
//
// Written against the oracle (oracle/ngsqc_oracle.c, "Genomic Features facet"); the per-record logic is shared with a
// host model (tools/features_model.cpp, tests/test_features_model.py); GPU parity: tests/test_gpu_features.py.  Launched
// per wave when NGSQ_F_FEATURES is set.
//
// No interval tree on the device.  The facet never needs WHICH features overlap a read, only HOW MANY of each name:
//   * the 5'/3'/CDS chain of features.rs:185-209 hands the k-th overlapping feature of a name to the k-th of the
//     three slots that carries that name (names may coincide: GTF models call both UTRs "UTR"), so slot j is counted
//     iff the overlap count of its name exceeds the number of earlier slots with the same name;
//   * gene regions (212-236) need "any gene" and "any exon".
// For features with start <= stop the number that overlap the half-open query [a, b) under rust-lapper's rule
// (start < b && stop > a) is  #(start < b) - #(stop <= a): two binary searches over the name's starts and stops,
// each sorted on its own, per contig.  The caller files every feature under a CLASS: the smallest slot index
// (0 five_prime_utr, 1 three_prime_utr, 2 coding_sequence, 3 exon, 4 gene) whose configured name equals the feature's
// type; classes 0-2 live in the "exonic translation" set, 3-4 in the "gene regions" set, exactly as try_from files them.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include "recscan.cuh"
#endif

#ifndef NGSQ_HD
#if defined(__CUDACC__)
#define NGSQ_HD __host__ __device__ __forceinline__
#else
#define NGSQ_HD inline
#endif
#endif

namespace ngsq {

// result block (u64 words, additive across shards); field order of features/metrics.rs
constexpr uint32_t F_UTR5 = 0, F_UTR3 = 1, F_CDS = 2, F_INTERGENIC = 3, F_EXONIC = 4, F_INTRONIC = 5, F_PROCESSED = 6,
                   F_IGNORED_FLAGS = 7, F_IGNORED_NONPRIMARY = 8, F_ERR = 9, F_WORDS = 16;

enum : uint32_t {
  kFtNoName = 1,      // "Could not parse read name" (features.rs:117-120; checked before anything else)
  kFtNoReference = 2, // mapped flag without a reference id (features.rs:131-139)
  kFtNoStart = 3,     // "Could not parse record's start position." (features.rs:167-170)
  kFtBadCigar = 4,
};

struct FeatureContig {        // one per reference sequence of the header
  const uint32_t* starts[5];  // per class: feature starts, ascending (nullptr: none)
  const uint32_t* stops[5];   // per class: feature stops (the GFF end, used as an exclusive bound), ascending
  uint32_t n[5];
  uint32_t primary;           // the sequence is in the genome's primary assembly (features.rs:157-164)
};

NGSQ_HD uint32_t ft_ld32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// number of values < x in a[0, n), a ascending
NGSQ_HD uint32_t ft_count_below(const uint32_t* a, uint32_t n, uint64_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if ((uint64_t)a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

NGSQ_HD uint32_t ft_overlaps(const FeatureContig& C, uint32_t cls, uint64_t a, uint64_t b) {  // features of a class over [a, b)
  if (!C.n[cls]) return 0;
  return ft_count_below(C.starts[cls], C.n[cls], b) - ft_count_below(C.stops[cls], C.n[cls], a + 1);
}

// One record.  Returns 0 and a bit mask of the F_* counters to increment, or a kFt* error.
// slot_class[j] = smallest slot index whose configured name equals slot j's name.
NGSQ_HD uint32_t features_record(const uint8_t* rec, int32_t n_ref, const FeatureContig* contigs, const uint8_t* slot_class, uint32_t* bits) {
  const int32_t ref = (int32_t)ft_ld32(rec + 4), pos = (int32_t)ft_ld32(rec + 8);
  const uint32_t w3 = ft_ld32(rec + 12), w4 = ft_ld32(rec + 16), lseq = ft_ld32(rec + 20);
  const uint32_t lname = w3 & 255, ncig = w4 & 0xFFFF, flag = w4 >> 16;
  *bits = 0;
  if (32ull + lname + 4ull * ncig + (lseq + 1ull) / 2 + lseq > ft_ld32(rec)) return 0;  // malformed: the facet kernel fails the run
  if (lname <= 1 || (lname == 2 && rec[36] == '*')) return kFtNoName;
  if (flag & 0x4u) { *bits = 1u << F_IGNORED_FLAGS; return 0; }
  if (ref < 0 || ref >= n_ref) return kFtNoReference;
  const FeatureContig& C = contigs[ref];
  if (!C.primary) { *bits = 1u << F_IGNORED_NONPRIMARY; return 0; }
  if (pos < 0) return kFtNoStart;
  const uint8_t* cig = rec + 36 + lname;
  uint64_t span = 0;
  for (uint32_t i = 0; i < ncig; ++i) {
    const uint32_t op = ft_ld32(cig + 4 * i), k = op & 15;
    if (k > 8) return kFtBadCigar;
    if ((0x18Du >> k) & 1u) span += op >> 4;  // M D N = X (utils/cigar.rs:6-11)
  }
  const uint64_t a = (uint64_t)pos + 1, b = a + span + 1;  // utrs.find(start, end + 1), end = start + span (features.rs:172-186)
  uint32_t out = 1u << F_PROCESSED;
  // exonic translation regions: slot j counts iff its name's overlap count exceeds the earlier slots of that name
  uint32_t cnt[3] = {0, 0, 0};
  for (uint32_t j = 0; j < 3; ++j)
    if (slot_class[j] == j) cnt[j] = ft_overlaps(C, j, a, b);
  uint32_t earlier[3] = {0, 0, 0};
  for (uint32_t j = 0; j < 3; ++j) {
    const uint32_t c = slot_class[j];
    if (cnt[c] > earlier[c]) out |= 1u << (F_UTR5 + j);
    earlier[c]++;
  }
  // gene regions: classes 3 (exon) and 4 (gene); a name shared with an earlier slot lives in the other set (never found here),
  // gene == exon name: the gene test wins (features.rs:215-219)
  const uint32_t gc = slot_class[4], ec = slot_class[3];
  const bool has_gene = gc >= 3 && ft_overlaps(C, gc, a, b) != 0;
  const bool has_exon = ec == 3 && gc != 3 && ft_overlaps(C, 3, a, b) != 0;
  out |= 1u << (has_gene ? (has_exon ? F_EXONIC : F_INTRONIC) : F_INTERGENIC);
  *bits = out;
  return 0;
}

#if defined(__CUDACC__)

struct FeatureParams {
  const uint8_t* d;      // base of the wave's slot
  const uint64_t* rec;   // record table of the wave (recscan.cuh): slot offset in the low 40 bits
  const RunState* st;    // wave_rec, rec_base, fatal
  uint64_t max_records;  // `-n`: first N records in file order; 0 = all
  int32_t n_ref;
  const FeatureContig* contigs;
  uint8_t slot_class[8];
  unsigned long long* res;  // F_* words
};

__global__ void __launch_bounds__(256) features_kernel(FeatureParams P) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t acc = 0, err = 0;  // lane k owns counter k
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n_rec = P.st->fatal ? 0 : P.st->wave_rec, rec_base = P.st->rec_base;
  const uint64_t n_iter = (n_rec + stride - 1) / stride;
  for (uint64_t it = 0; it < n_iter; ++it) {  // whole warps stay in the loop: the tallies below are ballots
    const uint64_t r = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bits = 0;
    if (r < n_rec && (P.max_records == 0 || rec_base + r < P.max_records)) {
      const uint32_t st = features_record(P.d + (P.rec[r] & kRecOffMask), P.n_ref, P.contigs, P.slot_class, &bits);
      if (st) { err = err > st ? err : st; bits = 0; }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, (bits >> k) & 1u);
      if ((int)lane == k) acc += __popc(m);
    }
  }
  if (lane < 9 && acc) atomicAdd(&P.res[lane], (unsigned long long)acc);
  if (err) atomicMax(&P.res[F_ERR], (unsigned long long)err);
}

#endif  // __CUDACC__

}  // namespace ngsq
