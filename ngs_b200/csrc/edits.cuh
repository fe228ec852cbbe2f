// K11 — the Edits facet (reference: src/qc/sequence_based/edits.rs:217-344 over
// src/utils/alignment.rs:48-107 and src/utils/cigar.rs:6-23): per record, the number of M-positions whose read
// base differs from the reference FASTA base, per reference position the counts of matching / differing reads,
// and at the end of the run the histogram of per-position variant allele fractions.
//
// Written against the oracle (oracle/ngsqc_oracle.c, "Edits"); the per-record logic is shared with a host model
// (tools/edits_model.cpp, tests/test_edits_model.py); GPU parity: tests/test_gpu_edits.py.  Launched per wave when
// NGSQ_F_EDITS is set.
//
// Layout in HBM (per contig with a loaded sequence):
//   codes      (one allocation per contig) 1 byte per FASTA base: the BAM 4-bit code of the letter (index in "=ACMGRSVTWYHKDBN"), 0xFF for a
//              letter noodles' Base::try_from rejects
//   bad_bits   1 bit per base (letter rejected) + bad_prefix, one u32 per 32 bases (rejected letters before the word):
//              "is any letter of [a, b) rejected" in O(1) — the reference converts the whole slice under a record
//              before it walks the CIGAR (edits.rs:259-265), deletions and skipped regions included
//   refs/alts  u32 per reference position 0..L (1-based positions, like Histogram::zero_based_with_capacity(L))
// One record per lane: records are short and the per-base work is a compare and one global reduction.
#pragma once
#include <stdint.h>
#include <string.h>

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

// result block (u64 words, additive across shards)
constexpr uint32_t E_READ_ONE = 0;     // 513: Histogram::default() = 0..=512 (histogram.rs:472-481)
constexpr uint32_t E_READ_TWO = 513;   // 513
constexpr uint32_t E_VAF = 1026;       // 101
constexpr uint32_t E_RECORDS = 1127;   // records that reached the step-through
constexpr uint32_t E_ERR = 1128;       // first error kind seen (max of codes; 0 = none)
constexpr uint32_t E_WORDS = 1136;

// outcome of one record
enum : uint32_t {
  kEdCounted = 0,       // edits counted
  kEdSkipped = 1,       // not part of the facet's input (filter / flags)
  // errors: the reference aborts the run on each of these
  kEdNoName = 2,        // read name "*" (edits.rs:233-236)
  kEdNoSequence = 3,    // no FASTA sequence loaded for the record's contig (edits.rs:196-198)
  kEdSlice = 4,         // record reaches past the end of the FASTA sequence (edits.rs:259-262, unwrap of None)
  kEdBadBase = 5,       // a letter of the slice is not a Base (edits.rs:263-265)
  kEdRefShort = 6,      // alignment.rs:60-70 "...consume a reference base, but no such base was found"
  kEdSeqShort = 7,      // alignment.rs:72-82 "...consume a record base..."
  kEdRefLeft = 8,       // alignment.rs:96-100 "reference sequence was not fully consumed"
  kEdSeqLeft = 9,       // alignment.rs:102-106 "record sequence was not fully consumed"
  kEdTooMany = 10,      // more than 512 edits in one read: Histogram::increment fails, unwrap (edits.rs:296-300)
  kEdPosition = 11,     // a matched position beyond the header's sequence length (edits.rs:281-290)
  kEdBadCigar = 12,     // CIGAR op kind > 8
};

struct EditsContig {          // one per reference sequence of the header
  const uint8_t* codes;       // code of every FASTA base; nullptr: no sequence loaded
  const uint32_t* bad_bits;   // code_len / 32 + 1 words each
  const uint32_t* bad_prefix;
  uint64_t code_len;          // bases in the FASTA sequence
  uint64_t pos_off;           // offset of position 0 in refs[] / alts[]
  uint32_t hdr_len;           // sequence length of the BAM header
  uint32_t pad;
};

NGSQ_HD uint32_t ed_ld32(const uint8_t* p) {  // any alignment
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// letters noodles' sam::record::sequence::Base accepts: the sixteen of the BAM code table, upper case
NGSQ_HD uint8_t edits_letter_code(uint8_t ch) {
  const char* tab = "=ACMGRSVTWYHKDBN";
  for (uint32_t i = 0; i < 16; ++i)
    if ((uint8_t)tab[i] == ch) return (uint8_t)i;
  return 0xFF;
}

// host side of ngsq_set_reference_bases: letters -> codes, rejected-letter bitmap and its per-word prefix counts
// (bits / prefix hold n / 32 + 1 words)
inline void edits_encode(const uint8_t* letters, uint64_t n, uint8_t* codes, uint32_t* bits, uint32_t* prefix) {
  uint8_t lut[256];
  for (uint32_t c = 0; c < 256; ++c) lut[c] = edits_letter_code((uint8_t)c);
  const uint64_t words = n / 32 + 1;
  uint32_t run = 0;
  for (uint64_t w = 0; w < words; ++w) {
    uint32_t m = 0;
    const uint64_t lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
    for (uint64_t i = lo; i < hi; ++i) {
      const uint8_t c = lut[letters[i]];
      codes[i] = c;
      if (c == 0xFF) m |= 1u << (i - lo);
    }
    bits[w] = m;
    prefix[w] = run;
    run += (uint32_t)__builtin_popcount(m);
  }
}

// rejected letters in [a, b) of a contig (a <= b <= code_len)
NGSQ_HD uint32_t edits_bad_in(const uint32_t* bits, const uint32_t* prefix, uint64_t a, uint64_t b) {
  const uint32_t ra = prefix[a >> 5] + (uint32_t)
#if defined(__CUDA_ARCH__)
      __popc
#else
      __builtin_popcount
#endif
      (bits[a >> 5] & ((1u << (a & 31)) - 1u));
  const uint32_t rb = prefix[b >> 5] + (uint32_t)
#if defined(__CUDA_ARCH__)
      __popc
#else
      __builtin_popcount
#endif
      (bits[b >> 5] & ((1u << (b & 31)) - 1u));
  return rb - ra;
}

// One record.  rec points at the record's block_size field (any alignment).  on_match(p, is_edit) is called for every
// Kind::Match position p (1-based) with p <= hdr_len, in stream order.  *edits, *first are set when the record counted.
// The checks follow the order in which the reference meets them: flags, name, slice, letters, walk.
template <class OnMatch>
NGSQ_HD uint32_t edits_record(const uint8_t* rec, int32_t n_ref, const EditsContig* contigs, uint32_t* edits, bool* first, OnMatch on_match) {
  const int32_t ref = (int32_t)ed_ld32(rec + 4), pos = (int32_t)ed_ld32(rec + 8);
  const uint32_t w3 = ed_ld32(rec + 12), w4 = ed_ld32(rec + 16), lseq = ed_ld32(rec + 20);
  const uint32_t lname = w3 & 255, ncig = w4 & 0xFFFF, flag = w4 >> 16;
  if (ref < 0 || ref >= n_ref || pos < 0) return kEdSkipped;  // not returned by the per-contig query (command.rs:369-377)
  // a record whose fields overrun its block_size fails the run in the facet kernel (R_ERR_RECORD): never walk it
  if (32ull + lname + 4ull * ncig + (lseq + 1ull) / 2 + lseq > ed_ld32(rec)) return kEdSkipped;
  const uint8_t* cig = rec + 36 + lname;
  const uint8_t* seq = cig + 4 * (size_t)ncig;
  uint64_t span = 0;
  for (uint32_t i = 0; i < ncig; ++i) {
    const uint32_t op = ed_ld32(cig + 4 * i), k = op & 15;
    if (k > 8) return kEdBadCigar;
    if ((0x18Du >> k) & 1u) span += op >> 4;  // M D N = X consume the reference (utils/cigar.rs:6-11)
  }
  const EditsContig& C = contigs[ref];
  const uint64_t start = (uint64_t)pos + 1, end = start + span - 1;  // 1-based, inclusive
  if (end == 0 || !(start <= C.hdr_len && end >= 1)) return kEdSkipped;  // the query's interval filter
  if (flag & (0x4u | 0x400u)) return kEdSkipped;                          // unmapped or duplicate (edits.rs:227-229)
  if (lname <= 1 || (lname == 2 && rec[36] == '*')) return kEdNoName;
  if (!C.codes) return kEdNoSequence;
  if ((uint64_t)pos + span > C.code_len) return kEdSlice;
  if (span && edits_bad_in(C.bad_bits, C.bad_prefix, (uint64_t)pos, (uint64_t)pos + span)) return kEdBadBase;
  const uint8_t* rc = C.codes + (uint64_t)pos;  // the slice
  uint64_t rp = 0, qp = 0;                                  // reference_ptr, record_ptr (alignment.rs:48-52)
  uint32_t e = 0, overflow = 0;
  for (uint32_t i = 0; i < ncig; ++i) {
    const uint32_t op = ed_ld32(cig + 4 * i), k = op & 15, len = op >> 4;
    const bool cr = (0x18Du >> k) & 1u;   // M D N = X
    const bool cs = (0x193u >> k) & 1u;   // M I S = X consume the read (utils/cigar.rs:13-23)
    if (cr && rp + len > span) return kEdRefShort;
    if (cs && qp + len > lseq) return kEdSeqShort;
    if (k == 0) {  // Kind::Match only: "=" and "X" are not compared (edits.rs:274)
      for (uint32_t t = 0; t < len; ++t) {
        const uint64_t qi = qp + t;
        const uint32_t qb = (qi & 1) ? (seq[qi >> 1] & 15u) : (uint32_t)(seq[qi >> 1] >> 4);
        const uint32_t is_edit = qb != rc[rp + t];
        e += is_edit;
        const uint64_t p = start + rp + t;
        if (p > C.hdr_len) overflow = 1; else on_match(p, is_edit != 0);
      }
    }
    if (cr) rp += len;
    if (cs) qp += len;
  }
  if (rp != span) return kEdRefLeft;
  if (qp != lseq) return kEdSeqLeft;
  if (overflow) return kEdPosition;
  if (e > 512) return kEdTooMany;
  *edits = e;
  *first = (flag & 0x40u) != 0;
  return kEdCounted;
}

// (alts as f32 / total as f32 * 100.0) as usize  (edits.rs:326-333): IEEE single-precision divide, then multiply
NGSQ_HD uint32_t edits_vaf_bin(uint32_t refs, uint32_t alts) {
  const float v = (float)alts / (float)(refs + alts);
  return (uint32_t)(v * 100.0f);
}

#if defined(__CUDACC__)

struct EditsParams {
  const uint8_t* d;            // base of the wave's slot
  const uint64_t* rec;         // record table of the wave (recscan.cuh): slot offset in the low 40 bits
  const RunState* st;          // wave_rec, fatal
  int32_t n_ref;
  const EditsContig* contigs;
  uint32_t* refs;
  uint32_t* alts;
  unsigned long long* res;     // E_* words
  const uint32_t* mark;        // `-n`: 1 = the second pass processes this record (cov_n.cuh); nullptr = every record
};

__global__ void __launch_bounds__(256) edits_kernel(EditsParams P) {
  __shared__ unsigned int s_hist[2][64];  // edits 0..63 of read one / read two; larger counts go straight to global
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  uint32_t n_counted = 0, err = 0;
  const uint64_t n_rec = P.st->fatal ? 0 : P.st->wave_rec;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (uint64_t)gridDim.x * blockDim.x) {
    if (P.mark && !P.mark[r]) continue;  // not yielded, or behind the second pass's record counter
    const uint8_t* rec = P.d + (P.rec[r] & kRecOffMask);
    const int32_t ref = (int32_t)ed_ld32(rec + 4);
    uint32_t* refs = P.refs;
    uint32_t* alts = P.alts;
    uint64_t poff = 0;
    if (ref >= 0 && ref < P.n_ref) poff = P.contigs[ref].pos_off;
    uint32_t e = 0;
    bool first = false;
    const uint32_t st = edits_record(rec, P.n_ref, P.contigs, &e, &first,
                                     [=](uint64_t p, bool is_edit) { atomicAdd((is_edit ? alts : refs) + poff + p, 1u); });
    if (st == kEdCounted) {
      ++n_counted;
      if (e < 64) atomicAdd(&s_hist[first ? 0 : 1][e], 1u);
      else atomicAdd(&P.res[(first ? E_READ_ONE : E_READ_TWO) + e], 1ull);
    } else if (st != kEdSkipped) {
      err = err > st ? err : st;
    }
  }
  // a failed record may have left per-position counts behind: the run is reported as failed, nothing is read back
  if (err) atomicMax(&P.res[E_ERR], (unsigned long long)err);
  n_counted = __reduce_add_sync(0xFFFFFFFFu, n_counted);
  if ((threadIdx.x & 31) == 0 && n_counted) atomicAdd(&P.res[E_RECORDS], (unsigned long long)n_counted);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) {
    const unsigned int v = (&s_hist[0][0])[i];
    if (v) atomicAdd(&P.res[(i < 64 ? E_READ_ONE : E_READ_TWO) + (i & 63)], (unsigned long long)v);
  }
}

// teardown (edits.rs:318-334) for all loaded contigs at once: positions [0, n) of refs[] / alts[]
__global__ void __launch_bounds__(256) edits_vaf_kernel(const uint32_t* __restrict__ refs, const uint32_t* __restrict__ alts, uint64_t n,
                                                        unsigned long long* __restrict__ res) {
  __shared__ unsigned int s_vaf[101];
  for (uint32_t i = threadIdx.x; i < 101; i += blockDim.x) s_vaf[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = refs[i], a = alts[i];
    if (r + a) atomicAdd(&s_vaf[edits_vaf_bin(r, a)], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 101; i += blockDim.x)
    if (s_vaf[i]) atomicAdd(&res[E_VAF + i], (unsigned long long)s_vaf[i]);
}

#endif  // __CUDACC__

}  // namespace ngsq
