/* ngsqc_oracle.c — CPU restatement of the reference `ngs qc` BAM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ngs_b200/, the
 * C-ABI library, the host driver) may link, import or execute this file; it
 * is the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED at the decode boundary: the reference (stjude-rust-labs/ngs
 * v0.4.0) cannot be built here (no cargo/rustc, crates not vendored) and its
 * BGZF/BAM/BAI arithmetic lives in noodles 0.34.0 (noodles-bam 0.28.0,
 * noodles-bgzf 0.20.0, noodles-sam 0.25.0, noodles-csi 0.14.0; inflate via
 * flate2 1.0.24 -> miniz_oxide 0.5.4), which is absent from /root/reference
 * (Cargo.lock:902-1072).  The reference has no test that reads a BAM or calls
 * a facet's process().  What IS pinned: the Histogram known-answer vectors of
 * src/utils/histogram.rs:405-523 (tests/test_oracle_kat.py drives
 * oracle_hist_* below against them).  Decode follows the SAM/BAM spec; inflate
 * is zlib 1.3 (any conformant inflater yields the same bytes).
 *
 * Each function cites the reference lines it restates.  The structure mirrors
 * the reference on purpose: single thread, two passes (the file is inflated
 * twice), one record at a time.
 *
 * One documented deviation: the reference draws the GC window
 * offset from an OS-seeded ThreadRng (gc_content.rs:69-74), so two reference
 * runs disagree with each other.  Oracle and engine share the policy
 *   offset = ((splitmix64(gc_seed ^ record_virtual_offset) >> 32) * (l_seq - 100)) >> 32
 * keyed by the record's BGZF virtual offset, which is shard-independent.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define MAX_SCORE 93          /* quality_scores.rs:26 */
#define TRUNCATION_LENGTH 100 /* gc_content.rs:20 */
#define TLEN_CAP 1024         /* qc.rs:62 */
#define COV_BIN 50000         /* qc.rs:86-89 */
#define COV_HIST 2048         /* coverage.rs:76 */

/* ------------------------------------------------------------------ errors */
static char g_err[512];
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return -1; } while (0)
const char* oracle_last_error(void) { return g_err; }

/* ---------------------------------------------------- Histogram (a13) ---- */
/* src/utils/histogram.rs:152-392 */
typedef struct { uint64_t* v; uint64_t cap; } hist_t;

static void hist_init(hist_t* h, uint64_t cap) { h->v = calloc(cap + 1, 8); h->cap = cap; }
/* histogram.rs:190-197: Err if bin out of [0, cap] */
static int hist_inc_by(hist_t* h, uint64_t bin, uint64_t n) { if (bin > h->cap) return -1; h->v[bin] += n; return 0; }
/* histogram.rs:258-269 */
static double hist_mean(const hist_t* h) {
  double sum = 0.0, den = 0.0;
  for (uint64_t i = 0; i <= h->cap; ++i) { den += (double)h->v[i]; sum += (double)(h->v[i] * i); }
  return sum / den;
}
/* histogram.rs:272-337; returns 0 and *out when Some, 1 when None, -1 on the
 * paths where the reference panics (runs off the end looking for a non-zero bin). */
static int hist_percentile(const hist_t* h, double p, double* out) {
  uint64_t n = 0;
  for (uint64_t i = 0; i <= h->cap; ++i) n += h->v[i];
  if (n == 0) return 1;
  double need = p * (double)n, got = 0.0;
  uint64_t idx = 0;
  for (;;) {
    if (idx > h->cap) return -1;
    got += (double)h->v[idx];
    if (got > need) { *out = (double)idx; return 0; }
    if (got == need) {
      uint64_t lo = idx;
      ++idx;
      for (;;) { if (idx > h->cap) return -1; if (h->v[idx] != 0) break; ++idx; }
      *out = (double)lo + ((double)(idx - lo) / 2.0);
      return 0;
    }
    ++idx;
  }
}
static uint64_t hist_sum(const hist_t* h) { uint64_t s = 0; for (uint64_t i = 0; i <= h->cap; ++i) s += h->v[i]; return s; }
/* histogram.rs:384-391 */
static uint64_t hist_top_until(const hist_t* h, uint64_t bin) { uint64_t s = 0; for (uint64_t i = bin; i <= h->cap; ++i) s += h->v[i]; return s; }
static uint64_t hist_bottom_until(const hist_t* h, uint64_t bin) { uint64_t s = 0; for (uint64_t i = 0; i <= bin && i <= h->cap; ++i) s += h->v[i]; return s; }

/* KAT entry points (tests drive these against histogram.rs:405-523) */
void* oracle_hist_new(uint64_t cap) { hist_t* h = malloc(sizeof *h); hist_init(h, cap); return h; }
void oracle_hist_free(void* p) { hist_t* h = p; free(h->v); free(h); }
int oracle_hist_inc_by(void* p, uint64_t bin, uint64_t n) { return hist_inc_by(p, bin, n); }
uint64_t oracle_hist_get(void* p, uint64_t bin) { return ((hist_t*)p)->v[bin]; }
uint64_t oracle_hist_len(void* p) { return ((hist_t*)p)->cap + 1; }
double oracle_hist_mean(void* p) { return hist_mean(p); }
int oracle_hist_percentile(void* p, double q, double* out) { return hist_percentile(p, q, out); }
uint64_t oracle_hist_sum(void* p) { return hist_sum(p); }
uint64_t oracle_hist_top_until(void* p, uint64_t b) { return hist_top_until(p, b); }
uint64_t oracle_hist_bottom_until(void* p, uint64_t b) { return hist_bottom_until(p, b); }

/* ------------------------------------------------------- BGZF reader (a1) */
/* Restates noodles-bgzf's blocking reader as used at utils/formats/bam.rs:41-44
 * and qc/command.rs:350: gzip member framing, raw inflate, CRC32 + ISIZE check,
 * empty blocks skipped, virtual positions, seek. */
typedef struct {
  const uint8_t* data; size_t len;
  size_t next_coff;             /* file offset of the next block to load */
  uint64_t cur_coff;            /* file offset of the loaded block */
  uint8_t buf[65536]; uint32_t cur_len, cur_off;
  uint64_t inflated_total, blocks;
} bgzf_t;

static void bgzf_open(bgzf_t* r, const uint8_t* d, size_t n) { r->data = d; r->len = n; r->cur_len = r->cur_off = 0; r->next_coff = 0; r->cur_coff = 0; r->inflated_total = 0; r->blocks = 0; }

/* loads the block at next_coff; returns 1 ok, 0 eof, -1 error */
static int bgzf_load(bgzf_t* r) {
  for (;;) {
    size_t o = r->next_coff;
    if (o >= r->len) return 0;
    if (r->len - o < 18) { snprintf(g_err, sizeof g_err, "truncated BGZF header at %zu", o); return -1; }
    const uint8_t* p = r->data + o;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) { snprintf(g_err, sizeof g_err, "bad BGZF magic at %zu", o); return -1; }
    uint32_t xlen = p[10] | (p[11] << 8), bsize = 0; int found = 0;
    for (uint32_t q = 0; q + 4 <= xlen;) {
      const uint8_t* s = p + 12 + q; uint32_t slen = s[2] | (s[3] << 8);
      if (s[0] == 'B' && s[1] == 'C' && slen == 2) { bsize = s[4] | (s[5] << 8); found = 1; }
      q += 4 + slen;
    }
    if (!found) { snprintf(g_err, sizeof g_err, "no BC subfield at %zu", o); return -1; }
    size_t total = (size_t)bsize + 1;
    if (total < 12 + xlen + 8 || o + total > r->len) { snprintf(g_err, sizeof g_err, "truncated BGZF block at %zu", o); return -1; }
    const uint8_t* cdata = p + 12 + xlen; size_t clen = total - 12 - xlen - 8;
    uint32_t crc, isize; memcpy(&crc, p + total - 8, 4); memcpy(&isize, p + total - 4, 4);
    if (isize > 65536) { snprintf(g_err, sizeof g_err, "ISIZE too large at %zu", o); return -1; }
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { snprintf(g_err, sizeof g_err, "inflateInit2"); return -1; }
    zs.next_in = (Bytef*)cdata; zs.avail_in = (uInt)clen; zs.next_out = r->buf; zs.avail_out = 65536;
    int rc = inflate(&zs, Z_FINISH); uint32_t got = (uint32_t)zs.total_out; inflateEnd(&zs);
    if (rc != Z_STREAM_END || got != isize) { snprintf(g_err, sizeof g_err, "inflate failed at %zu (rc %d, %u vs ISIZE %u)", o, rc, got, isize); return -1; }
    if ((uint32_t)crc32(crc32(0, NULL, 0), r->buf, got) != crc) { snprintf(g_err, sizeof g_err, "CRC mismatch at %zu", o); return -1; }
    r->cur_coff = o; r->next_coff = o + total; r->cur_len = got; r->cur_off = 0;
    r->inflated_total += got; r->blocks++;
    if (got == 0) continue; /* empty blocks (incl. EOF marker) are skipped */
    return 1;
  }
}
static int bgzf_read(bgzf_t* r, uint8_t* dst, size_t n) { /* 1 ok, 0 clean eof before any byte, -1 error/short */
  size_t done = 0;
  while (done < n) {
    if (r->cur_off == r->cur_len) { int rc = bgzf_load(r); if (rc < 0) return -1; if (rc == 0) { if (done == 0) return 0; snprintf(g_err, sizeof g_err, "unexpected EOF inside record"); return -1; } }
    size_t k = r->cur_len - r->cur_off; if (k > n - done) k = n - done;
    memcpy(dst + done, r->buf + r->cur_off, k); r->cur_off += (uint32_t)k; done += k;
  }
  return 1;
}
/* virtual position of the next byte; a fully consumed block reports the start of the next one */
static uint64_t bgzf_tell(const bgzf_t* r) {
  if (r->cur_off == r->cur_len) return (uint64_t)r->next_coff << 16;
  return (r->cur_coff << 16) | r->cur_off;
}
static int bgzf_seek(bgzf_t* r, uint64_t v) {
  r->next_coff = (size_t)(v >> 16); r->cur_len = r->cur_off = 0;
  uint32_t u = (uint32_t)(v & 0xFFFF);
  if (r->next_coff >= r->len) return 0;
  int rc = bgzf_load(r); if (rc <= 0) return rc;
  /* bgzf_load skips empty blocks; the uoffset applies only to the addressed block */
  if (r->cur_coff == (v >> 16)) { if (u > r->cur_len) { snprintf(g_err, sizeof g_err, "bad seek uoffset"); return -1; } r->cur_off = u; }
  return 1;
}

/* ------------------------------------------------------------ BAM (a2,a3) */
typedef struct { char* name; uint32_t len; int primary; int known; } ref_t;
typedef struct {
  int32_t ref_id, pos, next_ref, next_pos, tlen;
  uint32_t l_seq; uint16_t flag, n_cigar; uint8_t mapq, l_name;
  const uint32_t* cigar; const uint8_t* seq; const uint8_t* qual;
  uint64_t voff; uint32_t span;
} rec_t;
typedef struct { uint8_t* buf; size_t cap; } recbuf_t;

/* primary assembly of GRCh38_no_alt_AnalysisSet: autosomes + sex + unlocalized +
 * unplaced (utils/genome.rs:59-83, genome/ncbi/grch38_no_alt.rs:46-283).  The name
 * rule below is equivalent to that table: chr1..22, chrX, chrY, chr*_random, chrUn_*. */
static int is_known_name(const char* n, int* primary) {
  *primary = 0;
  if (strncmp(n, "chr", 3) != 0) return 0;
  const char* s = n + 3;
  if (!strcmp(s, "M") || !strcmp(s, "EBV")) return 1;
  if (!strcmp(s, "X") || !strcmp(s, "Y")) { *primary = 1; return 1; }
  char* e; long v = strtol(s, &e, 10);
  if (e != s && *e == 0 && v >= 1 && v <= 22 && s[0] != '0') { *primary = 1; return 1; }
  size_t L = strlen(n);
  if (!strncmp(n, "chrUn_", 6) || (L > 7 && !strcmp(n + L - 7, "_random"))) { *primary = 1; return 1; }
  return 0;
}

static int read_header(bgzf_t* r, ref_t** refs_out, uint32_t* n_ref_out) {
  uint8_t b[8];
  if (bgzf_read(r, b, 8) != 1 || memcmp(b, "BAM\1", 4)) FAIL("bad BAM magic");
  uint32_t l_text; memcpy(&l_text, b + 4, 4);
  uint8_t* text = malloc(l_text + 1);
  if (l_text && bgzf_read(r, text, l_text) != 1) FAIL("short header text");
  free(text);
  if (bgzf_read(r, b, 4) != 1) FAIL("short n_ref");
  uint32_t n_ref; memcpy(&n_ref, b, 4);
  ref_t* refs = calloc(n_ref ? n_ref : 1, sizeof *refs);
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (bgzf_read(r, b, 4) != 1) FAIL("short ref");
    uint32_t l; memcpy(&l, b, 4);
    refs[i].name = malloc(l + 1);
    if (bgzf_read(r, (uint8_t*)refs[i].name, l) != 1) FAIL("short ref name");
    refs[i].name[l] = 0;
    if (bgzf_read(r, b, 4) != 1) FAIL("short ref len");
    memcpy(&refs[i].len, b, 4);
    refs[i].known = is_known_name(refs[i].name, &refs[i].primary);
  }
  *refs_out = refs; *n_ref_out = n_ref;
  return 0;
}

/* 1 = record, 0 = EOF, -1 = error. Decode rules: SAM/BAM spec 4.2 + SURVEY App. D. */
static int read_record(bgzf_t* r, recbuf_t* rb, rec_t* rec, uint32_t n_ref) {
  uint64_t v = bgzf_tell(r);
  if (r->cur_off == r->cur_len) { /* normalise to the block the first byte really lives in */
    int rc = bgzf_load(r); if (rc < 0) return -1; if (rc == 0) return 0;
    v = (r->cur_coff << 16) | r->cur_off;
  }
  uint8_t b4[4];
  int rc = bgzf_read(r, b4, 4); if (rc <= 0) return rc;
  uint32_t bs; memcpy(&bs, b4, 4);
  if (bs < 32) { snprintf(g_err, sizeof g_err, "record block_size %u < 32", bs); return -1; }
  if (bs > rb->cap) { rb->cap = bs * 2; rb->buf = realloc(rb->buf, rb->cap); }
  if (bgzf_read(r, rb->buf, bs) != 1) { if (!g_err[0]) snprintf(g_err, sizeof g_err, "short record"); return -1; }
  const uint8_t* p = rb->buf;
  memcpy(&rec->ref_id, p, 4); memcpy(&rec->pos, p + 4, 4);
  rec->l_name = p[8]; rec->mapq = p[9];
  memcpy(&rec->n_cigar, p + 12, 2); memcpy(&rec->flag, p + 14, 2); memcpy(&rec->l_seq, p + 16, 4);
  memcpy(&rec->next_ref, p + 20, 4); memcpy(&rec->next_pos, p + 24, 4); memcpy(&rec->tlen, p + 28, 4);
  uint64_t need = 32ull + rec->l_name + 4ull * rec->n_cigar + (rec->l_seq + 1ull) / 2 + rec->l_seq;
  if (need > bs) { snprintf(g_err, sizeof g_err, "record fields overrun block_size"); return -1; }
  if (rec->ref_id < -1 || rec->ref_id >= (int32_t)n_ref || rec->next_ref < -1 || rec->next_ref >= (int32_t)n_ref) { snprintf(g_err, sizeof g_err, "reference id out of range"); return -1; }
  rec->cigar = (const uint32_t*)(p + 32 + rec->l_name); /* alignment: buf is malloc'd, 32+l_name may be odd -> memcpy on use */
  rec->seq = p + 32 + rec->l_name + 4ull * rec->n_cigar;
  rec->qual = rec->seq + (rec->l_seq + 1) / 2;
  rec->voff = v;
  uint32_t span = 0;
  for (uint32_t i = 0; i < rec->n_cigar; ++i) {
    uint32_t op; memcpy(&op, (const uint8_t*)rec->cigar + 4 * i, 4);
    uint32_t k = op & 15;
    if (k > 8) { snprintf(g_err, sizeof g_err, "invalid CIGAR op %u", k); return -1; }
    if (k == 0 || k == 2 || k == 3 || k == 7 || k == 8) span += op >> 4; /* utils/cigar.rs:6-11 */
  }
  rec->span = span;
  /* quality scores: l_seq==0 or all 0xFF -> empty; else each must be <= 93 (App. D.5) */
  return 1;
}
static int qual_present(const rec_t* r, int* bad) {
  *bad = 0;
  if (r->l_seq == 0) return 0;
  int all_ff = 1;
  for (uint32_t i = 0; i < r->l_seq; ++i) if (r->qual[i] != 0xFF) { all_ff = 0; break; }
  if (all_ff) return 0;
  for (uint32_t i = 0; i < r->l_seq; ++i) if (r->qual[i] > MAX_SCORE) { *bad = 1; return 0; }
  return 1;
}

/* ----------------------------------------------------------------- results */
enum { G_TOTAL, G_UNMAPPED, G_DUPLICATE, G_PRIMARY, G_SECONDARY, G_SUPPLEMENTARY, G_PRIMARY_MAPPED,
       G_PRIMARY_DUPLICATE, G_PAIRED, G_READ_1, G_READ_2, G_PROPER_PAIR, G_SINGLETON, G_MATE_MAPPED,
       G_MISMATCH, G_MISMATCH_HQ, G_NCOUNT };

typedef struct {
  uint32_t n_ref; ref_t* refs;
  int do_records, do_coverage;
  /* general (a5) */
  uint64_t general[G_NCOUNT]; uint64_t cigar_ops[2][9]; /* [0]=read one, [1]=read two */
  /* template length (a6) */
  hist_t tlen; uint64_t tlen_processed, tlen_ignored;
  /* gc (a8) */
  hist_t gc; uint64_t gc_nuc[3]; /* gc, at, other */ uint64_t gc_rec[3]; /* processed, ignored_flags, ignored_too_short */
  /* quality (a7) */
  uint64_t* qual; uint64_t qual_positions, qual_cap; /* qual[pos*94+score], pos 0-based; positions ever seen = [0,qual_positions) */
  /* coverage (a10-a12) */
  uint8_t* touched; uint64_t* cov_ignored; double *cov_mean, *cov_median, *cov_mom; double** cov_bins; uint64_t* cov_nbins;
  uint64_t** cov_bin_sums; hist_t* cov_hist_contig;
  hist_t cov_dist; uint64_t nonsensical; float covered_by[6];
  uint64_t pass1_records, pass2_records, inflated_bytes, bgzf_blocks;
} results_t;

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

/* general.rs:31-124 */
static int general_process(results_t* R, const rec_t* r) {
  uint64_t* g = R->general; uint16_t f = r->flag;
  g[G_TOTAL]++;
  if (f & 0x4) g[G_UNMAPPED]++;
  if (f & 0x400) g[G_DUPLICATE]++;
  if (f & 0x100) g[G_SECONDARY]++;
  else if (f & 0x800) g[G_SUPPLEMENTARY]++;
  else {
    g[G_PRIMARY]++;
    if (!(f & 0x4)) g[G_PRIMARY_MAPPED]++;
    if (f & 0x400) g[G_PRIMARY_DUPLICATE]++;
    if (f & 0x1) {
      g[G_PAIRED]++;
      if (f & 0x40) g[G_READ_1]++;
      if (f & 0x80) g[G_READ_2]++;
      if (!(f & 0x4)) {
        if (f & 0x2) g[G_PROPER_PAIR]++;
        if (f & 0x8) g[G_SINGLETON]++;
        else {
          g[G_MATE_MAPPED]++;
          /* general.rs:81-83: unwrap() on both ids -> the reference panics */
          if (r->ref_id < 0 || r->next_ref < 0) FAIL("reference would panic: mapped pair without reference ids (general.rs:81-83)");
          if (r->ref_id != r->next_ref) {
            g[G_MISMATCH]++;
            if (r->mapq >= 5) g[G_MISMATCH_HQ]++; /* 255 (missing) counts, general.rs:88-95 */
          }
        }
      }
    }
  }
  int which = (f & 0x40) ? 0 : 1; /* general.rs:103-121 */
  for (uint32_t i = 0; i < r->n_cigar; ++i) { uint32_t op; memcpy(&op, (const uint8_t*)r->cigar + 4 * i, 4); R->cigar_ops[which][op & 15]++; }
  return 0;
}
/* template_length.rs:79-87: `tlen as usize` (sign-extends: negatives become huge) */
static void tlen_process(results_t* R, const rec_t* r) {
  uint64_t t = (uint64_t)(int64_t)r->tlen;
  if (hist_inc_by(&R->tlen, t, 1) == 0) R->tlen_processed++; else R->tlen_ignored++;
}
/* quality_scores.rs:37-49 */
static int quality_process(results_t* R, const rec_t* r) {
  int bad; int present = qual_present(r, &bad);
  if (bad) FAIL("quality score > 93: noodles rejects the record (App. D.5)");
  if (!present) return 0;
  if (r->l_seq > R->qual_cap) {
    uint64_t nc = R->qual_cap ? R->qual_cap : 256; while (nc < r->l_seq) nc *= 2;
    R->qual = realloc(R->qual, nc * 94 * 8); memset(R->qual + R->qual_cap * 94, 0, (nc - R->qual_cap) * 94 * 8); R->qual_cap = nc;
  }
  if (r->l_seq > R->qual_positions) R->qual_positions = r->l_seq;
  for (uint32_t i = 0; i < r->l_seq; ++i) R->qual[(uint64_t)i * 94 + r->qual[i]]++;
  return 0;
}
/* gc_content.rs:38-100 */
static void gc_process(results_t* R, const rec_t* r, uint64_t gc_seed) {
  if (r->flag & (0x400 | 0x100)) { R->gc_rec[1]++; return; }
  if (r->l_seq < TRUNCATION_LENGTH) { R->gc_rec[2]++; return; }
  uint64_t offset = 0;
  if (TRUNCATION_LENGTH < r->l_seq) offset = ((splitmix64(gc_seed ^ r->voff) >> 32) * (uint64_t)(r->l_seq - TRUNCATION_LENGTH)) >> 32; /* gen_range(0..max_offset): exclusive upper bound */
  uint64_t gc_this = 0;
  for (uint64_t i = 0; i < TRUNCATION_LENGTH; ++i) {
    uint64_t k = offset + i; uint8_t code = (k & 1) ? (r->seq[k >> 1] & 15) : (r->seq[k >> 1] >> 4);
    if (code == 2 || code == 4) { gc_this++; R->gc_nuc[0]++; }      /* C, G */
    else if (code == 1 || code == 8) R->gc_nuc[1]++;                /* A, T */
    else R->gc_nuc[2]++;
  }
  uint64_t pct = (uint64_t)round(((double)gc_this / (double)TRUNCATION_LENGTH) * 100.0);
  hist_inc_by(&R->gc, pct, 1);
  R->gc_rec[0]++;
}

/* --------------------------------------------------------------- BAI (a9) */
typedef struct { uint64_t beg, end; } chunk_t;
typedef struct { uint32_t n_bin; uint32_t* bin_id; uint32_t* n_chunk; chunk_t** chunks; uint32_t n_intv; uint64_t* ioff; } bai_ref_t;
typedef struct { uint32_t n_ref; bai_ref_t* refs; } bai_t;

static int bai_parse(const uint8_t* d, size_t n, bai_t* out) {
  size_t o = 0;
#define NEED(k) do { if (o + (k) > n) FAIL("truncated BAI"); } while (0)
  NEED(8); if (memcmp(d, "BAI\1", 4)) FAIL("bad BAI magic");
  memcpy(&out->n_ref, d + 4, 4); o = 8;
  out->refs = calloc(out->n_ref ? out->n_ref : 1, sizeof *out->refs);
  for (uint32_t r = 0; r < out->n_ref; ++r) {
    bai_ref_t* R = &out->refs[r];
    NEED(4); memcpy(&R->n_bin, d + o, 4); o += 4;
    R->bin_id = calloc(R->n_bin ? R->n_bin : 1, 4); R->n_chunk = calloc(R->n_bin ? R->n_bin : 1, 4); R->chunks = calloc(R->n_bin ? R->n_bin : 1, sizeof(chunk_t*));
    for (uint32_t b = 0; b < R->n_bin; ++b) {
      NEED(8); memcpy(&R->bin_id[b], d + o, 4); memcpy(&R->n_chunk[b], d + o + 4, 4); o += 8;
      NEED(16ull * R->n_chunk[b]); R->chunks[b] = malloc(16ull * R->n_chunk[b] + 1); memcpy(R->chunks[b], d + o, 16ull * R->n_chunk[b]); o += 16ull * R->n_chunk[b];
    }
    NEED(4); memcpy(&R->n_intv, d + o, 4); o += 4;
    NEED(8ull * R->n_intv); R->ioff = malloc(8ull * R->n_intv + 1); memcpy(R->ioff, d + o, 8ull * R->n_intv); o += 8ull * R->n_intv;
  }
#undef NEED
  return 0;
}
static int cmp_chunk(const void* a, const void* b) { const chunk_t *x = a, *y = b; return x->beg < y->beg ? -1 : x->beg > y->beg ? 1 : 0; }
/* noodles-csi query: bins overlapping [beg,end) (0-based half-open), chunks whose end > linear
 * min offset, sorted and merged.  Used at qc/command.rs:369-373 with the whole contig. */
static size_t bai_query(const bai_t* B, uint32_t ref, int64_t beg, int64_t end, chunk_t** out) {
  *out = NULL; if (ref >= B->n_ref) return 0;
  const bai_ref_t* R = &B->refs[ref];
  if (end > (1ll << 29)) end = 1ll << 29;
  --end;
  size_t n = 0, cap = 64; chunk_t* v = malloc(cap * sizeof *v);
  uint64_t min_off = 0; { uint64_t w = (uint64_t)beg >> 14; if (w < R->n_intv) min_off = R->ioff[w]; else if (R->n_intv) min_off = 0; }
  for (uint32_t b = 0; b < R->n_bin; ++b) {
    uint32_t id = R->bin_id[b]; if (id == 37450) continue;
    int hit = id == 0;
    static const int first[5] = {1, 9, 73, 585, 4681}; static const int shift[5] = {26, 23, 20, 17, 14};
    for (int l = 0; l < 5 && !hit; ++l) if (id >= (uint32_t)(first[l] + (beg >> shift[l])) && id <= (uint32_t)(first[l] + (end >> shift[l])) && id < (uint32_t)(l < 4 ? first[l + 1] : 37449)) hit = 1;
    if (!hit) continue;
    for (uint32_t c = 0; c < R->n_chunk[b]; ++c) if (R->chunks[b][c].end > min_off) { if (n == cap) { cap *= 2; v = realloc(v, cap * sizeof *v); } v[n++] = R->chunks[b][c]; }
  }
  qsort(v, n, sizeof *v, cmp_chunk);
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) { if (m && v[i].beg <= v[m - 1].end) { if (v[i].end > v[m - 1].end) v[m - 1].end = v[i].end; } else v[m++] = v[i]; }
  *out = v; return m;
}

/* ------------------------------------------------------------- the two passes */
static void results_init(results_t* R) {
  memset(R, 0, sizeof *R);
  hist_init(&R->tlen, TLEN_CAP); hist_init(&R->gc, 100); hist_init(&R->cov_dist, COV_HIST);
}

/* coverage.rs:182-262 */
static int coverage_teardown(results_t* R, uint32_t c, const uint32_t* depth) {
  uint64_t L = R->refs[c].len;
  hist_t cov; hist_init(&cov, COV_HIST);
  uint64_t ignored = 0, binsum = 0, nb = 0, capb = L / COV_BIN + 3;
  double* bins = malloc(capb * sizeof(double)); uint64_t* sums = malloc(capb * 8);
  for (uint64_t i = 0; i <= L; ++i) {
    uint64_t d = depth[i];
    if (hist_inc_by(&cov, d, 1)) ignored++;
    binsum += d;
    if (i % COV_BIN == 0) { sums[nb] = binsum; bins[nb++] = (double)binsum / (double)COV_BIN; binsum = 0; }
  }
  uint64_t mod = L % COV_BIN;
  if (mod != 0) { sums[nb] = binsum; bins[nb++] = (double)binsum / (double)mod; }
  double mean = hist_mean(&cov), median;
  if (hist_percentile(&cov, 0.5, &median) != 0) FAIL("reference would panic: coverage median undefined (coverage.rs:233)");
  for (uint64_t i = 0; i <= COV_HIST; ++i) hist_inc_by(&R->cov_dist, i, cov.v[i]);
  R->cov_mean[c] = mean; R->cov_median[c] = median; R->cov_mom[c] = median / mean; R->cov_ignored[c] = ignored;
  R->cov_bins[c] = bins; R->cov_nbins[c] = nb; R->cov_bin_sums[c] = sums; R->cov_hist_contig[c] = cov;
  return 0;
}

typedef struct { uint64_t n_records; /* 0 = all (-n) */ uint64_t gc_seed; int do_records, do_coverage; } oracle_opts;

static int run(const uint8_t* bam, size_t bam_len, const uint8_t* bai, size_t bai_len, const oracle_opts* opt, results_t* R) {
  g_err[0] = 0;
  results_init(R);
  R->do_records = opt->do_records; R->do_coverage = opt->do_coverage;
  bgzf_t* rd = malloc(sizeof *rd); recbuf_t rb = {malloc(1 << 16), 1 << 16}; rec_t rec;
  bgzf_open(rd, bam, bam_len);
  if (read_header(rd, &R->refs, &R->n_ref)) return -1;
  /* qc/command.rs:258-272: every binary reference name must be in the genome */
  for (uint32_t i = 0; i < R->n_ref; ++i) if (!R->refs[i].known) FAIL("Sequence \"%s\" not found in specified reference genome.", R->refs[i].name);
  /* utils/formats/bam.rs:86-96: the index must exist and parse */
  bai_t B; if (!bai || bai_parse(bai, bai_len, &B)) { if (!g_err[0]) snprintf(g_err, sizeof g_err, "missing BAM index"); return -1; }

  if (opt->do_records) { /* pass 1, command.rs:305-316 */
    uint64_t counter = 0; int rc;
    while ((rc = read_record(rd, &rb, &rec, R->n_ref)) == 1) {
      if (general_process(R, &rec)) return -1;
      tlen_process(R, &rec);
      gc_process(R, &rec, opt->gc_seed);
      if (quality_process(R, &rec)) return -1;
      ++counter;
      if (opt->n_records && counter >= opt->n_records) break;
    }
    if (rc < 0) return -1;
    R->pass1_records = counter; R->inflated_bytes = rd->inflated_total; R->bgzf_blocks = rd->blocks;
  }
  if (opt->do_coverage) { /* pass 2, command.rs:350-397: fresh reader, per contig in header order */
    uint32_t n = R->n_ref;
    R->touched = calloc(n + 1, 1); R->cov_ignored = calloc(n + 1, 8); R->cov_mean = calloc(n + 1, 8); R->cov_median = calloc(n + 1, 8); R->cov_mom = calloc(n + 1, 8);
    R->cov_bins = calloc(n + 1, sizeof(double*)); R->cov_nbins = calloc(n + 1, 8); R->cov_bin_sums = calloc(n + 1, sizeof(uint64_t*)); R->cov_hist_contig = calloc(n + 1, sizeof(hist_t));
    bgzf_open(rd, bam, bam_len);
    uint64_t counter = 0;
    for (uint32_t c = 0; c < n; ++c) {
      uint64_t L = R->refs[c].len; int supported = R->refs[c].primary; /* coverage.rs:133-138 */
      uint32_t* depth = NULL;
      chunk_t* ch; size_t nch = bai_query(&B, c, 0, (int64_t)L, &ch);
      for (size_t k = 0; k < nch; ++k) {
        int src = bgzf_seek(rd, ch[k].beg); if (src < 0) return -1; if (src == 0) break;
        int stop = 0;
        while (bgzf_tell(rd) < ch[k].end) {
          int rc = read_record(rd, &rb, &rec, R->n_ref); if (rc < 0) return -1; if (rc == 0) break;
          /* noodles query filter (App. D.6): same reference, start and end present, interval intersects [1, L] */
          if (rec.ref_id != (int32_t)c || rec.pos < 0) continue;
          uint64_t start = (uint64_t)rec.pos + 1, end = start + rec.span - 1; /* end==0 -> None */
          if (end == 0) continue;
          if (!(start <= L && end >= 1)) continue;
          if (supported) { /* coverage.rs:148-180 */
            if (!depth) { depth = calloc(L + 1, 4); R->touched[c] = 1; }
            for (uint64_t i = start; i <= end; ++i) { if (i <= L) depth[i]++; else R->nonsensical++; }
          }
          ++counter;
          if (opt->n_records && counter >= opt->n_records) { stop = 1; break; } /* command.rs:384-388: breaks only this contig's loop */
        }
        if (stop) break;
      }
      free(ch);
      if (supported && depth) { if (coverage_teardown(R, c, depth)) return -1; free(depth); }
    }
    R->pass2_records = counter;
    /* coverage.rs:264-287 (f32 arithmetic) */
    uint64_t total = hist_sum(&R->cov_dist);
    for (uint32_t c = 0; c < n; ++c) if (R->touched[c]) total += R->cov_ignored[c];
    static const uint64_t X[6] = {10, 20, 30, 40, 50, 60};
    for (int k = 0; k < 6; ++k) R->covered_by[k] = ((float)hist_top_until(&R->cov_dist, X[k]) / (float)total) * 100.0f;
  }
  free(rd); free(rb.buf);
  return 0;
}

/* ------------------------------------------------------------------ JSON */
static void jnum(FILE* f, const char* fmt, double v) { /* floats always carry a '.' or exponent so parsers keep them floats */
  char b[64]; snprintf(b, sizeof b, fmt, v); fputs(b, f); if (!strpbrk(b, ".eE")) fputs(".0", f);
}
static void jf(FILE* f, double v) { if (isnan(v) || isinf(v)) fputs("null", f); else jnum(f, "%.17g", v); }
static void jhist(FILE* f, const uint64_t* v, uint64_t cap) {
  fputs("{\"values\":[", f);
  for (uint64_t i = 0; i <= cap; ++i) fprintf(f, "%s%llu", i ? "," : "", (unsigned long long)v[i]);
  fprintf(f, "],\"range_start\":0,\"range_stop\":%llu}", (unsigned long long)cap);
}
static const char* kOpNames[9] = {"M", "I", "D", "N", "S", "H", "P", "=", "X"};

/* results.rs:24-60 schema; map keys emitted in sorted-independent order (compare parsed) */
static int write_json(const results_t* R, FILE* f) {
  fputs("{", f);
  if (R->do_records) {
    const uint64_t* g = R->general;
    fprintf(f, "\"general\":{\"records\":{\"total\":%llu,\"unmapped\":%llu,\"duplicate\":%llu,\"designation\":{\"primary\":%llu,\"secondary\":%llu,\"supplementary\":%llu},"
               "\"primary_mapped\":%llu,\"primary_duplicate\":%llu,\"paired\":%llu,\"read_1\":%llu,\"read_2\":%llu,\"proper_pair\":%llu,\"singleton\":%llu,\"mate_mapped\":%llu,"
               "\"mate_reference_sequence_id_mismatch\":%llu,\"mate_reference_sequence_id_mismatch_hq\":%llu},",
            (unsigned long long)g[G_TOTAL], (unsigned long long)g[G_UNMAPPED], (unsigned long long)g[G_DUPLICATE], (unsigned long long)g[G_PRIMARY], (unsigned long long)g[G_SECONDARY],
            (unsigned long long)g[G_SUPPLEMENTARY], (unsigned long long)g[G_PRIMARY_MAPPED], (unsigned long long)g[G_PRIMARY_DUPLICATE], (unsigned long long)g[G_PAIRED],
            (unsigned long long)g[G_READ_1], (unsigned long long)g[G_READ_2], (unsigned long long)g[G_PROPER_PAIR], (unsigned long long)g[G_SINGLETON], (unsigned long long)g[G_MATE_MAPPED],
            (unsigned long long)g[G_MISMATCH], (unsigned long long)g[G_MISMATCH_HQ]);
    fputs("\"cigar\":{", f);
    for (int w = 0; w < 2; ++w) {
      fprintf(f, "%s\"%s\":{", w ? "," : "", w ? "read_two_cigar_ops" : "read_one_cigar_ops");
      int first = 1;
      for (int k = 0; k < 9; ++k) if (R->cigar_ops[w][k]) { fprintf(f, "%s\"%s\":%llu", first ? "" : ",", kOpNames[k], (unsigned long long)R->cigar_ops[w][k]); first = 0; }
      fputs("}", f);
    }
    /* general.rs:126-148 */
    double tot = (double)g[G_TOTAL];
    fputs("},\"summary\":{\"duplication_pct\":", f); jf(f, (double)g[G_DUPLICATE] / tot * 100.0);
    fputs(",\"mapped_pct\":", f); jf(f, (1.0 - (double)g[G_UNMAPPED] / tot) * 100.0);
    fputs(",\"mate_reference_sequence_id_mismatch_pct\":", f); jf(f, (double)g[G_MISMATCH] / tot * 100.0);
    fputs(",\"mate_reference_sequence_id_mismatch_hq_pct\":", f); jf(f, (double)g[G_MISMATCH_HQ] / tot * 100.0);
    fputs("}},\"features\":null,", f);
    /* gc_content.rs:102-122 */
    fputs("\"gc_content\":{\"histogram\":", f); jhist(f, R->gc.v, 100);
    fprintf(f, ",\"nucleobases\":{\"total_gc_count\":%llu,\"total_at_count\":%llu,\"total_other_count\":%llu},\"records\":{\"processed\":%llu,\"ignored_flags\":%llu,\"ignored_too_short\":%llu},",
            (unsigned long long)R->gc_nuc[0], (unsigned long long)R->gc_nuc[1], (unsigned long long)R->gc_nuc[2], (unsigned long long)R->gc_rec[0], (unsigned long long)R->gc_rec[1], (unsigned long long)R->gc_rec[2]);
    fputs("\"summary\":{\"gc_content_pct\":", f); jf(f, ((double)R->gc_nuc[0] / (double)(R->gc_nuc[0] + R->gc_nuc[1] + R->gc_nuc[2])) * 100.0);
    double rs = (double)(R->gc_rec[1] + R->gc_rec[2] + R->gc_rec[0]);
    fputs(",\"ignored_flags_pct\":", f); jf(f, ((double)R->gc_rec[1] / rs) * 100.0);
    fputs(",\"ignored_too_short_pct\":", f); jf(f, ((double)R->gc_rec[2] / rs) * 100.0);
    /* template_length.rs:89-100 */
    fputs("}},\"template_length\":{\"histogram\":", f); jhist(f, R->tlen.v, TLEN_CAP);
    fprintf(f, ",\"records\":{\"processed\":%llu,\"ignored\":%llu},\"summary\":{\"template_length_unknown_pct\":", (unsigned long long)R->tlen_processed, (unsigned long long)R->tlen_ignored);
    double den = (double)R->tlen_processed + (double)R->tlen_ignored;
    jf(f, ((double)R->tlen.v[0] / den) * 100.0);
    fputs(",\"template_length_out_of_range_pct\":", f); jf(f, ((double)R->tlen_ignored / den) * 100.0);
    fputs("}},\"quality_scores\":{\"scores\":{", f);
    for (uint64_t p = 0; p < R->qual_positions; ++p) { fprintf(f, "%s\"%llu\":", p ? "," : "", (unsigned long long)(p + 1)); jhist(f, R->qual + p * 94, MAX_SCORE); }
    fputs("}},", f);
  } else fputs("\"general\":null,\"features\":null,\"gc_content\":null,\"template_length\":null,\"quality_scores\":null,", f);
  if (R->do_coverage) {
    const char* keys[4] = {"mean_coverage", "median_coverage", "median_over_mean_coverage", NULL};
    fputs("\"coverage\":{", f);
    for (int k = 0; k < 3; ++k) {
      fprintf(f, "\"%s\":{", keys[k]); int first = 1;
      for (uint32_t c = 0; c < R->n_ref; ++c) if (R->touched[c]) { fprintf(f, "%s\"%s\":", first ? "" : ",", R->refs[c].name); jf(f, k == 0 ? R->cov_mean[c] : k == 1 ? R->cov_median[c] : R->cov_mom[c]); first = 0; }
      fputs("},", f);
    }
    fputs("\"mean_coverage_per_bin\":{", f); { int first = 1;
      for (uint32_t c = 0; c < R->n_ref; ++c) if (R->touched[c]) { fprintf(f, "%s\"%s\":[", first ? "" : ",", R->refs[c].name); for (uint64_t b = 0; b < R->cov_nbins[c]; ++b) { if (b) fputc(',', f); jf(f, R->cov_bins[c][b]); } fputs("]", f); first = 0; } }
    fprintf(f, "},\"ignored\":{\"nonsensical_records\":%llu,\"pileup_too_large_positions\":{", (unsigned long long)R->nonsensical); { int first = 1;
      for (uint32_t c = 0; c < R->n_ref; ++c) if (R->touched[c]) { fprintf(f, "%s\"%s\":%llu", first ? "" : ",", R->refs[c].name, (unsigned long long)R->cov_ignored[c]); first = 0; } }
    fputs("}},\"coverage_distribution\":", f); jhist(f, R->cov_dist.v, COV_HIST);
    fputs(",\"genome_covered_by\":{", f);
    for (int k = 0; k < 6; ++k) { fprintf(f, "%s\"%dx\":", k ? "," : "", 10 * (k + 1)); if (isnan(R->covered_by[k]) || isinf(R->covered_by[k])) fputs("null", f); else jnum(f, "%.9g", (double)R->covered_by[k]); }
    fputs("}},", f);
  } else fputs("\"coverage\":null,", f);
  fputs("\"edits\":null}\n", f);
  return 0;
}

/* --------------------------------------------------------- library surface */
void* oracle_run(const uint8_t* bam, size_t bam_len, const uint8_t* bai, size_t bai_len, uint64_t n_records, uint64_t gc_seed, int do_records, int do_coverage) {
  oracle_opts o = {n_records, gc_seed, do_records, do_coverage};
  results_t* R = malloc(sizeof *R);
  if (run(bam, bam_len, bai, bai_len, &o, R)) { free(R); return NULL; }
  return R;
}
void oracle_free(void* p) { free(p); } /* leaks the inner arrays: test infrastructure, short-lived processes */
int oracle_write_json(void* p, const char* path) { FILE* f = fopen(path, "w"); if (!f) FAIL("cannot open %s", path); write_json(p, f); fclose(f); return 0; }
uint32_t oracle_n_ref(void* p) { return ((results_t*)p)->n_ref; }
uint64_t oracle_pass1_records(void* p) { return ((results_t*)p)->pass1_records; }
uint64_t oracle_pass2_records(void* p) { return ((results_t*)p)->pass2_records; }
uint64_t oracle_inflated_bytes(void* p) { return ((results_t*)p)->inflated_bytes; }
/* integer getters use the same layouts as include/ngs_cuda.h's ngsq_get_* */
void oracle_get_general(void* p, uint64_t out[34]) { results_t* R = p; memcpy(out, R->general, 16 * 8); memcpy(out + 16, R->cigar_ops, 18 * 8); }
void oracle_get_tlen(void* p, uint64_t hist[1025], uint64_t* processed, uint64_t* ignored) { results_t* R = p; memcpy(hist, R->tlen.v, 1025 * 8); *processed = R->tlen_processed; *ignored = R->tlen_ignored; }
void oracle_get_gc(void* p, uint64_t hist[101], uint64_t nuc[3], uint64_t rec[3]) { results_t* R = p; memcpy(hist, R->gc.v, 101 * 8); memcpy(nuc, R->gc_nuc, 24); memcpy(rec, R->gc_rec, 24); }
uint64_t oracle_quality_positions(void* p) { return ((results_t*)p)->qual_positions; }
void oracle_get_quality(void* p, uint64_t* out) { results_t* R = p; memcpy(out, R->qual, R->qual_positions * 94 * 8); }
int oracle_cov_touched(void* p, uint32_t c) { results_t* R = p; return R->touched ? R->touched[c] : 0; }
uint64_t oracle_cov_nbins(void* p, uint32_t c) { return ((results_t*)p)->cov_nbins[c]; }
void oracle_get_cov_contig(void* p, uint32_t c, uint64_t hist[2049], uint64_t* ignored, uint64_t* bin_sums) { results_t* R = p; memcpy(hist, R->cov_hist_contig[c].v, 2049 * 8); *ignored = R->cov_ignored[c]; memcpy(bin_sums, R->cov_bin_sums[c], R->cov_nbins[c] * 8); }
void oracle_get_cov_dist(void* p, uint64_t hist[2049], uint64_t* nonsensical) { results_t* R = p; memcpy(hist, R->cov_dist.v, 2049 * 8); *nonsensical = R->nonsensical; }
void oracle_get_cov_floats(void* p, uint32_t c, double out[3]) { results_t* R = p; out[0] = R->cov_mean[c]; out[1] = R->cov_median[c]; out[2] = R->cov_mom[c]; }
void oracle_get_covered_by(void* p, float out[6]) { memcpy(out, ((results_t*)p)->covered_by, 24); }

/* Inflates every BGZF block of `bgzf` into `out` (zlib); returns bytes written or -1. */
int64_t oracle_inflate_all(const uint8_t* bgzf, size_t n, uint8_t* out, size_t cap) {
  g_err[0] = 0;
  bgzf_t* rd = malloc(sizeof *rd); bgzf_open(rd, bgzf, n); size_t w = 0; int rc;
  while ((rc = bgzf_load(rd)) == 1) { if (w + rd->cur_len > cap) { free(rd); FAIL("output too small"); } memcpy(out + w, rd->buf, rd->cur_len); w += rd->cur_len; rd->cur_off = rd->cur_len; }
  free(rd);
  return rc < 0 ? -1 : (int64_t)w;
}

/* ------------------------------------------------------------ Edits
 * Oracle FIRST for the next row of the scope table: no device path consumes this yet.
 * Restates src/utils/cigar.rs:6-23 (which ops consume what), src/utils/alignment.rs:13-126
 * (flatten + ReferenceRecordStepThrough::stepthrough / edits) and the Edits facet
 * src/qc/sequence_based/edits.rs:170-344 (supports every sequence name; setup loads the contig from
 * the FASTA; process; teardown = VAF histogram in f32; aggregate = the two means).
 * Pinned by the reference's own tests src/utils/alignment.rs:134-202 (tests/test_oracle_edits.py). */
static int edits_consumes_reference(uint32_t k) { return k == 0 || k == 2 || k == 3 || k == 7 || k == 8; } /* M D N = X */
static int edits_consumes_sequence(uint32_t k) { return k == 0 || k == 1 || k == 4 || k == 7 || k == 8; }  /* M I S = X */

/* Bases are compared as noodles `Base` values: one code per letter of "=ACMGRSVTWYHKDBN" (the BAM
 * 4-bit codes).  reference/record: arrays of those codes.  cigar: BAM ops (len << 4 | kind).
 * on_match(ctx, reference_ptr, is_edit) is called for every M position (edits.rs:274-292).
 * Returns 0, or: 1 "...consume a reference base, but no such base was found", 2 "...consume a record
 * base, but no such base was found", 3 "reference sequence was not fully consumed", 4 "record sequence
 * was not fully consumed" (alignment.rs:60-106), 6 invalid CIGAR op. */
typedef void (*edits_cb)(void* ctx, uint64_t reference_ptr, int is_edit);
static int edits_stepthrough(const uint8_t* reference, uint64_t n_reference, const uint8_t* record, uint64_t n_record,
                             const uint32_t* cigar, uint64_t n_ops, uint64_t* edits, edits_cb on_match, void* ctx) {
  uint64_t record_ptr = 0, reference_ptr = 0, e = 0;
  for (uint64_t o = 0; o < n_ops; ++o) {
    uint32_t op; memcpy(&op, (const uint8_t*)cigar + 4 * o, 4);
    uint32_t kind = op & 15, len = op >> 4;
    if (kind > 8) return 6;
    for (uint32_t t = 0; t < len; ++t) { /* the flattened CIGAR, one kind per step (alignment.rs:13-25,53) */
      int cr = edits_consumes_reference(kind), cs = edits_consumes_sequence(kind);
      int rb = -1, qb = -1; /* None */
      if (cr) { if (reference_ptr >= n_reference) return 1; rb = reference[reference_ptr]; }
      if (cs) { if (record_ptr >= n_record) return 2; qb = record[record_ptr]; }
      if (kind == 0) { /* Kind::Match only: "=" and "X" are not counted (edits.rs:274) */
        int is_edit = rb != qb;
        e += is_edit;
        if (on_match) on_match(ctx, reference_ptr, is_edit);
      }
      if (cr) reference_ptr++;
      if (cs) record_ptr++;
    }
  }
  if (reference_ptr != n_reference) return 3;
  if (record_ptr != n_record) return 4;
  *edits = e;
  return 0;
}
int oracle_stepthrough_edits(const uint8_t* reference, uint64_t n_reference, const uint8_t* record, uint64_t n_record,
                             const uint32_t* cigar, uint64_t n_ops, uint64_t* edits) {
  return edits_stepthrough(reference, n_reference, record, n_record, cigar, n_ops, edits, NULL, NULL);
}

/* noodles Base::try_from(char): the sixteen upper-case letters of the BAM code table, anything else
 * (lower case included) is an error and aborts the run (edits.rs:259-265).  Returns the 4-bit code or -1. */
static int edits_base_code(uint8_t ch) {
  static const char tab[] = "=ACMGRSVTWYHKDBN";
  for (int i = 0; i < 16; ++i) if ((uint8_t)tab[i] == ch) return i;
  return -1;
}

typedef struct {
  hist_t read_one, read_two, vaf;      /* Histogram::default() = 0..=512 (histogram.rs:472-481), VAF 0..=100 */
  double mean_read_one, mean_read_two; /* aggregate (edits.rs:336-344) */
  uint64_t records;                    /* records that reached the step-through */
} edits_t;
typedef struct { uint64_t* refs; uint64_t* alts; uint64_t start, cap; int overflow; } edits_pos_ctx;
static void edits_on_match(void* c, uint64_t reference_ptr, int is_edit) {
  edits_pos_ctx* x = c;
  uint64_t p = x->start + reference_ptr; /* 1-based reference position (edits.rs:276-278) */
  if (p > x->cap) { x->overflow = 1; return; } /* increment(position).unwrap() beyond the header's length panics (edits.rs:281-290) */
  if (is_edit) x->alts[p]++; else x->refs[p]++;
}

/* fasta: text of a FASTA file whose record names (first word after '>') are looked up per contig. */
static int edits_find_contig(const char* fasta, size_t n, const char* name, uint8_t** codes, uint64_t* len) {
  size_t i = 0, ln = strlen(name);
  while (i < n) {
    if (fasta[i] != '>') { while (i < n && fasta[i] != '\n') ++i; ++i; continue; }
    size_t b = i + 1, e = b;
    while (e < n && fasta[e] != '\n' && fasta[e] != ' ' && fasta[e] != '\t') ++e;
    size_t eol = e; while (eol < n && fasta[eol] != '\n') ++eol;
    int hit = (e - b == ln) && !memcmp(fasta + b, name, ln);
    i = eol + 1;
    size_t s = i;
    while (i < n && fasta[i] != '>') { while (i < n && fasta[i] != '\n') ++i; ++i; }
    if (!hit) continue;
    uint8_t* out = malloc(i - s + 1); uint64_t k = 0;
    for (size_t q = s; q < i && q < n; ++q) {
      uint8_t ch = (uint8_t)fasta[q];
      if (ch == '\n' || ch == '\r') continue;
      out[k++] = ch; /* kept as characters: conversion to Base happens per record slice (edits.rs:259-265) */
    }
    *codes = out; *len = k;
    return 0;
  }
  return -1;
}

/* n_records: `-n` (0 = all).  Pass 2 keeps ONE counter over all sequences, incremented for every record a sequence's query yields —
 * after the facets saw it, whether or not they skipped it — and reaching the limit only leaves the current sequence's loop
 * (command.rs:375-388, display.rs:43-65). */
void* oracle_edits_run_n(const uint8_t* bam, size_t bam_len, const uint8_t* bai, size_t bai_len, const char* fasta, size_t fasta_len, uint64_t n_records) {
  g_err[0] = 0;
  uint64_t counter = 0;
  edits_t* E = calloc(1, sizeof *E);
  hist_init(&E->read_one, 512); hist_init(&E->read_two, 512); hist_init(&E->vaf, 100);
  bgzf_t* rd = malloc(sizeof *rd); recbuf_t rb = {malloc(1 << 16), 1 << 16}; rec_t rec;
  ref_t* refs; uint32_t n_ref;
  bgzf_open(rd, bam, bam_len);
  if (read_header(rd, &refs, &n_ref)) return NULL;
  bai_t B; if (!bai || bai_parse(bai, bai_len, &B)) { if (!g_err[0]) snprintf(g_err, sizeof g_err, "missing BAM index"); return NULL; }
  for (uint32_t c = 0; c < n_ref; ++c) { /* pass 2 driver, command.rs:356-397; supports_sequence_name is always true (edits.rs:178-180) */
    uint64_t L = refs[c].len;
    uint8_t* seq; uint64_t seq_len;
    if (edits_find_contig(fasta, fasta_len, refs[c].name, &seq, &seq_len)) { snprintf(g_err, sizeof g_err, "sequence %s not found in reference FASTA.", refs[c].name); return NULL; }
    uint64_t* refs_pp = calloc(L + 1, 8); uint64_t* alts_pp = calloc(L + 1, 8); /* zero_based_with_capacity(seq_length) */
    chunk_t* ch; size_t nch = bai_query(&B, c, 0, (int64_t)L, &ch);
    int stop = 0;
    for (size_t k = 0; k < nch && !stop; ++k) {
      int src = bgzf_seek(rd, ch[k].beg); if (src < 0) return NULL; if (src == 0) break;
      while (bgzf_tell(rd) < ch[k].end) {
        int rc = read_record(rd, &rb, &rec, n_ref); if (rc < 0) return NULL; if (rc == 0) break;
        if (rec.ref_id != (int32_t)c || rec.pos < 0) continue;
        uint64_t start = (uint64_t)rec.pos + 1, end = start + rec.span - 1;
        if (end == 0) continue;
        if (!(start <= L && end >= 1)) continue;
        if (!(rec.flag & (0x4 | 0x400))) { /* unmapped or duplicate: the facet returns at once (edits.rs:227-229) */
        { const char* nm = (const char*)rb.buf + 32; /* read name "*" = missing: the facet bails (edits.rs:233-236) */
          if (rec.l_name <= 1 || (rec.l_name == 2 && nm[0] == '*')) { snprintf(g_err, sizeof g_err, "Could not parse read name"); return NULL; } }
        /* reference slice start .. start + span, 1-based, end exclusive (edits.rs:241-243,259-262): out of the
         * FASTA sequence -> the reference unwraps a None and panics */
        if (start - 1 + rec.span > seq_len) { snprintf(g_err, sizeof g_err, "record reaches past the end of the reference sequence"); return NULL; }
        uint8_t* rcodes = malloc(rec.span + 1);
        for (uint32_t t = 0; t < rec.span; ++t) {
          int code = edits_base_code(seq[start - 1 + t]);
          if (code < 0) { snprintf(g_err, sizeof g_err, "invalid base in the reference sequence"); return NULL; }
          rcodes[t] = (uint8_t)code;
        }
        uint8_t* qcodes = malloc(rec.l_seq + 1);
        for (uint32_t t = 0; t < rec.l_seq; ++t) qcodes[t] = (t & 1) ? (rec.seq[t >> 1] & 15) : (rec.seq[t >> 1] >> 4);
        edits_pos_ctx ctx = {refs_pp, alts_pp, start, L, 0};
        uint64_t edits = 0;
        int st = edits_stepthrough(rcodes, rec.span, qcodes, rec.l_seq, rec.cigar, rec.n_cigar, &edits, edits_on_match, &ctx);
        free(rcodes); free(qcodes);
        if (st) { snprintf(g_err, sizeof g_err, "step-through failed (%d)", st); return NULL; }
        if (ctx.overflow) { snprintf(g_err, sizeof g_err, "matched position beyond the sequence length of the header"); return NULL; }
        /* increment(edits).unwrap(): more than 512 edits panics the reference (edits.rs:296-300) */
        if (hist_inc_by((rec.flag & 0x40) ? &E->read_one : &E->read_two, edits, 1)) { snprintf(g_err, sizeof g_err, "more than 512 edits in one read"); return NULL; }
        E->records++;
        }
        ++counter;
        if (n_records && counter >= n_records) { stop = 1; break; }
      }
    }
    free(ch);
    for (uint64_t i = 0; i <= L; ++i) { /* teardown, edits.rs:318-334 */
      uint64_t r = refs_pp[i], a = alts_pp[i], t = r + a;
      if (!t) continue;
      float vaf = (float)a / (float)t;
      if (hist_inc_by(&E->vaf, (uint64_t)(vaf * 100.0f), 1)) return NULL;
    }
    free(refs_pp); free(alts_pp); free(seq);
  }
  E->mean_read_one = hist_mean(&E->read_one);
  E->mean_read_two = hist_mean(&E->read_two);
  free(rd); free(rb.buf);
  return E;
}
void* oracle_edits_run(const uint8_t* bam, size_t bam_len, const uint8_t* bai, size_t bai_len, const char* fasta, size_t fasta_len) {
  return oracle_edits_run_n(bam, bam_len, bai, bai_len, fasta, fasta_len, 0);
}
void oracle_edits_get(void* p, uint64_t read_one[513], uint64_t read_two[513], uint64_t vaf[101], double means[2], uint64_t* records) {
  edits_t* E = p;
  memcpy(read_one, E->read_one.v, 513 * 8); memcpy(read_two, E->read_two.v, 513 * 8); memcpy(vaf, E->vaf.v, 101 * 8);
  means[0] = E->mean_read_one; means[1] = E->mean_read_two; *records = E->records;
}


/* ------------------------------------------------------------ Genomic Features facet
 * Oracle first, like Edits: no device path consumes this yet.  Restates
 *   src/qc/record_based/features.rs:115-242 (process), 244-262 (summarize), 270-355 (try_from),
 *   src/qc/record_based/features/utils.rs:33-43 (strand parse), src/utils/formats/gff.rs (plain-text GFF only here).
 * Third-party behaviour restated from the published crates (absent from /root/reference, Cargo.lock pins):
 *   rust-lapper 1.1.0: Lapper::new sorts intervals by (start, stop); find(start, stop) yields, in that order, every
 *     interval with iv.start < stop && iv.stop > start (half-open);
 *   noodles-gff (noodles 0.34.0): records() skips comments and directives, ends at ##FASTA, and the facet unwraps every
 *     record, so a malformed line aborts; Display of a strand is "+", "-", "." or "?".
 * Reference quirks kept: a feature's GFF end (inclusive) is used as an exclusive stop, so its last base does not count;
 * the query is [start, start + span + 1) with the 1-based start, two bases past the last aligned base. */
typedef struct { uint64_t start, stop; uint8_t is5, is3, iscds, isgene, isexon; } feat_iv;
typedef struct { feat_iv* v; size_t n, cap; } feat_list;
typedef struct {
  uint64_t utr5, utr3, cds, intergenic, exonic, intronic, processed, ignored_flags, ignored_nonprimary;
  double ignored_flags_pct, ignored_nonprimary_pct;
} features_t;
static void feat_push(feat_list* l, feat_iv iv) { if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 64; l->v = realloc(l->v, l->cap * sizeof *l->v); } l->v[l->n++] = iv; }
static int feat_cmp(const void* a, const void* b) {
  const feat_iv *x = a, *y = b;
  if (x->start != y->start) return x->start < y->start ? -1 : 1;
  if (x->stop != y->stop) return x->stop < y->stop ? -1 : 1;
  return 0;
}
static int feat_parse_pos(const char* s, size_t n, uint64_t* out) { /* noodles Position: decimal, >= 1 */
  if (!n || n > 18) return -1;
  uint64_t v = 0;
  for (size_t i = 0; i < n; ++i) { if (s[i] < '0' || s[i] > '9') return -1; v = v * 10 + (uint64_t)(s[i] - '0'); }
  if (!v) return -1;
  *out = v; return 0;
}
static int feat_eq(const char* s, size_t n, const char* name) { return strlen(name) == n && !memcmp(s, name, n); }

/* names: five_prime_utr, three_prime_utr, coding_sequence, exon, gene (command.rs:78-101 defaults
 * "five_prime_UTR", "three_prime_UTR", "CDS", "exon", "gene") */
void* oracle_features_run(const uint8_t* bam, size_t bam_len, const char* gff, size_t gff_len, const char* const names[5], uint64_t n_records) {
  g_err[0] = 0;
  features_t* F = calloc(1, sizeof *F);
  bgzf_t* rd = malloc(sizeof *rd); recbuf_t rb = {malloc(1 << 16), 1 << 16}; rec_t rec;
  ref_t* refs; uint32_t n_ref;
  bgzf_open(rd, bam, bam_len);
  if (read_header(rd, &refs, &n_ref)) return NULL;
  /* try_from: one (exonic translation, gene region) pair of lists per primary contig of the header */
  feat_list* utr = calloc(n_ref ? n_ref : 1, sizeof *utr); feat_list* gene = calloc(n_ref ? n_ref : 1, sizeof *gene);
  size_t o = 0;
  while (o < gff_len) {
    size_t e = o; while (e < gff_len && gff[e] != '\n') ++e;
    size_t n = e - o; const char* ln = gff + o; o = e + 1;
    if (n && ln[n - 1] == '\r') --n;
    if (n >= 7 && !memcmp(ln, "##FASTA", 7)) break;
    if (n && ln[0] == '#') continue;
    const char* f[9]; size_t fl[9]; int nf = 0; size_t b = 0;
    for (size_t i = 0; i <= n && nf < 9; ++i) if (i == n || ln[i] == '\t') { f[nf] = ln + b; fl[nf] = i - b; ++nf; b = i + 1; }
    uint64_t start, stop;
    if (nf < 9 || b <= n || feat_parse_pos(f[3], fl[3], &start) || feat_parse_pos(f[4], fl[4], &stop)) { snprintf(g_err, sizeof g_err, "invalid GFF record"); return NULL; }
    uint32_t c = n_ref;
    for (uint32_t i = 0; i < n_ref; ++i) if (refs[i].primary && feat_eq(f[0], fl[0], refs[i].name)) { c = i; break; }
    if (c == n_ref) continue; /* only primary-assembly sequence names are tabulated (features.rs:297-301) */
    if (!(fl[6] == 1 && (f[6][0] == '+' || f[6][0] == '-'))) { /* utils.rs:36-42, parsed before the type is looked at */
      snprintf(g_err, sizeof g_err, "attempted to parse strand from value: %.*s", (int)fl[6], f[6]); return NULL; }
    feat_iv iv = {start, stop, (uint8_t)feat_eq(f[2], fl[2], names[0]), (uint8_t)feat_eq(f[2], fl[2], names[1]), (uint8_t)feat_eq(f[2], fl[2], names[2]),
                  (uint8_t)feat_eq(f[2], fl[2], names[4]), (uint8_t)feat_eq(f[2], fl[2], names[3])};
    if (iv.is5 || iv.is3 || iv.iscds) feat_push(&utr[c], iv);              /* features.rs:322-326 */
    else if (iv.isexon || iv.isgene) feat_push(&gene[c], iv);             /* features.rs:327-331 */
  }
  for (uint32_t c = 0; c < n_ref; ++c) { qsort(utr[c].v, utr[c].n, sizeof(feat_iv), feat_cmp); qsort(gene[c].v, gene[c].n, sizeof(feat_iv), feat_cmp); }
  uint64_t seen = 0;
  for (;;) { /* pass 1 driver, command.rs:291-316 */
    int rc = read_record(rd, &rb, &rec, n_ref); if (rc < 0) return NULL; if (rc == 0) break;
    const char* nm = (const char*)rb.buf + 32;
    if (rec.l_name <= 1 || (rec.l_name == 2 && nm[0] == '*')) { snprintf(g_err, sizeof g_err, "Could not parse read name"); return NULL; }
    if (rec.flag & 0x4) F->ignored_flags++;
    else {
      if (rec.ref_id < 0) { snprintf(g_err, sizeof g_err, "Could not parse reference sequence id for read"); return NULL; }
      if (!refs[rec.ref_id].primary) F->ignored_nonprimary++;
      else {
        if (rec.pos < 0) { snprintf(g_err, sizeof g_err, "Could not parse record's start position."); return NULL; }
        const uint64_t start = (uint64_t)rec.pos + 1, end = start + rec.span, qstop = end + 1;
        int c5 = 0, c3 = 0, cc = 0;
        const feat_list* U = &utr[rec.ref_id];
        for (size_t i = 0; i < U->n && U->v[i].start < qstop; ++i) { /* features.rs:185-209, in Lapper order */
          const feat_iv* iv = &U->v[i];
          if (!(iv->stop > start)) continue;
          if (!c5 && iv->is5) { c5 = 1; F->utr5++; }
          else if (!c3 && iv->is3) { c3 = 1; F->utr3++; }
          else if (!cc && iv->iscds) { cc = 1; F->cds++; }
        }
        int has_gene = 0, has_exon = 0;
        const feat_list* G = &gene[rec.ref_id];
        for (size_t i = 0; i < G->n && G->v[i].start < qstop; ++i) { /* features.rs:212-226 */
          const feat_iv* iv = &G->v[i];
          if (!(iv->stop > start)) continue;
          if (iv->isgene) has_gene = 1; else if (iv->isexon) has_exon = 1;
          if (has_gene && has_exon) break;
        }
        if (has_gene) { if (has_exon) F->exonic++; else F->intronic++; } else F->intergenic++; /* features.rs:228-236 */
        F->processed++;
      }
    }
    if (n_records && ++seen >= n_records) break;
  }
  const double tot = (double)(F->ignored_flags + F->ignored_nonprimary + F->processed); /* features.rs:244-259: 0/0 = NaN -> JSON null */
  F->ignored_flags_pct = (double)F->ignored_flags / tot * 100.0;
  F->ignored_nonprimary_pct = (double)F->ignored_nonprimary / tot * 100.0;
  free(rd); free(rb.buf);
  return F;
}
/* counts: utr_five_prime, utr_three_prime, coding_sequence, intergenic, exonic, intronic, processed, ignored_flags,
 * ignored_nonprimary_chromosome (field order of features/metrics.rs) */
void oracle_features_get(void* p, uint64_t counts[9], double pct[2]) {
  features_t* F = p;
  counts[0] = F->utr5; counts[1] = F->utr3; counts[2] = F->cds; counts[3] = F->intergenic; counts[4] = F->exonic; counts[5] = F->intronic;
  counts[6] = F->processed; counts[7] = F->ignored_flags; counts[8] = F->ignored_nonprimary;
  pct[0] = F->ignored_flags_pct; pct[1] = F->ignored_nonprimary_pct;
}

#ifdef ORACLE_MAIN
static uint8_t* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb"); if (!f) return NULL;
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  uint8_t* b = malloc(sz ? sz : 1); if (fread(b, 1, sz, f) != (size_t)sz) { fclose(f); return NULL; } fclose(f); *n = sz; return b;
}
/* usage: ngsqc_oracle <in.bam> <out.json> [-n N] [--only records|coverage] [--gc-seed S] */
int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <in.bam> <out.json> [-n N] [--only records|coverage] [--gc-seed S]\n", argv[0]); return 2; }
  uint64_t n = 0, seed = 0; int dr = 1, dc = 1;
  for (int i = 3; i < argc; ++i) {
    if (!strcmp(argv[i], "-n") && i + 1 < argc) n = strtoull(argv[++i], NULL, 10);
    else if (!strcmp(argv[i], "--gc-seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 0);
    else if (!strcmp(argv[i], "--only") && i + 1 < argc) { ++i; if (!strcmp(argv[i], "records")) dc = 0; else dr = 0; }
  }
  size_t bl, il = 0; uint8_t* bam = slurp(argv[1], &bl); if (!bam) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
  char ip[4096]; snprintf(ip, sizeof ip, "%s.bai", argv[1]); /* utils/pathbuf.rs:59-75: x.bam -> x.bam.bai */
  uint8_t* bai = slurp(ip, &il);
  void* R = oracle_run(bam, bl, bai, il, n, seed, dr, dc);
  if (!R) { fprintf(stderr, "error: %s\n", oracle_last_error()); return 1; }
  if (oracle_write_json(R, argv[2])) { fprintf(stderr, "error: %s\n", oracle_last_error()); return 1; }
  fprintf(stderr, "pass1=%llu pass2=%llu inflated=%llu\n", (unsigned long long)oracle_pass1_records(R), (unsigned long long)oracle_pass2_records(R), (unsigned long long)oracle_inflated_bytes(R));
  return 0;
}
#endif
