"""A second, independent restatement of the reference's facet rules (pure Python + numpy, written from
SURVEY App. C / the cited reference lines, sharing no code with oracle/ngsqc_oracle.c) must agree with the
oracle integer for integer on generator-written BAMs — not only on hand-built micro-BAMs.  Two
restatements agreeing is what pins the oracle in the absence of a runnable reference."""
import struct
import zlib

import numpy as np
import pytest

from helpers import oracle_ints

MASK64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def _records(raw: bytes):
    """Yields (virtual offset, record bytes without block_size) and returns via StopIteration nothing; header first."""
    blocks, off = [], 0
    stream = bytearray()
    starts = []  # (stream offset, coffset) per block
    while off < len(raw):
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        data = zlib.decompress(raw[off + 18:off + bsize - 8], -15)
        if data:
            starts.append((len(stream), off))
            stream += data
        off += bsize
    s = bytes(stream)
    l_text = int.from_bytes(s[4:8], "little")
    p = 8 + l_text
    n_ref = int.from_bytes(s[p:p + 4], "little")
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = int.from_bytes(s[p:p + 4], "little")
        refs.append((s[p + 4:p + 4 + ln - 1].decode(), int.from_bytes(s[p + 4 + ln:p + 8 + ln], "little")))
        p += 8 + ln
    so = np.array([a for a, _ in starts])
    co = [c for _, c in starts]
    recs = []
    while p < len(s):
        bs = int.from_bytes(s[p:p + 4], "little")
        k = int(np.searchsorted(so, p, side="right")) - 1
        recs.append(((co[k] << 16) | (p - int(so[k])), s[p + 4:p + 4 + bs]))
        p += 4 + bs
    return refs, recs


def _python_facets(raw: bytes, gc_seed: int, primary):
    refs, recs = _records(raw)
    general = np.zeros(34, np.uint64)
    tlen_hist, tl = np.zeros(1025, np.uint64), [0, 0]
    gc_hist, nuc, gcrec = np.zeros(101, np.uint64), [0, 0, 0], [0, 0, 0]
    qual = {}
    diffs = {c: None for c in range(len(refs))}  # contig -> {position: depth delta}, sparse (contigs are Gbp long)
    nonsensical = 0
    for voff, b in recs:
        ref, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", b, 0)
        o = 32 + l_name
        cig = struct.unpack_from(f"<{n_cig}I", b, o)
        o += 4 * n_cig
        seq = b[o:o + (l_seq + 1) // 2]
        q = b[o + (l_seq + 1) // 2:o + (l_seq + 1) // 2 + l_seq]
        # General (general.rs:31-124)
        g = general
        g[0] += 1
        if flag & 0x4: g[1] += 1
        if flag & 0x400: g[2] += 1
        if flag & 0x100: g[4] += 1
        elif flag & 0x800: g[5] += 1
        else:
            g[3] += 1
            if not flag & 0x4: g[6] += 1
            if flag & 0x400: g[7] += 1
            if flag & 0x1:
                g[8] += 1
                if flag & 0x40: g[9] += 1
                if flag & 0x80: g[10] += 1
                if not flag & 0x4:
                    if flag & 0x2: g[11] += 1
                    if flag & 0x8: g[12] += 1
                    else:
                        g[13] += 1
                        if ref != nref:
                            g[14] += 1
                            if mapq >= 5: g[15] += 1
        base = 16 if flag & 0x40 else 25
        span = 0
        for op in cig:
            g[base + (op & 15)] += 1
            if (op & 15) in (0, 2, 3, 7, 8): span += op >> 4
        # Template length (template_length.rs:79-87)
        if 0 <= tlen <= 1024:
            tlen_hist[tlen] += 1; tl[0] += 1
        else:
            tl[1] += 1
        # GC content (gc_content.rs:38-100) with the shared window policy
        if flag & (0x400 | 0x100): gcrec[1] += 1
        elif l_seq < 100: gcrec[2] += 1
        else:
            offset = ((_splitmix64(gc_seed ^ voff) >> 32) * (l_seq - 100)) >> 32 if l_seq > 100 else 0
            gc = at = 0
            for i in range(offset, offset + 100):
                c = (seq[i >> 1] >> 4) if not i & 1 else (seq[i >> 1] & 15)
                if c in (2, 4): gc += 1
                elif c in (1, 8): at += 1
            gc_hist[gc] += 1; gcrec[0] += 1
            nuc[0] += gc; nuc[1] += at; nuc[2] += 100 - gc - at
        # Quality (quality_scores.rs:37-49)
        if l_seq and any(x != 0xFF for x in q):
            for i, x in enumerate(q):
                qual.setdefault(i, np.zeros(94, np.uint64))[x] += 1
        # Coverage (coverage.rs:148-180 behind the query filter)
        if ref >= 0 and pos >= 0 and primary(refs[ref][0]):
            L = refs[ref][1]
            start, end = pos + 1, pos + span
            if end != 0 and start <= L and end >= 1:
                if diffs[ref] is None: diffs[ref] = {}
                if span:
                    d = diffs[ref]
                    d[start] = d.get(start, 0) + 1
                    d[min(end, L) + 1] = d.get(min(end, L) + 1, 0) - 1
                    nonsensical += max(0, end - L)
    cov = {}
    for c, d in diffs.items():
        if d is None: continue
        L = refs[c][1]
        # depth is piecewise constant: value dep[i] on positions [brk[i], brk[i+1]) of 0..L
        brk = np.array(sorted(set(d) | {0, L + 1}), np.int64)
        brk = brk[brk <= L + 1]
        dep = np.cumsum([d.get(int(x), 0) for x in brk[:-1]]).astype(np.int64)
        run = np.diff(brk)
        hist = np.zeros(2049, np.uint64)
        small = dep <= 2048
        np.add.at(hist, dep[small], run[small].astype(np.uint64))
        too_large = int(run[~small].sum())
        before = np.concatenate([[0], np.cumsum(dep * run)])  # sum of depth over positions < brk[i]

        def upto(x):  # sum of depth over positions 0..x
            i = int(np.searchsorted(brk, x, side="right")) - 1
            return int(before[i] + dep[i] * (x - brk[i] + 1))
        # bins: {0}, then (50000(k-1), 50000k], then the tail (coverage.rs:206-230)
        edges = list(range(0, L + 1, 50000))
        sums = [upto(0)] + [upto(b) - upto(a) for a, b in zip(edges[:-1], edges[1:])]
        if L % 50000: sums.append(upto(L) - upto(edges[-1]))
        cov[c] = {"hist": hist, "too_large": too_large, "bin_sums": np.array(sums, np.uint64)}
    n_pos = max(qual) + 1 if qual else 0
    qarr = np.zeros((n_pos, 94), np.uint64)
    for i, h in qual.items(): qarr[i] = h
    return dict(general=general, tlen_hist=tlen_hist, tlen_processed=tl[0], tlen_ignored=tl[1], gc_hist=gc_hist,
                gc_nuc=np.array(nuc, np.uint64), gc_rec=np.array(gcrec, np.uint64), quality=qarr, coverage=cov, nonsensical=nonsensical)


@pytest.mark.parametrize("shape,n,seed", [(0, 12000, 3), (1, 6000, 5), (3, 6000, 8), (2, 120, 1)])
def test_python_restatement_agrees_with_the_oracle(shape, n, seed):
    from ngs_b200 import ffi, formats
    from helpers import assert_same_ints
    bam, bai, _ = ffi.synth_bam(shape, n, level=6)
    want = oracle_ints(bam, bai, gc_seed=seed)
    got = _python_facets(bam.tobytes(), seed, formats.is_primary)
    assert_same_ints(got, want)


def test_restatements_agree_on_an_adversarial_bam():
    """Deep pile-up (> 2048, coverage.rs:186-199), reads hanging over the contig end (nonsensical records),
    zero-span CIGARs, a non-primary contig, odd sequence lengths, missing qualities, TLEN out of range."""
    import random
    from bamutil import as_u8, rec, write_bam
    from ngs_b200 import formats
    rng = random.Random(4)
    refs = [("chr1", 120000), ("chrUn_KI270302v1", 5000), ("chrM", 700)]
    rs = []
    seq = lambda n: "".join(rng.choice("ACGTN") for _ in range(n))
    for i in range(2200):                                    # 2200-deep column at chr1:1001-1050
        n = 101 + (i % 3)
        rs.append(rec(name=f"d{i}", flag=[0x41, 0x81, 0x400, 0x100, 0x800, 0x63][i % 6], ref=0, pos=1000, mapq=i % 60,
                      cigar=f"50M{n - 50}S", next_ref=[0, 2][i % 2], next_pos=5, tlen=[0, 300, 1024, 1025, -7][i % 5],
                      seq=seq(n), qual=[rng.randrange(94) for _ in range(n)]))
    for i in range(40):
        rs.append(rec(name=f"z{i}", flag=0, ref=0, pos=60000 + i, cigar="30S71I", seq=seq(101)))          # span 0, no qualities
    for i in range(30):
        rs.append(rec(name=f"e{i}", flag=0, ref=0, pos=119990, cigar="20M5D10N66M15H", seq=seq(86), qual=[i] * 86))  # overhang
    for i in range(25):
        rs.append(rec(name=f"u{i}", flag=0, ref=1, pos=10 * i, cigar="100M", seq=seq(100), qual=[40] * 100))  # not a primary contig
    for i in range(35):
        rs.append(rec(name=f"m{i}", flag=0x10, ref=2, pos=650, cigar="10=5X90M", seq=seq(105), qual=[93] * 105))  # chrM overhang
    for i in range(10):
        rs.append(rec(name=f"n{i}", flag=0x4 | 0x1 | 0x8, seq=seq(99 + i), qual=[1] * (99 + i)))           # no coordinate
    bam, bai = write_bam(refs, rs, block_payload=[4000, 65280, 900])
    b, x = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, x, gc_seed=11)
    got = _python_facets(bam, 11, formats.is_primary)
    assert want["coverage"][0]["too_large"] == 50 and want["nonsensical"] > 0
    from helpers import assert_same_ints
    assert_same_ints(got, want)
