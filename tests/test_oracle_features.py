"""Oracle first for SURVEY 8(f) rank 3, the Genomic Features facet (oracle/ngsqc_oracle.c, "Genomic Features facet"):
hand-derived answers for every rule of features.rs:115-242 / 270-355 on micro-BAMs, and agreement with an independent,
order-free Python formulation on a generator-written RNA-seq BAM with a random gene model.  No device path yet."""
import ctypes as C
import random
import struct
import zlib

import numpy as np
import pytest

from bamutil import as_u8, rec, write_bam
from helpers import oracle_lib

NAMES = ("five_prime_UTR", "three_prime_UTR", "CDS", "exon", "gene")
KEYS = ["utr5", "utr3", "cds", "intergenic", "exonic", "intronic", "processed", "ignored_flags", "ignored_nonprimary"]
REFS = [("chr1", 100000), ("chr2", 50000), ("chrM", 16569), ("chrUn_KI270302v1", 2274)]


def features(bam: bytes, gff: str, names=NAMES, n_records=0):
    lib = oracle_lib()
    lib.oracle_features_run.restype = C.c_void_p
    lib.oracle_features_run.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_uint64]
    lib.oracle_features_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    b = as_u8(bam)
    g = gff.encode()
    arr = (C.c_char_p * 5)(*[n.encode() for n in names])
    h = lib.oracle_features_run(b.ctypes.data, b.size, g, len(g), arr, n_records)
    if not h:
        raise RuntimeError(lib.oracle_last_error().decode())
    counts, pct = np.zeros(9, np.uint64), np.zeros(2, np.float64)
    lib.oracle_features_get(h, counts.ctypes.data, pct.ctypes.data)
    out = dict(zip(KEYS, (int(x) for x in counts)))
    out["pct"] = (float(pct[0]), float(pct[1]))
    return out


def gff_line(seq, ty, start, end, strand="+"):
    return f"{seq}\ttest\t{ty}\t{start}\t{end}\t.\t{strand}\t.\tID=x\n"


def one(pos, cigar="100M", ref=0, flag=0, name="r"):
    return rec(name=name, flag=flag, ref=ref, pos=pos, cigar=cigar, seq="A" * 100)


def run_one(gff, *records, **kw):
    bam, _ = write_bam(REFS, list(records))
    return features(bam, gff, **kw)


def test_gene_regions_decision_tree():
    gff = "##gff-version 3\n# a comment\n" + gff_line("chr1", "gene", 1000, 5000) + gff_line("chr1", "exon", 1000, 1500) + gff_line("chr1", "exon", 4000, 5000, "-")
    got = run_one(gff, one(1099), one(1999), one(8999), one(99, ref=1))
    # 1100-1199 lies in an exon of the gene; 2000-2099 only in the gene; 9000-9099 and chr2 (no features at all) in nothing
    assert (got["exonic"], got["intronic"], got["intergenic"], got["processed"]) == (1, 1, 2, 4)
    assert (got["utr5"], got["utr3"], got["cds"]) == (0, 0, 0)
    # an exon outside every gene still leaves the read intergenic (has_gene is false, features.rs:228-236)
    got = run_one(gff_line("chr1", "exon", 1000, 1500), one(1099))
    assert (got["exonic"], got["intronic"], got["intergenic"]) == (0, 0, 1)


def test_interval_edges_follow_the_reference_quirks():
    gff = gff_line("chr1", "gene", 2000, 3000)
    # query = [pos+1, pos+1+span+1): a 100M read at 1-based 1899 reaches 1-based 1998, the query stops before 2000
    assert run_one(gff, one(1898))["intergenic"] == 1
    # ... one base further the query's exclusive stop is 2001 > 2000: counted although the read ends at 1999
    assert run_one(gff, one(1899))["intronic"] == 1
    # the feature's inclusive end 3000 is used as an exclusive stop: a read starting at 1-based 3000 misses it
    assert run_one(gff, one(2999))["intergenic"] == 1
    assert run_one(gff, one(2998))["intronic"] == 1
    # deletions and skips stretch the span, insertions and clips do not (utils/cigar.rs:6-11)
    assert run_one(gff, one(1800, cigar="50M150N50M"))["intronic"] == 1
    assert run_one(gff, one(1800, cigar="50M150I"))["intergenic"] == 1


def test_exonic_translation_chain_counts_each_kind_once():
    gff = (gff_line("chr1", "five_prime_UTR", 100, 200) + gff_line("chr1", "five_prime_UTR", 150, 260) + gff_line("chr1", "CDS", 180, 400)
           + gff_line("chr1", "three_prime_UTR", 5000, 5100))
    got = run_one(gff, one(149), one(4990), one(20000))
    assert (got["utr5"], got["cds"], got["utr3"], got["processed"]) == (1, 1, 1, 3)
    # exonic-translation features never feed the gene-region tally (they are filed in the other Lapper, features.rs:322-331)
    assert got["intergenic"] == 3


def test_coinciding_feature_names_fill_the_chain_in_order():
    """GTF-style models call both UTRs "UTR": the first overlapping one counts as 5', the second as 3' (features.rs:188-207)."""
    names = ("UTR", "UTR", "CDS", "exon", "gene")
    gff = gff_line("chr1", "UTR", 100, 200) + gff_line("chr1", "UTR", 150, 260) + gff_line("chr1", "UTR", 170, 270)
    got = run_one(gff, one(149), one(99, cigar="10M90S"), names=names)
    assert (got["utr5"], got["utr3"], got["cds"]) == (2, 1, 0)
    # gene == exon name: the gene test wins, the read is never exonic (features.rs:215-219)
    got = run_one(gff_line("chr1", "gene", 100, 900), one(149), names=("a", "b", "c", "gene", "gene"))
    assert (got["exonic"], got["intronic"]) == (0, 1)


def test_record_filters_and_summary():
    gff = gff_line("chr1", "gene", 1, 90000)
    got = run_one(gff, one(10), one(20, flag=0x4), one(30, ref=2), one(5, ref=3), rec(name="u", flag=0x4 | 0x1, seq="A" * 100))
    # unmapped twice (placed and unplaced); chrM is not primary; chrUn_* is, and has no features -> intergenic
    assert (got["processed"], got["ignored_flags"], got["ignored_nonprimary"]) == (2, 2, 1)
    assert (got["intronic"], got["intergenic"]) == (1, 1)
    assert got["pct"] == (2 / 5 * 100.0, 1 / 5 * 100.0)
    # -n stops pass 1 after the first n records in file order (command.rs:312-315)
    got = run_one(gff, one(10), one(20, flag=0x4), one(30, ref=2), n_records=2)
    assert (got["processed"], got["ignored_flags"], got["ignored_nonprimary"]) == (1, 1, 0)
    # no records at all: 0/0 -> NaN (serialised as null)
    got = run_one(gff)
    assert all(np.isnan(x) for x in got["pct"])


def test_errors_abort_like_the_reference():
    gff = gff_line("chr1", "gene", 1, 90000)
    with pytest.raises(RuntimeError, match="reference sequence id"):
        run_one(gff, rec(name="m", flag=0, seq="A" * 10))                  # mapped flag, no reference id
    with pytest.raises(RuntimeError, match="read name"):
        run_one(gff, one(10, name="*"))
    with pytest.raises(RuntimeError, match="strand"):
        run_one(gff_line("chr1", "gene", 1, 9, "."), one(10))              # utils.rs:36-42
    run_one(gff_line("chrM", "gene", 1, 9, "."), one(10))                   # ... only for primary sequence names
    with pytest.raises(RuntimeError, match="GFF"):
        run_one("chr1\ttest\tgene\t0\t9\t.\t+\t.\tID=x\n", one(10))         # positions are 1-based
    with pytest.raises(RuntimeError, match="GFF"):
        run_one("chr1 gene 1 9\n", one(10))
    got = run_one(gff + "##FASTA\n>chr1\nACGT\n", one(10))                  # records() ends at the FASTA section
    assert got["intronic"] == 1


# ---- independent formulation: counts of overlapping intervals per kind, no ordering, no early exit ----

def _bam_records(raw: bytes):
    off, stream = 0, bytearray()
    while off < len(raw):
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        stream += zlib.decompress(raw[off + 18:off + bsize - 8], -15)
        off += bsize
    s = bytes(stream)
    p = 8 + int.from_bytes(s[4:8], "little")
    n_ref = int.from_bytes(s[p:p + 4], "little")
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = int.from_bytes(s[p:p + 4], "little")
        refs.append(s[p + 4:p + 4 + ln - 1].decode())
        p += 8 + ln
    while p < len(s):
        bs = int.from_bytes(s[p:p + 4], "little")
        ref, pos, l_name, _mq, _bin, n_cig, flag = struct.unpack_from("<iiBBHHH", s, p + 4)
        cig = struct.unpack_from(f"<{n_cig}I", s, p + 36 + l_name)
        yield refs, ref, pos, flag, sum(op >> 4 for op in cig if (op & 15) in (0, 2, 3, 7, 8))
        p += 4 + bs


def _python_features(raw, model, names, primary):
    """model: list of (seq, type, start, end)."""
    out = dict.fromkeys(KEYS, 0)
    by_seq = {}
    for seq, ty, a, b in model:
        by_seq.setdefault(seq, []).append((ty, a, b))
    for refs, ref, pos, flag, span in _bam_records(raw):
        if flag & 4:
            out["ignored_flags"] += 1
            continue
        if not primary(refs[ref]):
            out["ignored_nonprimary"] += 1
            continue
        lo, hi = pos + 1, pos + 1 + span + 1
        hits = [ty for ty, a, b in by_seq.get(refs[ref], []) if a < hi and b > lo]
        translation = [t for t in hits if t in names[:3]]
        chain = list(names[:3])                      # each overlapping feature takes the first free slot carrying its name
        slots = [False, False, False]
        for kind in set(translation):
            free = [i for i in range(3) if chain[i] == kind]
            for i in free[:translation.count(kind)]:
                slots[i] = True
        out["utr5"] += slots[0]
        out["utr3"] += slots[1]
        out["cds"] += slots[2]
        region = [t for t in hits if t not in names[:3] and t in names[3:]]
        has_gene = names[4] in region
        has_exon = names[3] in region and names[3] != names[4]
        out["exonic" if has_gene and has_exon else "intronic" if has_gene else "intergenic"] += 1
        out["processed"] += 1
    return out


@pytest.mark.parametrize("names", [NAMES, ("UTR", "UTR", "CDS", "exon", "gene"), ("UTR", "CDS", "CDS", "exon", "exon")])
def test_python_formulation_agrees_on_a_generated_bam(names):
    from ngs_b200 import ffi, formats
    bam, _, _ = ffi.synth_bam(3, 4000, level=1)
    raw = bam.tobytes()
    rng = random.Random(5)
    # a gene model dense around the positions the reads really occupy
    spots = [(refs[ref], pos) for refs, ref, pos, flag, _ in _bam_records(raw) if ref >= 0][::7]
    kinds = sorted(set(names)) + ["transcript"]
    model = []
    for seq, pos in spots:
        for _ in range(3):
            a = max(1, pos + rng.randrange(-3000, 3000))
            model.append((seq, rng.choice(kinds), a, a + rng.choice([0, 1, 50, 400, 5000, 60000])))
    rng.shuffle(model)
    gff = "##gff-version 3\n" + "".join(gff_line(s, t, a, b, rng.choice("+-")) for s, t, a, b in model)
    got = features(raw, gff, names=names)
    want = _python_features(raw, model, list(names), formats.is_primary)
    assert {k: got[k] for k in KEYS} == want
    assert got["processed"] > 1000 and got["intronic"] and got["intergenic"] and got["utr5"] and got["cds"]
    assert bool(got["exonic"]) == (names[3] != names[4])
