"""Oracle for the NEXT row of the scope table (the Edits facet): pinned against the
reference's own unit tests of the step-through (src/utils/alignment.rs:134-202) and cross-checked against
an independent pure-Python restatement of src/qc/sequence_based/edits.rs:217-344 on a hand-built BAM +
FASTA.  No device path consumes this yet."""
import ctypes as C

import numpy as np
import pytest

from bamutil import as_u8, parse_cigar, rec, write_bam
from helpers import oracle_lib

CODES = "=ACMGRSVTWYHKDBN"


def _lib():
    lib = oracle_lib()
    P = C.c_void_p
    lib.oracle_stepthrough_edits.restype = C.c_int
    lib.oracle_stepthrough_edits.argtypes = [P, C.c_uint64, P, C.c_uint64, P, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.oracle_edits_run.restype = P
    lib.oracle_edits_run.argtypes = [P, C.c_size_t, P, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.oracle_edits_run_n.restype = P
    lib.oracle_edits_run_n.argtypes = [P, C.c_size_t, P, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64]
    lib.oracle_edits_get.argtypes = [P, P, P, P, P, C.POINTER(C.c_uint64)]
    return lib


def _edits(reference: str, record: str, cigar: str):
    lib = _lib()
    r = np.array([CODES.index(c) for c in reference], dtype=np.uint8)
    q = np.array([CODES.index(c) for c in record], dtype=np.uint8)
    ops = np.array([(l << 4) | k for l, k in parse_cigar(cigar)], dtype=np.uint32)
    n = C.c_uint64(0)
    rc = lib.oracle_stepthrough_edits(r.ctypes.data, r.size, q.ctypes.data, q.size, ops.ctypes.data, ops.size, C.byref(n))
    return rc, n.value


# ---- known-answer vectors lifted from the reference's tests (alignment.rs:134-202)
def test_zero_edits_when_sequences_are_identical():
    assert _edits("ACTG", "ACTG", "4M") == (0, 0)


def test_one_edit():
    assert _edits("AATG", "ACTG", "4M") == (0, 1)


def test_softclips():
    assert _edits("ACTG", "ACTGACTG", "4M4S") == (0, 0)


def test_malformed_record_with_too_few_record_bases():
    rc, _ = _edits("ACTG", "ACTGACTG", "4M5S")
    assert rc == 2  # "...consume a record base, but no such base was found"


def test_malformed_record_with_too_few_reference_bases():
    rc, _ = _edits("ACTG", "ACT", "3M2D")
    assert rc == 1  # "...consume a reference base, but no such base was found"


def test_sequences_must_be_fully_consumed():
    assert _edits("ACTGA", "ACTG", "4M")[0] == 3  # reference sequence was not fully consumed
    assert _edits("ACTG", "ACTGA", "4M")[0] == 4  # record sequence was not fully consumed


def test_only_m_counts_as_a_comparison():
    # "=" and "X" consume both sequences but are not Kind::Match (edits.rs:274)
    assert _edits("ACGT", "TTTT", "2=2X") == (0, 0)
    assert _edits("ACGTAC", "AGTTAC", "6M") == (0, 2)
    # insertions / deletions / skips move the two pointers without comparing
    assert _edits("ACGTTTAC", "ACGGGAC", "3M2I3D2M") == (0, 0)


# ---- the facet on a BAM + FASTA, against a pure-Python restatement
REFS = [("chr1", 5000), ("chr2", 3000), ("chrM", 800)]


def _python_edits(refseqs, recs, positions=None, vaf_lines=None, n_records=0):
    """positions (dict) / vaf_lines (list), when given, receive the per-position counters of every sequence and the lines of the
    VAF file of edits.rs:317-340 (f32 printed like Rust's `{}`: shortest round-trip, positional)."""
    one, two, vaf = np.zeros(513, np.uint64), np.zeros(513, np.uint64), np.zeros(101, np.uint64)
    n = 0
    counter = 0  # `-n`: one counter over all sequences, incremented after every yielded record (command.rs:375-388)
    for c, (name, L) in enumerate(REFS):
        refs, alts = np.zeros(L + 1, np.uint64), np.zeros(L + 1, np.uint64)
        stop = False
        for r in recs:
            if r["ref"] != c or stop:   # every record of make_edits_case lies inside its sequence: the query yields it
                continue
            counter += 1
            if n_records and counter >= n_records:
                stop = True             # this record is still processed; the loop of THIS sequence ends after it
            if r["flag"] & (0x4 | 0x400):
                continue
            rp, qp, e = 0, 0, 0
            start = r["pos"] + 1
            for ln, k in parse_cigar(r["cigar"]):
                for _ in range(ln):
                    cr, cs = k in (0, 2, 3, 7, 8), k in (0, 1, 4, 7, 8)
                    if k == 0:
                        if refseqs[name][start - 1 + rp] != r["seq"][qp]:
                            e += 1
                            alts[start + rp] += 1
                        else:
                            refs[start + rp] += 1
                    rp += cr
                    qp += cs
            (one if r["flag"] & 0x40 else two)[e] += 1
            n += 1
        tot = refs + alts
        if positions is not None:
            positions[c] = (refs.copy(), alts.copy())
        for i in np.nonzero(tot)[0]:
            v = np.float32(alts[i]) / np.float32(tot[i])
            vaf[int(v * np.float32(100.0))] += 1
            if vaf_lines is not None:
                vaf_lines.append(f"{name}\t{i}\t" + np.format_float_positional(v, unique=True, trim="-"))
    return one, two, vaf, n


def make_edits_case(seed=17):
    """(bam bytes, bai bytes, fasta bytes, reference strings, record dicts): mixed CIGAR shapes, substitutions, flags."""
    rng = np.random.default_rng(seed)
    refseqs = {name: "".join(rng.choice(list("ACGT"), size=L)) for name, L in REFS}
    recs = []
    for c, (name, L) in enumerate(REFS):
        pos = 0
        for i in range(160):
            pos += int(rng.integers(0, L // 200 + 1))
            n = int(rng.integers(30, 90))
            if pos + n + 40 >= L:
                break
            shape = int(rng.integers(0, 6))
            ref = refseqs[name]
            if shape == 0:
                cig, parts = f"{n}M", [("M", ref[pos:pos + n])]
            elif shape == 1:
                cig, parts = f"5S{n}M", [("S", "ACGTA"), ("M", ref[pos:pos + n])]
            elif shape == 2:
                cig, parts = f"{n // 2}M3I{n - n // 2}M", [("M", ref[pos:pos + n // 2]), ("I", "GGG"), ("M", ref[pos + n // 2:pos + n])]
            elif shape == 3:
                cig, parts = f"{n // 2}M4D{n - n // 2}M", [("M", ref[pos:pos + n // 2]), ("M", ref[pos + n // 2 + 4:pos + n + 4])]
            elif shape == 4:
                cig, parts = f"{n // 2}M20N{n - n // 2}M", [("M", ref[pos:pos + n // 2]), ("M", ref[pos + n // 2 + 20:pos + n + 20])]
            else:
                cig, parts = f"10={n - 10}M", [("M", ref[pos:pos + n])]
            seq = list("".join(p for _, p in parts))
            for k in range(len(seq)):  # substitutions; some records share positions so VAFs vary
                if rng.random() < 0.06:
                    seq[k] = str(rng.choice(list("ACGTN")))
            flag = int(rng.choice([0x43, 0x83, 0x63, 0x93, 0x400 | 0x43, 0x4 | 0x41, 0x0, 0x100 | 0x83]))
            recs.append(dict(ref=c, pos=pos, cigar=cig, seq="".join(seq), flag=flag, name=f"r{c}_{i}"))
    # paired records carry their mate's reference id: the General facet of the reference panics on a mapped pair
    # without one (general.rs:81-83), and the GPU tests run this case with every facet enabled
    raw = [rec(name=r["name"], flag=r["flag"], ref=r["ref"], pos=r["pos"], mapq=30, cigar=r["cigar"], seq=r["seq"],
               next_ref=r["ref"] if r["flag"] & 1 else -1, next_pos=r["pos"] if r["flag"] & 1 else -1,
               qual=[30] * len(r["seq"])) for r in recs]
    bam, bai = write_bam(REFS, raw, block_payload=3000)
    fasta = "".join(f">{name} synthetic\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n" for name, s in refseqs.items())
    return bam, bai, fasta.encode(), refseqs, recs


def oracle_edits(bam: bytes, bai: bytes, fa: bytes, n_records=0):
    """(read_one, read_two, vaf, records, means) from the oracle, or raises RuntimeError with its message."""
    lib = _lib()
    b, x = as_u8(bam), as_u8(bai)
    h = lib.oracle_edits_run_n(b.ctypes.data, b.size, x.ctypes.data, x.size, fa, len(fa), n_records)
    if not h:
        raise RuntimeError(lib.oracle_last_error().decode())
    one, two, vaf = np.zeros(513, np.uint64), np.zeros(513, np.uint64), np.zeros(101, np.uint64)
    means = (C.c_double * 2)()
    n = C.c_uint64(0)
    lib.oracle_edits_get(h, one.ctypes.data, two.ctypes.data, vaf.ctypes.data, means, C.byref(n))
    return one, two, vaf, n.value, (means[0], means[1])


def test_facet_matches_python_restatement():
    bam, bai, fa, refseqs, recs = make_edits_case(17)
    lib = _lib()
    b, x = as_u8(bam), as_u8(bai)
    h = lib.oracle_edits_run(b.ctypes.data, b.size, x.ctypes.data, x.size, fa, len(fa))
    assert h, lib.oracle_last_error().decode()
    one, two, vaf = np.zeros(513, np.uint64), np.zeros(513, np.uint64), np.zeros(101, np.uint64)
    means = (C.c_double * 2)()
    n = C.c_uint64(0)
    lib.oracle_edits_get(h, one.ctypes.data, two.ctypes.data, vaf.ctypes.data, means, C.byref(n))
    w_one, w_two, w_vaf, w_n = _python_edits(refseqs, recs)
    assert n.value == w_n > 100
    np.testing.assert_array_equal(one, w_one)
    np.testing.assert_array_equal(two, w_two)
    np.testing.assert_array_equal(vaf, w_vaf)
    k = np.arange(513, dtype=np.float64)
    assert means[0] == pytest.approx(float((w_one * k).sum() / w_one.sum()), rel=1e-12)
    assert means[1] == pytest.approx(float((w_two * k).sum() / w_two.sum()), rel=1e-12)
    assert vaf.sum() > 1000 and vaf[1:100].sum() > 0  # intermediate VAFs occur: positions covered by several reads


def test_lower_case_reference_bases_abort_the_run():
    """Base::try_from accepts the sixteen upper-case letters only (edits.rs:259-265 propagates the error)."""
    refseqs = {name: "acgt" * (L // 4) + "a" * (L % 4) for name, L in REFS}
    raw = [rec(name="r", flag=0x43, ref=0, pos=10, mapq=30, cigar="20M", seq="ACGT" * 5, qual=[30] * 20)]
    bam, bai = write_bam(REFS, raw)
    fasta = "".join(f">{name}\n{s}\n" for name, s in refseqs.items()).encode()
    lib = _lib()
    b, x = as_u8(bam), as_u8(bai)
    assert not lib.oracle_edits_run(b.ctypes.data, b.size, x.ctypes.data, x.size, fasta, len(fasta))
    assert b"invalid base" in lib.oracle_last_error()


@pytest.mark.parametrize("n", [1, 7, 150, 161, 300, 100000])
def test_num_records_counts_every_yielded_record_and_only_ends_the_current_sequence(n):
    """`-n` in pass 2 (command.rs:375-388): records the facet skips (unmapped / duplicate) count too, and once the limit is
    reached every later sequence still contributes its first yielded record."""
    bam, bai, fa, refseqs, recs = make_edits_case(11)
    one, two, vaf, cnt = _python_edits(refseqs, recs, n_records=n)
    o1, o2, ov, on, _ = oracle_edits(bam, bai, fa, n_records=n)
    assert on == cnt
    np.testing.assert_array_equal(o1, one)
    np.testing.assert_array_equal(o2, two)
    np.testing.assert_array_equal(ov, vaf)
    if n == 1:
        assert cnt <= len(REFS)   # at most one record per sequence
