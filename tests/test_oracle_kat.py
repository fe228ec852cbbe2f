"""Pins the oracle (CPU restatement) against every known-answer vector the reference's own tests
hold for this path (src/utils/histogram.rs:405-523), and against hand-derived
answers for each facet rule on hand-built micro-BAMs (the reference ships no BAM fixture)."""
import ctypes as C
import json

import numpy as np
import pytest

from bamutil import as_u8, rec, write_bam
from helpers import OracleError, oracle_ints, oracle_lib


class H:
    def __init__(self, cap):
        self.lib = oracle_lib()
        self.h = self.lib.oracle_hist_new(cap)

    def inc(self, b, n=1):
        return self.lib.oracle_hist_inc_by(self.h, b, n)

    def pct(self, p):
        out = C.c_double(0)
        rc = self.lib.oracle_hist_percentile(self.h, p, C.byref(out))
        return None if rc == 1 else out.value

    def mean(self):
        return self.lib.oracle_hist_mean(self.h)


def test_histogram_mean_median_quartiles():  # histogram.rs:414-431
    h = H(100)
    h.inc(25); h.inc(50); h.inc(75, 3); h.inc(100, 5)
    assert h.mean() == 80.0
    assert h.pct(0.25) == 75.0
    assert h.pct(0.5) == 87.5
    assert h.pct(0.75) == 100.0
    assert h.pct(0.75) - h.pct(0.25) == 25.0


def test_histogram_empty_median_is_none():  # histogram.rs:434-437
    assert H(5000).pct(0.5) is None


def test_histogram_tie_median():  # histogram.rs:440-463
    h = H(5000)
    h.inc(0, 2500); h.inc(10, 2500); h.inc(100, 2500); h.inc(5000, 5000)
    assert h.pct(0.5) == 100.0
    h.inc(200, 2500)
    assert h.pct(0.5) == 150.0
    h.inc(200)
    assert h.pct(0.5) == 200.0


def test_histogram_out_of_bounds():  # histogram.rs:466-469
    assert H(100).inc(101) != 0


def test_histogram_values_and_cumulative_counts():  # histogram.rs:484-523
    lib = oracle_lib()
    h = H(3)
    h.inc(1); h.inc(2); h.inc(3, 3)
    assert [lib.oracle_hist_get(h.h, i) for i in range(4)] == [0, 1, 1, 3]
    assert lib.oracle_hist_len(h.h) == 4
    g = H(3)
    g.inc(0, 5); g.inc(1, 3); g.inc(2, 6)
    assert [lib.oracle_hist_bottom_until(g.h, i) for i in range(4)] == [5, 8, 14, 14]
    assert [lib.oracle_hist_top_until(g.h, i) for i in (3, 2, 1, 0)] == [0, 6, 9, 14]


REFS = [("chr1", 1000), ("chr2", 500), ("chrM", 300)]
P, U, MU, R1, R2, SEC, DUP, SUP, PROPER = 0x1, 0x4, 0x8, 0x40, 0x80, 0x100, 0x400, 0x800, 0x2


def run(recs_, **kw):
    bam, bai = write_bam(REFS, recs_)
    return oracle_ints(as_u8(bam), as_u8(bai), **kw)


def test_general_decision_tree():  # general.rs:31-101, hand-derived
    recs = [
        rec(name="a", flag=P | R1 | PROPER, ref=0, pos=10, mapq=60, cigar="10M", next_ref=0, next_pos=50, tlen=50, seq="ACGTACGTAC"),
        rec(name="b", flag=P | R2, ref=0, pos=20, mapq=255, cigar="5M2I3M", next_ref=1, next_pos=5, tlen=0, seq="ACGTACGTAC"),  # mismatch, hq (255 counts)
        rec(name="c", flag=P | R1, ref=0, pos=30, mapq=4, cigar="10M", next_ref=1, next_pos=5, seq="ACGTACGTAC"),             # mismatch, not hq
        rec(name="d", flag=P | R1 | MU, ref=0, pos=40, mapq=60, cigar="4S6M", next_ref=0, next_pos=40, seq="ACGTACGTAC"),     # singleton
        rec(name="e", flag=P | R2 | SEC | DUP, ref=0, pos=50, mapq=60, cigar="10M", next_ref=0, next_pos=1, seq="ACGTACGTAC"),
        rec(name="f", flag=SUP, ref=0, pos=60, mapq=60, cigar="3H10M", seq="ACGTACGTAC"),
        rec(name="g", flag=DUP, ref=1, pos=5, mapq=0, cigar="10=", seq="ACGTACGTAC"),                                         # unpaired primary dup
        rec(name="h", flag=P | U | R1, ref=1, pos=7, next_ref=1, next_pos=7, seq="ACGT"),                                      # placed unmapped
        rec(name="i", flag=P | U | MU | R2, seq="ACGT"),                                                                       # unplaced
    ]
    g = run(recs, coverage=False)["general"]
    names = ["total", "unmapped", "duplicate", "primary", "secondary", "supplementary", "primary_mapped", "primary_duplicate", "paired",
             "read_1", "read_2", "proper_pair", "singleton", "mate_mapped", "mismatch", "mismatch_hq"]
    got = dict(zip(names, map(int, g[:16])))
    assert got == dict(total=9, unmapped=2, duplicate=2, primary=7, secondary=1, supplementary=1, primary_mapped=5, primary_duplicate=1,
                       paired=6, read_1=4, read_2=2, proper_pair=1, singleton=1, mate_mapped=3, mismatch=2, mismatch_hq=1)
    # CIGAR kinds, read one = flag 0x40 regardless of pairing; everything else is "read two" (general.rs:103-121)
    ops = "MIDNSHP=X"
    one = {ops[k]: int(g[16 + k]) for k in range(9) if g[16 + k]}
    two = {ops[k]: int(g[25 + k]) for k in range(9) if g[25 + k]}
    assert one == {"M": 3, "S": 1}
    assert two == {"M": 4, "I": 1, "H": 1, "=": 1}


def test_general_panics_without_reference_ids():  # general.rs:81-83 unwrap()
    with pytest.raises(OracleError, match="panic"):
        run([rec(name="x", flag=P | R1, ref=0, pos=1, cigar="4M", next_ref=-1, seq="ACGT")], coverage=False)


def test_template_length_rules():  # template_length.rs:79-100: negatives are ignored, 0..=1024 binned
    recs = [rec(name=f"t{i}", flag=0, ref=0, pos=i, cigar="4M", seq="ACGT", tlen=t) for i, t in enumerate([0, 0, 1, 1024, 1025, -1, -300, 500])]
    r = run(recs, coverage=False)
    assert (r["tlen_processed"], r["tlen_ignored"]) == (5, 3)
    assert r["tlen_hist"][0] == 2 and r["tlen_hist"][1] == 1 and r["tlen_hist"][1024] == 1 and r["tlen_hist"][500] == 1


def test_gc_rules():  # gc_content.rs:38-100
    s100 = "G" * 30 + "C" * 20 + "A" * 25 + "T" * 20 + "N" * 5
    recs = [
        rec(name="ok", flag=0, ref=0, pos=1, cigar="100M", seq=s100),
        rec(name="dup", flag=DUP, ref=0, pos=2, cigar="100M", seq=s100),
        rec(name="sec", flag=SEC, ref=0, pos=3, cigar="100M", seq=s100),
        rec(name="short", flag=0, ref=0, pos=4, cigar="99M", seq=s100[:99]),
        rec(name="l101", flag=0, ref=0, pos=5, cigar="101M", seq=s100 + "G"),  # max_offset 1 -> offset always 0
    ]
    r = run(recs, coverage=False)
    assert list(r["gc_rec"]) == [2, 2, 1]
    assert list(r["gc_nuc"]) == [100, 90, 10]
    assert r["gc_hist"][50] == 2 and r["gc_hist"].sum() == 2


def test_quality_rules():  # quality_scores.rs:37-49 + presence rule
    recs = [
        rec(name="q1", flag=0, ref=0, pos=1, cigar="4M", seq="ACGT", qual=[0, 10, 93, 40]),
        rec(name="q2", flag=0, ref=0, pos=2, cigar="2M", seq="AC", qual=[10, 10]),
        rec(name="none", flag=0, ref=0, pos=3, cigar="6M", seq="ACGTAC"),  # all 0xFF: no scores, no positions
        rec(name="empty", flag=U),
    ]
    q = run(recs, coverage=False)["quality"]
    assert q.shape == (4, 94)
    assert q[0, 0] == 1 and q[0, 10] == 1 and q[1, 10] == 2 and q[2, 93] == 1 and q[3, 40] == 1 and q.sum() == 6


def test_quality_above_93_aborts():
    with pytest.raises(OracleError, match="93"):
        run([rec(name="bad", flag=0, ref=0, pos=1, cigar="2M", seq="AC", qual=[94, 10])], coverage=False)


def test_coverage_rules():  # coverage.rs:148-287 + query filter
    recs = [
        rec(name="a", flag=0, ref=0, pos=0, cigar="10M", seq="A" * 10),             # positions 1..10
        rec(name="b", flag=DUP | SEC, ref=0, pos=4, cigar="2M3D2M", seq="A" * 4),  # 5..11: D counts, flags ignored
        rec(name="c", flag=0, ref=0, pos=8, cigar="2M100N2M", seq="A" * 4),        # 9..112: N skip counts as covered
        rec(name="d", flag=U | P, ref=0, pos=20, next_ref=0, next_pos=20, seq="A" * 4),  # placed unmapped, span 0, start 21: touches, adds nothing
        rec(name="e", flag=0, ref=0, pos=995, cigar="10M", seq="A" * 10),          # 996..1005: 5 positions beyond L=1000
        rec(name="m", flag=0, ref=2, pos=5, cigar="10M", seq="A" * 10),            # chrM: not primary assembly
    ]
    r = run(recs, records=False)
    assert sorted(r["coverage"]) == [0]  # chr2 untouched, chrM unsupported
    c = r["coverage"][0]
    assert r["nonsensical"] == 5
    depth = np.zeros(1001, dtype=np.int64)
    depth[1:11] += 1; depth[5:12] += 1; depth[9:113] += 1; depth[996:1001] += 1
    want = np.bincount(depth, minlength=2049)
    np.testing.assert_array_equal(c["hist"], want)   # position 0 is counted as depth 0
    assert c["too_large"] == 0
    assert list(c["bin_sums"]) == [0, int(depth.sum())]  # bin 0 = {0}; tail bin = the rest (L < 50000)


def test_coverage_start_one_span_zero_is_filtered():  # alignment_end() == None -> query drops the record
    recs = [rec(name="z", flag=U | P, ref=1, pos=0, next_ref=1, next_pos=0, seq="ACGT")]
    assert run(recs, records=False)["coverage"] == {}


def test_num_records_pass2_counter_is_shared():  # command.rs:384-388, SURVEY App. C.7
    recs = [rec(name=f"a{i}", flag=0, ref=0, pos=10 + i, cigar="5M", seq="AAAAA") for i in range(4)]
    recs += [rec(name=f"b{i}", flag=0, ref=1, pos=10 + i, cigar="5M", seq="AAAAA") for i in range(3)]
    r = run(recs, n_records=2)
    assert r["general"][0] == 2
    assert r["coverage"][0]["hist"][1:].sum() > 0
    # chr2 still processes exactly its first returned record
    d = np.zeros(501, dtype=np.int64); d[11:16] += 1
    np.testing.assert_array_equal(r["coverage"][1]["hist"], np.bincount(d, minlength=2049))


def test_unknown_sequence_name_is_rejected():  # command.rs:258-272
    bam, bai = write_bam([("contigX", 100)], [])
    with pytest.raises(OracleError, match="not found"):
        oracle_ints(as_u8(bam), as_u8(bai))


def test_oracle_json_schema(tmp_path):  # results.rs:24-45 field set
    recs = [rec(name="a", flag=0, ref=0, pos=0, cigar="10M", seq="A" * 10, qual=[30] * 10)]
    bam, bai = write_bam(REFS, recs)
    p = tmp_path / "o.json"
    oracle_ints(as_u8(bam), as_u8(bai), json_path=str(p))
    j = json.load(open(p))
    assert list(j) == ["general", "features", "gc_content", "template_length", "quality_scores", "coverage", "edits"]
    assert j["features"] is None and j["edits"] is None
    assert j["coverage"]["mean_coverage"]["chr1"] == 10 / 1001
    assert set(j["coverage"]["genome_covered_by"]) == {"10x", "20x", "30x", "40x", "50x", "60x"}
    assert j["template_length"]["histogram"]["range_stop"] == 1024
    assert list(j["quality_scores"]["scores"]) == [str(i) for i in range(1, 11)]
