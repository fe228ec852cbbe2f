"""GPU parity for the Edits facet: every integer `ngsq_get_edits` returns must equal the oracle's
on the same BAM + FASTA, and the engine must fail where the oracle (the reference) aborts."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

from bamutil import as_u8, rec, write_bam  # noqa: E402
from test_edits_model import ERRORS, _one_record_case  # noqa: E402
from test_oracle_edits import REFS, _python_edits, make_edits_case, oracle_edits  # noqa: E402


def _fasta_sequences(fa: bytes):
    out, name = {}, None
    for ln in fa.decode().splitlines():
        if ln.startswith(">"):
            name = ln[1:].split()[0]
            out[name] = []
        elif name is not None:
            out[name].append(ln.strip())
    return {k: "".join(v).encode() for k, v in out.items()}


def engine_edits(bam: bytes, fa: bytes, record_facets=True, launch_blocks=0, want_engine=False, n_records=0, coverage=True):
    """record_facets=False is `--only`-style: the one-record cases below carry a mapped pair without mate reference ids,
    on which the General facet of the reference panics (general.rs:81-83) before Edits ever sees the record."""
    from ngs_b200 import ffi, formats
    b = as_u8(bam)
    eng = ffi.Engine(flags=(ffi.NGSQ_F_RECORD_FACETS if record_facets else 0) | (ffi.NGSQ_F_COVERAGE if coverage else 0) | ffi.NGSQ_F_VERIFY_CRC
                     | ffi.NGSQ_F_EDITS, launch_blocks=launch_blocks, max_records=n_records)
    hdr = formats.read_header(eng, b)
    eng.set_references([l for _, l in hdr.refs], [1 if formats.is_primary(n) else 0 for n, _ in hdr.refs])
    seqs = _fasta_sequences(fa)
    for c, (name, _) in enumerate(hdr.refs):
        if name in seqs:
            eng.set_reference_bases(c, seqs[name])
    eng.set_range(hdr.first_voffset, 0)
    eng.submit(np.ascontiguousarray(b), 0)
    eng.finish()
    if want_engine:
        return eng, hdr
    return eng.edits(), eng.stats()


@pytest.mark.parametrize("seed", [17, 3, 99])
def test_edits_match_oracle(seed):
    bam, bai, fa, _, _ = make_edits_case(seed)
    one, two, vaf, n, _ = oracle_edits(bam, bai, fa)
    for launch_blocks in (0, 3):  # one wave, and many: the per-position counters live across waves
        (g_one, g_two, g_vaf, g_n), st = engine_edits(bam, fa, launch_blocks=launch_blocks)
        assert g_n == n > 100
        np.testing.assert_array_equal(g_one, one)
        np.testing.assert_array_equal(g_two, two)
        np.testing.assert_array_equal(g_vaf, vaf)
        assert st["ms_edits"] > 0


@pytest.mark.parametrize("case,code,msg", ERRORS, ids=[str(e[1]) for e in ERRORS])
def test_engine_fails_where_the_oracle_aborts(case, code, msg):
    from ngs_b200 import ffi
    case = dict(case)
    bam, bai, fa = _one_record_case(case.pop("refseq"), **case)
    with pytest.raises(RuntimeError, match=msg):
        oracle_edits(bam, bai, fa)
    with pytest.raises(ffi.NgsqError) as ei:
        engine_edits(bam, fa, record_facets=False)
    assert ei.value.code == -11  # NGSQ_E_EDITS


def test_other_facets_are_unchanged_by_the_edits_flag():
    from helpers import assert_same_ints, engine_ints, oracle_ints
    from ngs_b200 import ffi
    bam, bai, fa, _, _ = make_edits_case(5)
    b, x = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, x, gc_seed=2)
    got = engine_ints(b, gc_seed=2)
    assert_same_ints(got, want)


def test_a_contig_without_a_sequence_fails_only_if_it_holds_records():
    from ngs_b200 import ffi
    ref = "ACGT" * 1250
    raw = [rec(name="a", flag=0x43, ref=0, pos=100, mapq=9, cigar="20M", seq=ref[100:120])]
    bam, bai = write_bam(REFS, raw)
    (one, two, vaf, n), _ = engine_edits(bam, f">chr1\n{ref}\n".encode(), record_facets=False)   # chr2 / chrM have no sequence and no records
    assert n == 1 and one[0] == 1 and int(vaf.sum()) == 20
    with pytest.raises(ffi.NgsqError):
        engine_edits(bam, b">chr2\nACGT\n", record_facets=False)


def test_per_position_counters_match_the_restated_step_through():
    """ngsq_get_edit_positions (the VAF file's input, edits.rs:279-288 / 317-340): refs and alts of every 1-based position."""
    from ngs_b200 import ffi
    bam, bai, fa, refseqs, recs = make_edits_case(5)
    want = {}
    _python_edits(refseqs, recs, positions=want)
    for launch_blocks in (0, 2):
        eng, hdr = engine_edits(bam, fa, launch_blocks=launch_blocks, want_engine=True)
        for c, (name, L) in enumerate(hdr.refs):
            refs, alts = eng.edit_positions(c, L)
            np.testing.assert_array_equal(refs, want[c][0].astype(np.uint32), err_msg=f"refs of {name}")
            np.testing.assert_array_equal(alts, want[c][1].astype(np.uint32), err_msg=f"alts of {name}")
            assert refs[0] == 0 and alts[0] == 0
        with pytest.raises(ffi.NgsqError):
            eng.edit_positions(0, hdr.refs[0][1] + 1)   # the length must be the header's


@pytest.mark.parametrize("n", [1, 7, 161, 300])
def test_num_records_follows_the_second_pass_counter(n):
    """`-n` with Edits (command.rs:375-388): one counter over all sequences, incremented for every yielded record (also those
    the facet skips), the limit only ends the current sequence's loop.  With and without Coverage beside it (they share the
    marks of cov_n.cuh), one wave and many."""
    bam, bai, fa, _, _ = make_edits_case(11)
    one, two, vaf, cnt, _ = oracle_edits(bam, bai, fa, n_records=n)
    for coverage, launch_blocks in ((True, 0), (False, 0), (True, 2)):
        (g_one, g_two, g_vaf, g_n), _ = engine_edits(bam, fa, launch_blocks=launch_blocks, n_records=n, coverage=coverage)
        assert g_n == cnt
        np.testing.assert_array_equal(g_one, one)
        np.testing.assert_array_equal(g_two, two)
        np.testing.assert_array_equal(g_vaf, vaf)
