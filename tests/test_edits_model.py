"""CPU check of the Edits kernel's per-record logic: ngs_b200/csrc/edits.cuh (edits_record, edits_encode,
edits_vaf_bin — what the CUDA kernel runs per lane), compiled for the host by tools/edits_model.cpp, must reproduce
the oracle's integers on BAM + FASTA cases and fail where the oracle (the reference) aborts.  Test tooling only:
nothing under ngs_b200/ links or calls the model, and the kernel itself has not run on a GPU yet."""
import os
import subprocess

import numpy as np
import pytest

from bamutil import rec, write_bam
from test_oracle_edits import REFS, make_edits_case, oracle_edits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("edits") / "edits_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "edits_model.cpp"), "-lz"], check=True)
    return exe


def run_model(model, tmp_path, bam: bytes, fa: bytes):
    (tmp_path / "x.bam").write_bytes(bam)
    (tmp_path / "x.fa").write_bytes(fa)
    r = subprocess.run([model, str(tmp_path / "x.bam"), str(tmp_path / "x.fa")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    if lines[0].startswith("error"):
        return int(lines[0].split()[1])
    out = {"records": int(lines[0].split()[1])}
    for ln in lines[1:]:
        k, *v = ln.split()
        out[k] = np.array([int(x) for x in v], dtype=np.uint64)
    return out


@pytest.mark.parametrize("seed", [17, 3, 99])
def test_model_matches_oracle(model, tmp_path, seed):
    bam, bai, fa, _, _ = make_edits_case(seed)
    one, two, vaf, n, _ = oracle_edits(bam, bai, fa)
    got = run_model(model, tmp_path, bam, fa)
    assert got["records"] == n > 100
    np.testing.assert_array_equal(got["read_one"], one)
    np.testing.assert_array_equal(got["read_two"], two)
    np.testing.assert_array_equal(got["vaf"], vaf)


def _one_record_case(refseq_chr1, **kw):
    seqs = {"chr1": refseq_chr1, "chr2": "A" * 3000, "chrM": "C" * 800}
    fa = "".join(f">{n}\n{s}\n" for n, s in seqs.items()).encode()
    r = dict(name="r", flag=0x43, ref=0, pos=10, mapq=30, cigar="20M", seq="ACGT" * 5, qual=[30] * 20)
    r.update(kw)
    bam, bai = write_bam(REFS, [rec(**r)])
    return bam, bai, fa


# (case, model's error kind, fragment of the oracle's message): the reference aborts the run on each of these
ERRORS = [
    (dict(refseq="acgt" * 1250), 5, "invalid base"),                                             # lower-case letters under the read
    (dict(refseq="ACGT" * 1250, name="*"), 2, "read name"),
    (dict(refseq="ACGT" * 100, pos=390), 4, "past the end"),                                     # FASTA shorter than the header says
    (dict(refseq="ACGT" * 1250, cigar="20M5S"), 7, "step-through"),                              # CIGAR wants more read bases
    (dict(refseq="ACGT" * 1250, cigar="15M"), 9, "step-through"),                                # read not fully consumed
    (dict(refseq="A" * 5000, cigar="600M", seq="C" * 600, qual=[30] * 600), 10, "512 edits"),
    (dict(refseq="ACGT" * 1500, pos=4990), 11, "beyond the sequence length"),                    # FASTA longer than the header: position > L
]


@pytest.mark.parametrize("case,code,msg", ERRORS, ids=[str(e[1]) for e in ERRORS])
def test_model_fails_where_the_oracle_aborts(model, tmp_path, case, code, msg):
    case = dict(case)
    bam, bai, fa = _one_record_case(case.pop("refseq"), **case)
    with pytest.raises(RuntimeError, match=msg):
        oracle_edits(bam, bai, fa)
    assert run_model(model, tmp_path, bam, fa) == code


def test_skipped_records_and_a_deletion_over_rejected_letters(model, tmp_path):
    """Unmapped / duplicate / unplaced records are skipped; a rejected letter under a deletion still aborts
    (the whole slice is converted first, edits.rs:259-265); outside every record it does not matter."""
    ref = "ACGT" * 1250
    seqs = {"chr1": ref[:2000] + "n" + ref[2001:], "chr2": "A" * 3000, "chrM": "C" * 800}
    fa = "".join(f">{n}\n{s}\n" for n, s in seqs.items()).encode()
    ok = [rec(name="a", flag=0x43, ref=0, pos=100, mapq=9, cigar="20M", seq=ref[100:120]),
          rec(name="d", flag=0x400 | 0x43, ref=0, pos=1990, mapq=9, cigar="20M", seq=ref[1990:2010]),   # duplicate over the 'n': skipped
          rec(name="u", flag=0x4, ref=0, pos=1990, mapq=0, cigar="20M", seq=ref[1990:2010]),            # unmapped: skipped
          rec(name="x", flag=0x4 | 0x1, seq="ACGT")]                                                    # unplaced
    bam, bai = write_bam(REFS, ok)
    one, two, vaf, n, _ = oracle_edits(bam, bai, fa)
    got = run_model(model, tmp_path, bam, fa)
    assert got["records"] == n == 1 and got["read_one"][0] == one[0] == 1 and int(got["vaf"].sum()) == int(vaf.sum()) == 20
    bad = ok[:1] + [rec(name="del", flag=0x83, ref=0, pos=1990, mapq=9, cigar="5M10D5M", seq=ref[1990:1995] + ref[2005:2010])]
    bam, bai = write_bam(REFS, bad)
    with pytest.raises(RuntimeError, match="invalid base"):
        oracle_edits(bam, bai, fa)
    assert run_model(model, tmp_path, bam, fa) == 5
