"""CPU check of the Genomic Features kernel's per-record logic: ngs_b200/csrc/features.cuh (features_record — counts of
overlapping features per name from two binary searches, no interval tree), compiled for the host by
tools/features_model.cpp, must reproduce the oracle's nine counters, which walk rust-lapper's ordered overlaps.
Test tooling only; the kernel itself has not run on a GPU yet."""
import os
import random
import subprocess

import pytest

from bamutil import write_bam
from test_oracle_features import KEYS, NAMES, REFS, _bam_records, features, gff_line, one, rec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("features") / "features_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "features_model.cpp"), "-lz"], check=True)
    return exe


def run_model(model, tmp_path, bam: bytes, gff: str, names=NAMES, n_records=0):
    (tmp_path / "x.bam").write_bytes(bam)
    (tmp_path / "x.gff").write_text(gff)
    r = subprocess.run([model, str(tmp_path / "x.bam"), str(tmp_path / "x.gff"), *names, str(n_records)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if r.stdout.startswith("error"):
        return int(r.stdout.split()[1])
    return dict(zip(KEYS, (int(x) for x in r.stdout.split())))


NAME_SETS = [NAMES, ("UTR", "UTR", "CDS", "exon", "gene"), ("UTR", "CDS", "CDS", "exon", "exon"), ("x", "x", "x", "gene", "gene"),
             ("five_prime_UTR", "three_prime_UTR", "exon", "exon", "gene"), ("gene", "three_prime_UTR", "CDS", "exon", "gene")]


@pytest.mark.parametrize("names", NAME_SETS, ids=["-".join(n) for n in NAME_SETS])
def test_model_matches_oracle_on_a_generated_bam(model, tmp_path, names):
    from ngs_b200 import ffi
    bam, _, _ = ffi.synth_bam(3, 4000, level=1)
    raw = bam.tobytes()
    rng = random.Random(11)
    spots = [(refs[ref], pos) for refs, ref, pos, flag, _ in _bam_records(raw) if ref >= 0][::5]
    kinds = sorted(set(names)) + ["transcript"]
    lines = []
    for seq, pos in spots:
        for _ in range(4):
            a = max(1, pos + rng.randrange(-3000, 3000))
            lines.append(gff_line(seq, rng.choice(kinds), a, a + rng.choice([0, 1, 2, 50, 400, 5000, 60000]), rng.choice("+-")))
    rng.shuffle(lines)
    gff = "##gff-version 3\n" + "".join(lines)
    want = features(raw, gff, names=names)
    got = run_model(model, tmp_path, raw, gff, names=names)
    assert got == {k: want[k] for k in KEYS}
    assert got["processed"] > 1000
    want = features(raw, gff, names=names, n_records=777)
    assert run_model(model, tmp_path, raw, gff, names=names, n_records=777) == {k: want[k] for k in KEYS}


def test_model_matches_oracle_on_the_interval_edges(model, tmp_path):
    gff = gff_line("chr1", "gene", 2000, 3000) + gff_line("chr1", "exon", 2500, 2500) + gff_line("chr1", "CDS", 2600, 2601)
    recs = [one(p) for p in (1897, 1898, 1899, 1900, 2398, 2399, 2400, 2498, 2499, 2500, 2598, 2599, 2600, 2998, 2999, 3000)]
    recs += [one(20, flag=0x4), one(30, ref=2), one(5, ref=3), rec(name="u", flag=0x4 | 0x1, seq="A" * 100)]
    bam, _ = write_bam(REFS, recs)
    want = features(bam, gff)
    assert run_model(model, tmp_path, bam, gff) == {k: want[k] for k in KEYS}
    assert want["exonic"] and want["intronic"] and want["intergenic"] and want["cds"] and want["ignored_flags"] == 2 and want["ignored_nonprimary"] == 1


@pytest.mark.parametrize("record,code,msg", [
    (dict(pos=10, name="*"), 1, "read name"),
    (dict(pos=10, name="*", flag=0x4), 1, "read name"),          # the name is parsed before the flags are looked at
    (dict(pos=-1, ref=-1), 2, "reference sequence id"),          # mapped flag, no reference id
])
def test_model_fails_where_the_oracle_aborts(model, tmp_path, record, code, msg):
    gff = gff_line("chr1", "gene", 1, 90000)
    bam, _ = write_bam(REFS, [one(**record)])
    with pytest.raises(RuntimeError, match=msg):
        features(bam, gff)
    assert run_model(model, tmp_path, bam, gff) == code
