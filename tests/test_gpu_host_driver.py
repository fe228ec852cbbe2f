"""End-to-end through the product's host driver (ngs-cuda-qc, C++ above the C ABI): the results JSON
must equal the oracle's after canonicalisation — integers identical, f64 identical (computed on the host
in the reference's operation order, so the 1e-12 relative bound of the north_star is met with zero
error), f32 identical after float32 rounding."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from helpers import canonical_results, oracle_ints, results_digest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "ngs_b200", "ngs-cuda-qc")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic_digests.json")))


def _write(tmp_path, shape, n, level=6, name="s.bam"):
    from ngs_b200 import ffi
    bam, bai, info = ffi.synth_bam(shape, n, level=level)
    p = str(tmp_path / name)
    bam.tofile(p)
    bai.tofile(p + ".bai")
    return p, bam, bai


def _run(args, check=True):
    r = subprocess.run([DRIVER] + args, capture_output=True, text=True)
    if check and r.returncode != 0:
        raise AssertionError(f"driver failed: {r.stderr}")
    return r


def _max_rel(a, b):
    worst = 0.0
    def walk(x, y):
        nonlocal worst
        if isinstance(x, dict):
            assert x.keys() == y.keys()
            for k in x:
                walk(x[k], y[k])
        elif isinstance(x, list):
            assert len(x) == len(y)
            for u, v in zip(x, y):
                walk(u, v)
        elif isinstance(x, float) or isinstance(y, float):
            if x != y:
                worst = max(worst, abs(x - y) / max(abs(x), abs(y)))
        else:
            assert x == y
    walk(a, b)
    return worst


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: f"shape{c['shape']}-{c['records']}-l{c['level']}")
def test_results_json_equals_oracle_and_golden(tmp_path, case):
    p, bam, bai = _write(tmp_path, case["shape"], case["records"], case["level"])
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path / "out"), "--cuda-gc-seed", str(case["gc_seed"])])
    got_path = str(tmp_path / "out" / "s.bam.results.json")  # prefix defaults to the file name (command.rs:156-162)
    want_path = str(tmp_path / "oracle.json")
    oracle_ints(bam, bai, gc_seed=case["gc_seed"], json_path=want_path)
    got, want = canonical_results(got_path), canonical_results(want_path)
    assert _max_rel(got, want) <= 1e-12  # tolerance stated by the north_star; observed 0
    assert got == want
    assert results_digest(got_path) == case["sha256"]
    # struct declaration order of the reference is preserved in the file itself
    raw = json.load(open(got_path))
    assert list(raw) == ["general", "features", "gc_content", "template_length", "quality_scores", "coverage", "edits"]
    assert list(raw["general"]["records"])[:4] == ["total", "unmapped", "duplicate", "designation"]


def test_only_flag_and_prefix(tmp_path):
    p, bam, bai = _write(tmp_path, 0, 8000)
    _run(["qc", p, "grch38_no_alt_analysisset", "-o", str(tmp_path), "-p", "x", "--only", "gc content"])
    j = json.load(open(tmp_path / "x.results.json"))
    assert j["general"] is None and j["coverage"] is None and j["gc_content"]["records"]["processed"] > 0
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "y", "--only", "Coverage"])
    j = json.load(open(tmp_path / "y.results.json"))
    assert j["general"] is None and set(j["coverage"]["mean_coverage"]) == {"chr1", "chr2"}
    r = _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "--only", "nope"], check=False)
    assert r.returncode != 0 and "No facets matched" in r.stderr


def test_num_records(tmp_path):
    p, bam, bai = _write(tmp_path, 0, 8000)
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "n", "-n", "1234", "--only", "General"])
    got = json.load(open(tmp_path / "n.results.json"))
    want_path = str(tmp_path / "o.json")
    oracle_ints(bam, bai, n_records=1234, coverage=False, json_path=want_path)
    assert got["general"] == json.load(open(want_path))["general"]


def test_num_records_with_every_default_facet(tmp_path):
    """`-n` without `--only`: pass 1 stops after n records, pass 2 applies the reference's shared counter
    (command.rs:350-397) — the whole results file must be the oracle's."""
    p, bam, bai = _write(tmp_path, 0, 30000)
    for n in (1234, 20000):
        _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", f"n{n}", "-n", str(n), "--cuda-gc-seed", "4"])
        want_path = str(tmp_path / f"o{n}.json")
        oracle_ints(bam, bai, n_records=n, gc_seed=4, json_path=want_path)
        assert canonical_results(str(tmp_path / f"n{n}.results.json")) == canonical_results(want_path)


def test_progress_lines_and_io_modes(tmp_path):
    """RecordCounter's log line (display.rs:43-52) every millionth record; O_DIRECT and buffered reads, small and large
    chunks give the same file."""
    p, bam, bai = _write(tmp_path, 0, 2_100_000, level=1)
    r = _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "a", "--cuda-chunk-mb", "8", "--cuda-perf"])
    assert "  [*] Processed 1,000,000 records." in r.stderr and "  [*] Processed 2,000,000 records." in r.stderr
    assert "Processed 2100000 records." in r.stderr
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "b", "--cuda-chunk-mb", "1", "--cuda-buffered-io"])
    assert results_digest(str(tmp_path / "a.results.json")) == results_digest(str(tmp_path / "b.results.json"))
    want_path = str(tmp_path / "o.json")
    oracle_ints(bam, bai, json_path=want_path)
    assert canonical_results(str(tmp_path / "a.results.json")) == canonical_results(want_path)
    perf = json.load(open(tmp_path / "a.perf.json"))
    assert perf[0]["records"] == 2_100_000 and perf[0]["wall_ms_file_to_results"] > 0 and perf[0]["waves"] >= 1


def test_read_longer_than_the_default_quality_table_is_retried(tmp_path):
    """One 140 kb read: the engine reports NGSQ_E_QUAL_CAP, the driver enlarges the table and streams the file again."""
    from bamutil import rec, write_bam
    n = 140_000
    r = [rec(name="short", flag=0, ref=0, pos=10, mapq=30, cigar="50M", seq="ACGTA" * 10, qual=[30] * 50),
         rec(name="ultralong", flag=0, ref=0, pos=100, mapq=30, cigar=f"{n}M", seq="ACGT" * (n // 4), qual=[7 + (i % 40) for i in range(n)])]
    bam, bai = write_bam([("chr1", 248956422)], r)
    p = str(tmp_path / "u.bam")
    open(p, "wb").write(bam)
    open(p + ".bai", "wb").write(bai)
    res = _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "u"])
    assert "enlarging the quality table" in res.stderr
    want_path = str(tmp_path / "o.json")
    oracle_ints(np.frombuffer(bam, dtype=np.uint8), np.frombuffer(bai, dtype=np.uint8), json_path=want_path)
    assert canonical_results(str(tmp_path / "u.results.json")) == canonical_results(want_path)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_devices_shard_one_file_by_bai_ranges(tmp_path):
    """--cuda-devices 0,1: plan_shards cuts the file at BAI contig extents, one engine per GPU, one NCCL reduce; the
    results file must be the single-device one (GC window histogram included: offsets are keyed by virtual offsets)."""
    p, bam, bai = _write(tmp_path, 1, 400000)
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "two", "--cuda-devices", "0,1", "--cuda-gc-seed", "6"])
    want_path = str(tmp_path / "o.json")
    oracle_ints(bam, bai, gc_seed=6, json_path=want_path)
    assert canonical_results(str(tmp_path / "two.results.json")) == canonical_results(want_path)


def test_errors_match_reference_behaviour(tmp_path):
    p, bam, bai = _write(tmp_path, 0, 2000)
    r = _run(["qc", p, "hg19", "-o", str(tmp_path)], check=False)
    assert r.returncode != 0 and "reference genome is not supported" in r.stderr  # command.rs:126-134
    os.remove(p + ".bai")
    r = _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path)], check=False)
    assert r.returncode != 0 and "index" in r.stderr.lower()  # IndexCheck::Full, bam.rs:86-96
    from bamutil import write_bam
    b, i = write_bam([("contigX", 100)], [])
    q = str(tmp_path / "bad.bam")
    open(q, "wb").write(b)
    open(q + ".bai", "wb").write(i)
    r = _run(["qc", q, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path)], check=False)
    assert r.returncode != 0 and "not found in specified reference genome" in r.stderr  # command.rs:258-272


def test_float_formatting_is_ryu_style(tmp_path):
    p, bam, bai = _write(tmp_path, 0, 3000)
    _run(["qc", p, "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "f"])
    text = open(tmp_path / "f.results.json").read()
    assert '"mean_coverage_per_bin": {\n      "chr1": [\n        0.0,' in text  # first per-bin mean is always 0/50000
    assert '\n  "features": null,' in text and text.rstrip().endswith("}")
