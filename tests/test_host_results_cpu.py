"""The product's host-side arithmetic and JSON writer on CPU: ngs_b200/host/facets.hpp (summarize /
teardown / aggregate in the reference's operation order) and results.hpp are fed the oracle's integers
through a test double of the C ABI getters (tests/cpp/fake_engine.cpp) and must write the oracle's results
JSON — integers identical, f64 identical, f32 identical after float32 rounding.  The GPU suite repeats the
comparison with the real engine (tests/test_gpu_host_driver.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import canonical_results, oracle_ints

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_results(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hostres") / "host_results")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_results.cpp"),
                    os.path.join(ROOT, "tests", "cpp", "fake_engine.cpp")], check=True)
    return exe


def _dump_ints(path, o, lens):
    u = lambda a: np.asarray(a, dtype=np.uint64).ravel()  # noqa: E731
    parts = [u(o["general"]), u(o["tlen_hist"]), u([o["tlen_processed"], o["tlen_ignored"]]), u(o["gc_hist"]), u(o["gc_nuc"]), u(o["gc_rec"]),
             u([o["quality"].shape[0]]), u(o["quality"]), u([o["nonsensical"]]), u([len(lens)])]
    for c, L in enumerate(lens):
        cc = o["coverage"].get(c)
        if cc is None:
            parts += [u([0, 0, 0]), np.zeros(2049, np.uint64)]
        else:
            parts += [u([1, len(cc["bin_sums"]), cc["too_large"]]), u(cc["hist"]), u(cc["bin_sums"])]
    np.concatenate(parts).tofile(path)


@pytest.mark.parametrize("shape,n,seed", [(0, 30000, 5), (3, 20000, 9), (2, 300, 2)])
def test_host_facets_write_the_oracles_json(host_results, tmp_path, shape, n, seed):
    from ngs_b200 import ffi, formats
    bam, bai, _ = ffi.synth_bam(shape, n, level=6)
    want_path = str(tmp_path / "oracle.json")
    o = oracle_ints(bam, bai, gc_seed=seed, json_path=want_path)
    # reference names / lengths from the BAM header (parsed on the host from zlib-inflated header blocks)
    import zlib
    raw = bam.tobytes()
    hdr, off = b"", 0
    while len(hdr) < 1 << 16 and off < len(raw):
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        hdr += zlib.decompress(raw[off + 18:off + bsize - 8], -15)
        off += bsize
    l_text = int.from_bytes(hdr[4:8], "little")
    p = 8 + l_text
    n_ref = int.from_bytes(hdr[p:p + 4], "little")
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = int.from_bytes(hdr[p:p + 4], "little")
        name = hdr[p + 4:p + 4 + ln - 1].decode()
        L = int.from_bytes(hdr[p + 4 + ln:p + 8 + ln], "little")
        refs.append((name, L))
        p += 8 + ln
    (tmp_path / "refs.tsv").write_text("".join(f"{nm}\t{L}\n" for nm, L in refs))
    ints = str(tmp_path / "ints.bin")
    _dump_ints(ints, o, [L for _, L in refs])
    env = dict(os.environ, NGSQ_FAKE_INTS=ints)
    subprocess.run([host_results, str(tmp_path / "refs.tsv"), str(tmp_path), "host"], check=True, env=env)
    got, want = canonical_results(str(tmp_path / "host.results.json")), canonical_results(want_path)
    assert got == want
    rawj = json.load(open(tmp_path / "host.results.json"))
    assert list(rawj) == ["general", "features", "gc_content", "template_length", "quality_scores", "coverage", "edits"]
