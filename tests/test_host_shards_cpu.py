"""The C++ host driver's shard planner (ngs_b200/host/bam.hpp) on CPU: for any number of devices the shards must be a
contiguous, contig-exclusive partition of the file cut at record starts taken from the BAI, every shard's byte range must
hold whole BGZF blocks through its last record, and the records the BAI counts per contig must add up.  (On GPUs this
path only runs with `--cuda-devices a,b,..`; bench.py shards in Python.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shards") / "host_shards")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "cpp", "host_shards.cpp"),
                    "-L" + os.path.join(ROOT, "ngs_b200"), "-lngs_cuda", "-Wl,-rpath," + os.path.join(ROOT, "ngs_b200")], check=True)
    return out


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    from ngs_b200 import ffi, formats
    d = tmp_path_factory.mktemp("shardcase")
    bam, bai, info = ffi.synth_bam(1, 60000, level=1)
    (d / "x.bam").write_bytes(bam.tobytes())
    (d / "x.bam.bai").write_bytes(bai.tobytes())
    import zlib
    raw = bam.tobytes()
    hdr, off, sizes = b"", 0, []
    while len(hdr) < info["header_bytes"]:
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        blk = zlib.decompress(raw[off + 18:off + bsize - 8], -15)
        sizes.append((off, len(blk)))
        hdr += blk
        off += bsize
    acc = 0
    for co, n in sizes:  # virtual offset of the first record = the byte right after the header
        if info["header_bytes"] < acc + n:
            first = (co << 16) | (info["header_bytes"] - acc)
            break
        acc += n
    else:
        first = off << 16
    return str(d / "x.bam"), first, info, formats.parse_bai(bai.tobytes()), raw


@pytest.mark.parametrize("n_shards", [1, 2, 3, 4, 8, 24, 40])
def test_shards_partition_the_file(exe, case, n_shards):
    path, first, info, bai, raw = case
    out = subprocess.run([exe, path, str(first), str(info["n_ref"]), str(n_shards)], capture_output=True, text=True, check=True).stdout
    assert not out.startswith("error"), out
    rows = [[int(x) for x in ln.split()] for ln in out.splitlines()]
    assert len(rows) == n_shards
    live = [r for r in rows if not r[2]]
    assert live and live[0][0] == first and live[-1][1] == 0                       # starts at the first record, runs to EOF
    starts = {r.ref_beg for r in bai.refs if r.ref_beg is not None and r.ref_end and r.ref_end > r.ref_beg}
    for a, b in zip(live, live[1:]):
        assert a[1] == b[0] and a[1] in starts                                       # cut at the first record of a contig
        assert a[3] < a[4] and b[3] <= a[4]                                           # byte ranges meet or share the cut block
    owned = [c for r in live for c in r[5:]]
    with_records = [c for c, r in enumerate(bai.refs) if r.ref_beg is not None and r.ref_end and r.ref_end > r.ref_beg]
    assert len(owned) == len(set(owned))                                              # contig-exclusive
    if len(live) > 1:
        assert sorted(owned) == with_records
    for r in live:                                                                    # whole BGZF blocks: lo and hi are block starts
        for o in (r[3], r[4]):
            assert o == len(raw) or raw[o:o + 2] == b"\x1f\x8b"
    assert len(live) == min(n_shards, len(with_records)) or n_shards == 1
