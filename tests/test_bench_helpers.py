"""Host-side planning logic of bench.py: contig-exclusive LPT sharding for 2/4/8 ranks and the
shape-preserving CPU sample (whole contigs of the logical BAM at full depth)."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("n_ranks,total", [(2, 150_000_000), (4, 300_000_000), (8, 600_000_000)])
def test_lpt_partition_is_exclusive_and_balanced(n_ranks, total):
    from bench import lpt_partition
    from ngs_b200 import ffi
    per_contig, tail = ffi.synth_layout(1, total)
    parts, loads = lpt_partition(per_contig, n_ranks)
    owned = sorted(c for p in parts for c in p)
    assert owned == list(range(len(per_contig)))  # every contig has exactly one owner: coverage needs that
    assert sum(loads) == sum(per_contig)
    assert max(loads) <= 1.08 * np.mean(loads)  # 25 contigs pack within a few per cent


@pytest.mark.parametrize("total", [10_000_000, 100_000_000, 150_000_000, 300_000_000, 600_000_000])
def test_cpu_sample_takes_whole_contigs_near_the_target(total, monkeypatch):
    import bench
    from ngs_b200 import ffi
    per_contig, _ = ffi.synth_layout(1, total)
    seen = {}

    def fake_synth_bam(shape, n, level=6, contig_mask=0, with_tail=True, **kw):
        seen.update(mask=contig_mask, n=n, tail=with_tail)
        cnt = sum(per_contig[c] for c in range(len(per_contig)) if contig_mask >> c & 1)
        return np.zeros(1, np.uint8), np.zeros(1, np.uint8), {"n_records": cnt}

    monkeypatch.setattr(ffi, "synth_bam", fake_synth_bam)
    args = types.SimpleNamespace(cpu_sample=3_000_000)
    wl = bench.workload("wgs", 1, total)
    _, _, info, desc = bench.cpu_sample(args, wl)
    assert seen["n"] == total and seen["tail"] is False
    chosen = [c for c in range(len(per_contig)) if seen["mask"] >> c & 1]
    big = [c for c in chosen if per_contig[c] > 0.01 * max(per_contig)]  # chromosomes, not chrM
    assert big, "at least one real chromosome"
    # about the requested size: within 1.5x unless a single chromosome is already larger
    assert info["n_records"] <= max(1.5 * args.cpu_sample, min(per_contig[c] for c in big) * 1.01)
    assert info["n_records"] >= 0.5 * args.cpu_sample
    assert str(info["n_records"]) in desc


@pytest.mark.parametrize("name,n_ranks", [("wgs", 1), ("wgs", 2), ("wgs", 4), ("wgs", 8), ("c1", 1), ("c4", 1), ("c5", 1)])
def test_workloads_are_the_named_configs_and_their_shards_cover_every_contig_once(name, n_ranks):
    import bench
    wl = bench.workload(name, n_ranks)
    want_total = {"wgs": 100_000_000 if n_ranks == 1 else 75_000_000 * n_ranks, "c1": 1_000_000, "c4": 2_000_000, "c5": 200_000_000}[name]
    assert wl["total_records"] == want_total and wl["key"] == f"{name}_n{n_ranks}_{want_total}_l{wl['level']}"
    masks = [bench.workload_shard(wl, r) for r in range(n_ranks)]
    assert sum(int(t) for _, t in masks) == 1                      # the unplaced tail has one owner
    union = 0
    for m, _ in masks:
        assert union & m == 0                                      # contig-exclusive
        union |= m
    from ngs_b200 import ffi
    assert union == (1 << len(ffi.synth_layout(wl["shape"], want_total)[0])) - 1


def test_committed_fullsize_goldens_load_and_name_their_workload():
    import glob
    import bench
    from helpers import GOLDEN_DIR, load_fullsize_golden
    files = glob.glob(os.path.join(GOLDEN_DIR, "fullsize_*.npz"))
    assert files, "no full-size golden committed"
    for f in files:
        key = os.path.basename(f)[len("fullsize_"):-4]
        ints, meta = load_fullsize_golden(key)
        name, n = meta["key"].split("_")[0], meta["n_ranks"]
        assert bench.workload(name, n)["key"] == key               # the bench would look this file up
        assert meta["records"] == meta["total_records"] == sum(meta["shard_records"])
        if "general" in ints:
            assert int(ints["general"][0]) == meta["records"]      # General.records.total counts every record once
