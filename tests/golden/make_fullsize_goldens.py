#!/usr/bin/env python
"""Full-size goldens for bench.py: the oracle's integers on exactly the BAMs the bench times.

bench.py's workloads are deterministic functions of (shape, records, zlib level, N): at N ranks every rank
generates the shard of the logical BAM that `bench.workload_shards` assigns to it (contig-exclusive, LPT by record
count; the unplaced tail goes to the lightest rank).  Every facet is additive over contig-exclusive shards, so
the golden of a workload is the sum of the oracle's integers over its shard files — the same files, byte for
byte, that the ranks write on the GPU box (same generator, same zlib).

    python tests/golden/make_fullsize_goldens.py wgs:1 wgs:2 wgs:4 wgs:8 c1:1 c5:1 ...   [--workers 3]

writes tests/golden/fullsize_<key>.npz (a few hundred KB each).  CPU only; about 5 minutes of one core per
75 M records.  bench.py compares the TIMED runs' own result buffers with these files (helpers.compare_fullsize).
"""
import argparse
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def shard_ints(job):
    """One shard file of a workload through the oracle (runs in a worker process)."""
    shape, total, level, mask, with_tail, gc_seed, records, coverage, threads = job
    from helpers import oracle_ints
    from ngs_b200 import ffi
    t0 = time.perf_counter()
    bam, bai, info = ffi.synth_bam(shape, total, level=level, contig_mask=mask, with_tail=with_tail, threads=threads)
    t1 = time.perf_counter()
    out = oracle_ints(bam, bai, gc_seed=gc_seed, records=records, coverage=coverage)
    t2 = time.perf_counter()
    out["_info"] = dict(info, compressed_bytes=int(bam.size), gen_s=t1 - t0, oracle_s=t2 - t1)
    return out


def main():
    import bench
    from helpers import merge_ints, save_fullsize_golden
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+", help="<shape name>:<n ranks>[:<records per gpu>]")
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0, help="generator threads per worker (0 = all cores)")
    ap.add_argument("--split", type=int, default=0,
                    help="compute every shard in this many contig-exclusive pieces (facets are additive over them): for workloads "
                         "whose shard file does not fit this machine's memory (configs[3]: 50 GB compressed)")
    args = ap.parse_args()
    t0 = time.perf_counter()
    for w in args.workloads:  # one workload at a time: a finished golden is on disk before the next one starts
        f = w.split(":")
        name, n = f[0], int(f[1])
        wl = bench.workload(name, n, int(f[2]) if len(f) > 2 else 0)
        jobs = []
        for rank in range(n):
            mask, with_tail = bench.workload_shard(wl, rank)
            pieces = [(mask, with_tail)]
            if args.split > 1:
                contigs = [c for c in range(mask.bit_length()) if mask >> c & 1]
                pieces = [(sum(1 << c for c in contigs[k::args.split]), with_tail and k == 0) for k in range(args.split)]
                pieces = [p for p in pieces if p[0] or p[1]]
            for m, t in pieces:
                jobs.append((wl["shape"], wl["total_records"], wl["level"], m, t, bench.GC_SEED, wl["records"], wl["coverage"], args.threads))
        with ProcessPoolExecutor(max_workers=min(args.workers, len(jobs))) as ex:
            parts = list(ex.map(shard_ints, jobs))
        merged = merge_ints(parts, records=wl["records"], coverage=wl["coverage"])
        meta = {"key": wl["key"], "shape": wl["shape"], "total_records": wl["total_records"], "level": wl["level"], "n_ranks": wl["n_ranks"],
                "gc_seed": bench.GC_SEED, "gc_window": args.split <= 1, "records": sum(p["_info"]["n_records"] for p in parts),
                "inflated_bytes": sum(p["_info"]["inflated_bytes"] for p in parts),
                "compressed_bytes": sum(p["_info"]["compressed_bytes"] for p in parts),
                "shard_records": [int(p["_info"]["n_records"]) for p in parts] if args.split <= 1 else
                                 [sum(int(p["_info"]["n_records"]) for p in parts)] if n == 1 else None,
                "oracle_s": sum(p["_info"]["oracle_s"] for p in parts)}
        path = save_fullsize_golden(merged, meta)
        print(f"{w}: {meta['records']} records, oracle {meta['oracle_s']:.0f} core-seconds -> {path} ({os.path.getsize(path)} bytes)", flush=True)
    print(f"done in {time.perf_counter() - t0:.0f} s")


if __name__ == "__main__":
    main()
