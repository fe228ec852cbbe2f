"""N>1 host logic on CPU (gloo, world_size 2): contig-exclusive shards + ONE sum-reduce of the packed
integer results reproduce the whole-file results, coverage included.  Each rank runs the
oracle on its shard of the logical file; the collective is the same single all-reduce bench.py issues."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack(o, n_ref, nbins):
    parts = [o["general"], o["tlen_hist"], np.array([o["tlen_processed"], o["tlen_ignored"]], dtype=np.uint64), o["gc_rec"],
             np.array([o["nonsensical"]], dtype=np.uint64)]
    q = np.zeros((160, 94), dtype=np.uint64)
    q[: o["quality"].shape[0]] = o["quality"]
    parts.append(q.ravel())
    for c in range(n_ref):
        slot = np.zeros(2 + 2049 + nbins[c], dtype=np.uint64)
        if c in o["coverage"]:
            cc = o["coverage"][c]
            slot[0] = 1
            slot[1] = cc["too_large"]
            slot[2:2051] = cc["hist"]
            slot[2051:2051 + len(cc["bin_sums"])] = cc["bin_sums"]
        parts.append(slot)
    return np.concatenate(parts).astype(np.int64)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle_ints
    from ngs_b200 import ffi
    sys.path.insert(0, ROOT)
    from bench import lpt_partition
    shape, n = 0, 30000
    per, tail = ffi.synth_layout(shape, n)
    parts, loads = lpt_partition(per, world)
    tail_rank = int(np.argmin(loads))
    mask = sum(1 << c for c in parts[rank])
    bam, bai, info = ffi.synth_bam(shape, n, contig_mask=mask, with_tail=(rank == tail_rank))
    o = oracle_ints(bam, bai)
    lens = [20000000, 10000000, 16569]
    nbins = [L // 50000 + 2 for L in lens]
    t = torch.from_numpy(_pack(o, 3, nbins))
    dist.all_reduce(t)  # the single merge
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_one_reduce_equals_whole_file(tmp_path):
    sys.path.insert(0, ROOT)
    from helpers import oracle_ints
    from ngs_b200 import ffi
    out = str(tmp_path / "merged.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = np.load(out)
    bam, bai, _ = ffi.synth_bam(0, 30000)
    whole = oracle_ints(bam, bai)
    lens = [20000000, 10000000, 16569]
    want = _pack(whole, 3, [L // 50000 + 2 for L in lens])
    np.testing.assert_array_equal(merged, want)
