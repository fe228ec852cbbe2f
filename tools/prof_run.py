"""Small driver for ncu captures and timing: passes over a synthetic BAM through the host-facing ABI.
usage: prof_run.py [n_records] [zlib level] [repetitions] [shape]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ngs_b200 import ffi, formats

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # 0 small 3-contig, 1 WGS 2x150, 2 long reads, 3 RNA-seq
bam, bai, info = ffi.synth_bam(shape, n, level=level)
eng = ffi.Engine(flags=7, gc_seed=7, reserve_compressed=bam.size, reserve_inflated=info["inflated_bytes"] + 65536, reserve_blocks=info["n_blocks"] + 16)
hdr = formats.read_header(eng, bam)
eng.set_references([l for _, l in hdr.refs], [1 if formats.is_primary(nm) else 0 for nm, _ in hdr.refs])
eng.set_range(hdr.first_voffset, 0)
for _ in range(reps):
    eng.reset()
    eng.submit(bam, 0)
    eng.finish()
    print(eng.stats())
