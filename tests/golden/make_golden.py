"""Regenerates tests/golden/synthetic_digests.json: sha256 digests of the oracle's canonical results
JSON on seeded synthetic BAMs (shape, records, zlib level, gc seed).  The reference itself cannot run
here (Rust, no toolchain) and ships no BAM fixture, so these pin the oracle + generator pair against
regressions; the hand-derived answers that pin the oracle's RULES live in tests/test_oracle_kat.py.
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = [  # (shape, n_records, level, gc_seed)
    (0, 20000, 6, 0), (0, 20000, 1, 5), (1, 30000, 6, 7), (3, 20000, 6, 7), (2, 400, 6, 7),
]


def digest_case(shape, n, level, seed):
    from helpers import oracle_ints, results_digest
    from ngs_b200 import ffi
    bam, bai, info = ffi.synth_bam(shape, n, level=level)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "o.json")
        o = oracle_ints(bam, bai, gc_seed=seed, json_path=p)
        return {"shape": shape, "records": n, "level": level, "gc_seed": seed, "bam_bytes": int(bam.size),
                "inflated_bytes": info["inflated_bytes"], "total": int(o["general"][0]), "unmapped": int(o["general"][1]),
                "quality_positions": int(o["quality"].shape[0]), "nonsensical": int(o["nonsensical"]), "sha256": results_digest(p)}


if __name__ == "__main__":
    out = [digest_case(*c) for c in CASES]
    with open(os.path.join(HERE, "synthetic_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
