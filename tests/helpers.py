"""Test helpers: the oracle (CPU restatement, oracle/ngsqc_oracle.c) through ctypes, and a
driver that pushes a BAM through the C ABI and collects every integer the engine returns.
Only tests/, smoke() and bench.py's cpu_baseline may touch oracle/."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(ORACLE_SO)
        P = C.c_void_p
        lib.oracle_run.restype = P
        lib.oracle_run.argtypes = [P, C.c_size_t, P, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_write_json.argtypes = [P, C.c_char_p]
        lib.oracle_free.argtypes = [P]
        for name in ["oracle_pass1_records", "oracle_pass2_records", "oracle_inflated_bytes", "oracle_quality_positions"]:
            getattr(lib, name).restype = C.c_uint64
            getattr(lib, name).argtypes = [P]
        lib.oracle_n_ref.restype = C.c_uint32
        lib.oracle_n_ref.argtypes = [P]
        lib.oracle_get_general.argtypes = [P, P]
        lib.oracle_get_tlen.argtypes = [P, P, P, P]
        lib.oracle_get_gc.argtypes = [P, P, P, P]
        lib.oracle_get_quality.argtypes = [P, P]
        lib.oracle_cov_touched.argtypes = [P, C.c_uint32]
        lib.oracle_cov_nbins.restype = C.c_uint64
        lib.oracle_cov_nbins.argtypes = [P, C.c_uint32]
        lib.oracle_get_cov_contig.argtypes = [P, C.c_uint32, P, P, P]
        lib.oracle_get_cov_dist.argtypes = [P, P, P]
        lib.oracle_inflate_all.restype = C.c_int64
        lib.oracle_inflate_all.argtypes = [P, C.c_size_t, P, C.c_size_t]
        lib.oracle_hist_new.restype = P
        lib.oracle_hist_new.argtypes = [C.c_uint64]
        lib.oracle_hist_free.argtypes = [P]
        lib.oracle_hist_inc_by.argtypes = [P, C.c_uint64, C.c_uint64]
        lib.oracle_hist_get.restype = C.c_uint64
        lib.oracle_hist_get.argtypes = [P, C.c_uint64]
        lib.oracle_hist_len.restype = C.c_uint64
        lib.oracle_hist_len.argtypes = [P]
        lib.oracle_hist_mean.restype = C.c_double
        lib.oracle_hist_mean.argtypes = [P]
        lib.oracle_hist_percentile.argtypes = [P, C.c_double, C.POINTER(C.c_double)]
        lib.oracle_hist_sum.restype = C.c_uint64
        lib.oracle_hist_sum.argtypes = [P]
        lib.oracle_hist_top_until.restype = C.c_uint64
        lib.oracle_hist_top_until.argtypes = [P, C.c_uint64]
        lib.oracle_hist_bottom_until.restype = C.c_uint64
        lib.oracle_hist_bottom_until.argtypes = [P, C.c_uint64]
        _oracle = lib
    return _oracle


class OracleError(RuntimeError):
    pass


def oracle_ints(bam: np.ndarray, bai: np.ndarray, n_records=0, gc_seed=0, records=True, coverage=True, json_path=None):
    """Runs the oracle; returns the same integer dict as engine_ints."""
    lib = oracle_lib()
    bam = np.ascontiguousarray(bam)
    bai = np.ascontiguousarray(bai)
    h = lib.oracle_run(bam.ctypes.data, bam.size, bai.ctypes.data, bai.size, n_records, gc_seed, int(records), int(coverage))
    if not h:
        raise OracleError(lib.oracle_last_error().decode())
    out = {}
    if records:
        g = np.zeros(34, dtype=np.uint64)
        lib.oracle_get_general(h, g.ctypes.data)
        out["general"] = g
        th = np.zeros(1025, dtype=np.uint64)
        p, i = C.c_uint64(0), C.c_uint64(0)
        lib.oracle_get_tlen(h, th.ctypes.data, C.addressof(p), C.addressof(i))
        out["tlen_hist"], out["tlen_processed"], out["tlen_ignored"] = th, p.value, i.value
        gh, nuc, rec = np.zeros(101, dtype=np.uint64), np.zeros(3, dtype=np.uint64), np.zeros(3, dtype=np.uint64)
        lib.oracle_get_gc(h, gh.ctypes.data, nuc.ctypes.data, rec.ctypes.data)
        out["gc_hist"], out["gc_nuc"], out["gc_rec"] = gh, nuc, rec
        nq = lib.oracle_quality_positions(h)
        q = np.zeros((max(nq, 1), 94), dtype=np.uint64)
        if nq:
            lib.oracle_get_quality(h, q.ctypes.data)
        out["quality"] = q[:nq]
    if coverage:
        n_ref = lib.oracle_n_ref(h)
        cov = {}
        for c in range(n_ref):
            if lib.oracle_cov_touched(h, c):
                nb = lib.oracle_cov_nbins(h, c)
                hist = np.zeros(2049, dtype=np.uint64)
                ign = C.c_uint64(0)
                bins = np.zeros(nb, dtype=np.uint64)
                lib.oracle_get_cov_contig(h, c, hist.ctypes.data, C.addressof(ign), bins.ctypes.data)
                cov[c] = {"hist": hist, "too_large": ign.value, "bin_sums": bins}
        out["coverage"] = cov
        dist = np.zeros(2049, dtype=np.uint64)
        ns = C.c_uint64(0)
        lib.oracle_get_cov_dist(h, dist.ctypes.data, C.addressof(ns))
        out["nonsensical"] = ns.value
    out["pass1_records"] = lib.oracle_pass1_records(h)
    if json_path:
        lib.oracle_write_json(h, json_path.encode())
    lib.oracle_free(h)
    return out


def engine_ints(bam: np.ndarray, n_records=0, gc_seed=0, records=True, coverage=True, chunk_bytes=None, launch_blocks=0,
                crc=True, device=0, shard=None, engine=None):
    """Pushes a whole BAM (or one shard of it) through the C ABI; returns integers + stats."""
    from ngs_b200 import ffi, formats
    flags = (ffi.NGSQ_F_RECORD_FACETS if records else 0) | (ffi.NGSQ_F_COVERAGE if coverage else 0) | (ffi.NGSQ_F_VERIFY_CRC if crc else 0)
    eng = engine or ffi.Engine(device=device, flags=flags, gc_seed=gc_seed, max_records=n_records, launch_blocks=launch_blocks)
    hdr = formats.read_header(eng, bam)
    names = [n for n, _ in hdr.refs]
    lens = [l for _, l in hdr.refs]
    enabled = [1 if formats.is_primary(n) else 0 for n in names]
    first, end, lo, hi = hdr.first_voffset, 0, hdr.first_voffset >> 16, bam.size
    if shard is not None:
        first, end, lo, hi, owned = shard
        enabled = [e if c in owned else 0 for c, e in enumerate(enabled)]
    eng.set_references(lens, enabled)
    eng.set_range(first, end)
    data = bam[lo:hi]
    if chunk_bytes is None:
        eng.submit(np.ascontiguousarray(data), lo)
    else:
        o = 0
        while o < data.size:
            piece = data[o:o + chunk_bytes]
            _, n, used = ffi.bgzf_walk(piece, lo + o)
            if used == 0:
                piece = data[o:]
                used = piece.size
            eng.submit(np.ascontiguousarray(piece[:used]), lo + o)
            o += used
    eng.finish()
    out = collect(eng, lens, enabled, records, coverage)
    out["refs"] = hdr.refs
    out["engine"] = eng
    return out


def collect(eng, lens, enabled, records=True, coverage=True):
    out = {}
    if records:
        out["general"] = eng.general()
        out["tlen_hist"], out["tlen_processed"], out["tlen_ignored"] = eng.tlen()
        out["gc_hist"], out["gc_nuc"], out["gc_rec"] = eng.gc()
        out["quality"] = eng.quality()
    if coverage:
        cov = {}
        for c, L in enumerate(lens):
            if not enabled[c]:
                continue
            cc = eng.coverage_contig(c, L)
            if cc.touched:
                cov[c] = {"hist": cc.hist, "too_large": cc.pileup_too_large, "bin_sums": cc.bin_sums}
        out["coverage"] = cov
        out["nonsensical"] = eng.nonsensical_records()
    out["stats"] = eng.stats()
    return out


def assert_same_ints(got, want, records=True, coverage=True, gc_window=True):
    """gc_window=False: the two sides saw the records at different BGZF virtual offsets (separately
    written shard files), so the GC window offsets differ: compare the
    deterministic GC fields and check the invariants of the others."""
    if records:
        for k in ["general", "tlen_hist", "gc_rec"] + (["gc_hist", "gc_nuc"] if gc_window else []):
            np.testing.assert_array_equal(got[k], want[k], err_msg=k)
        if not gc_window:
            assert int(got["gc_hist"].sum()) == int(got["gc_rec"][0]) == int(want["gc_hist"].sum())
            assert int(got["gc_nuc"].sum()) == 100 * int(got["gc_rec"][0])
        assert got["tlen_processed"] == want["tlen_processed"]
        assert got["tlen_ignored"] == want["tlen_ignored"]
        assert got["quality"].shape == want["quality"].shape, (got["quality"].shape, want["quality"].shape)
        np.testing.assert_array_equal(got["quality"], want["quality"], err_msg="quality")
    if coverage:
        assert sorted(got["coverage"]) == sorted(want["coverage"]), "touched contigs differ"
        for c in want["coverage"]:
            for k in ["hist", "bin_sums"]:
                np.testing.assert_array_equal(got["coverage"][c][k], want["coverage"][c][k], err_msg=f"coverage[{c}].{k}")
            assert got["coverage"][c]["too_large"] == want["coverage"][c]["too_large"]
        assert got["nonsensical"] == want["nonsensical"]


def canonical(obj):
    """Key-sorted JSON value for comparing results files (map order is random in the reference)."""
    return json.loads(json.dumps(obj, sort_keys=True))


def canonical_results(path_or_obj):
    """Parsed, key-sorted results JSON; genome_covered_by values pass through float32 (they are f32 in
    the reference, coverage.rs:282, and writers may print them with different digit counts)."""
    obj = json.load(open(path_or_obj)) if isinstance(path_or_obj, str) else path_or_obj
    cov = obj.get("coverage")
    if cov and cov.get("genome_covered_by"):
        cov["genome_covered_by"] = {k: (None if v is None else float(np.float32(v))) for k, v in cov["genome_covered_by"].items()}
    return json.loads(json.dumps(obj, sort_keys=True))


def results_digest(path_or_obj) -> str:
    import hashlib
    return hashlib.sha256(json.dumps(canonical_results(path_or_obj), sort_keys=True).encode()).hexdigest()


# ---------------------------------------------------------------- full-size goldens (bench.py workloads)
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def merge_ints(parts, records=True, coverage=True):
    """Sum of the integer dicts of contig-exclusive shards (what ngsq_reduce computes on the device)."""
    out = {}
    if records:
        for k in ["general", "tlen_hist", "gc_hist", "gc_nuc", "gc_rec"]:
            out[k] = sum(p[k].astype(np.uint64) for p in parts)
        for k in ["tlen_processed", "tlen_ignored"]:
            out[k] = int(sum(int(p[k]) for p in parts))
        rows = max(p["quality"].shape[0] for p in parts)
        q = np.zeros((rows, 94), dtype=np.uint64)
        for p in parts:
            q[: p["quality"].shape[0]] += p["quality"]
        out["quality"] = q
    if coverage:
        cov = {}
        for p in parts:
            for c, v in p["coverage"].items():
                assert c not in cov, f"contig {c} touched by two shards: shards must be contig-exclusive"
                cov[c] = v
        out["coverage"] = cov
        out["nonsensical"] = int(sum(int(p["nonsensical"]) for p in parts))
    return out


def save_fullsize_golden(ints, meta):
    arrays = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for k in ["general", "tlen_hist", "gc_hist", "gc_nuc", "gc_rec", "quality"]:
        if k in ints:
            arrays[k] = np.asarray(ints[k], dtype=np.uint64)
    if "tlen_processed" in ints:
        arrays["tlen_scalars"] = np.array([ints["tlen_processed"], ints["tlen_ignored"]], dtype=np.uint64)
    if "coverage" in ints:
        contigs = sorted(ints["coverage"])
        arrays["cov_contigs"] = np.array(contigs, dtype=np.int64)
        arrays["cov_hist"] = np.array([ints["coverage"][c]["hist"] for c in contigs], dtype=np.uint64).reshape(len(contigs), 2049)
        arrays["cov_too_large"] = np.array([ints["coverage"][c]["too_large"] for c in contigs], dtype=np.uint64)
        arrays["cov_nbins"] = np.array([len(ints["coverage"][c]["bin_sums"]) for c in contigs], dtype=np.int64)
        arrays["cov_bins"] = np.concatenate([np.asarray(ints["coverage"][c]["bin_sums"], dtype=np.uint64) for c in contigs]) if contigs else np.zeros(0, np.uint64)
        arrays["nonsensical"] = np.array([ints["nonsensical"]], dtype=np.uint64)
    path = os.path.join(GOLDEN_DIR, f"fullsize_{meta['key']}.npz")
    np.savez_compressed(path, **arrays)
    return path


def load_fullsize_golden(key):
    """(ints, meta) of tests/golden/fullsize_<key>.npz, or (None, None) when no golden is committed for the workload."""
    path = os.path.join(GOLDEN_DIR, f"fullsize_{key}.npz")
    if not os.path.exists(path):
        return None, None
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    ints = {}
    for k in ["general", "tlen_hist", "gc_hist", "gc_nuc", "gc_rec", "quality"]:
        if k in z:
            ints[k] = z[k]
    if "tlen_scalars" in z:
        ints["tlen_processed"], ints["tlen_ignored"] = int(z["tlen_scalars"][0]), int(z["tlen_scalars"][1])
    if "cov_contigs" in z:
        cov, o = {}, 0
        for i, c in enumerate(z["cov_contigs"]):
            nb = int(z["cov_nbins"][i])
            cov[int(c)] = {"hist": z["cov_hist"][i], "too_large": int(z["cov_too_large"][i]), "bin_sums": z["cov_bins"][o:o + nb]}
            o += nb
        ints["coverage"] = cov
        ints["nonsensical"] = int(z["nonsensical"][0])
    return ints, meta


def compare_fullsize(got, key, n_records=None):
    """Compares a run's integers with the committed full-size golden of workload `key`.
    Returns a one-line verdict; raises AssertionError on any difference."""
    want, meta = load_fullsize_golden(key)
    if want is None:
        return None
    if n_records is not None:
        assert int(n_records) == int(meta["records"]), f"records {n_records} != golden {meta['records']}"
    # a golden computed over contig-exclusive PIECES of a shard file (make_fullsize_goldens.py --split: the file does not fit
    # the build machine's memory) saw the records at other virtual offsets than the run did: its GC window histogram is not
    # comparable (same rule as assert_same_ints(gc_window=False)); every other integer is additive over the pieces
    gc_window = bool(meta.get("gc_window", True))
    assert_same_ints(got, want, records="general" in want, coverage="coverage" in want, gc_window=gc_window)
    return (f"bit-exact vs the oracle's integers on the full workload ({meta['records']} records, {meta['n_ranks']} shard file(s); "
            f"tests/golden/fullsize_{key}.npz)" + ("" if gc_window else " except the GC window histogram, whose golden was computed over "
            "pieces of the file at other virtual offsets (GC invariants checked; the histogram itself is covered by the sample parity)"))
