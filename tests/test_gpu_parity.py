"""GPU parity: every integer the engine returns through the C ABI must equal the oracle's on the
same seeded synthetic BAM (bit-exact; there is no floating point on the device)."""
import numpy as np
import pytest

from helpers import assert_same_ints, engine_ints, oracle_ints, oracle_lib

pytestmark = pytest.mark.gpu


def _synth(shape, n, level=6, seed=None):
    from ngs_b200 import ffi
    return ffi.synth_bam(shape, n, seed=seed, level=level)


@pytest.mark.parametrize("level", [1, 6, 9, 0])
def test_inflate_bytes_match_zlib(level):
    """K2: inflated bytes identical to zlib's for stored / fast / default / best streams."""
    from ngs_b200 import ffi
    bam, _, info = _synth(0, 20000, level=level)
    eng = ffi.Engine()
    got = eng.inflate_to_host(bam)
    want = np.empty(info["inflated_bytes"], dtype=np.uint8)
    n = oracle_lib().oracle_inflate_all(bam.ctypes.data, bam.size, want.ctypes.data, want.size)
    assert n == info["inflated_bytes"] == got.size
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape,n,level", [(1, 40000, 1), (2, 400, 6), (3, 30000, 1), (3, 30000, 9)])
def test_inflate_bytes_other_shapes(shape, n, level):
    """K2 on the WGS, long-read and RNA-seq shapes (different literal/match mixes and match lengths)."""
    from ngs_b200 import ffi
    bam, _, info = _synth(shape, n, level=level)
    eng = ffi.Engine()
    got = eng.inflate_to_host(bam)
    want = np.empty(info["inflated_bytes"], dtype=np.uint8)
    oracle_lib().oracle_inflate_all(bam.ctypes.data, bam.size, want.ctypes.data, want.size)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape,n", [(0, 200000), (1, 300000), (3, 200000), (2, 3000)])
def test_all_facets_match_oracle(shape, n):
    bam, bai, _ = _synth(shape, n)
    want = oracle_ints(bam, bai, gc_seed=7)
    got = engine_ints(bam, gc_seed=7)
    assert got["stats"]["records"] == n
    assert_same_ints(got, want)


def test_chunked_submit_matches_single_submit():
    bam, bai, _ = _synth(0, 150000)
    want = oracle_ints(bam, bai, gc_seed=1)
    got = engine_ints(bam, gc_seed=1, chunk_bytes=1 << 20, launch_blocks=64)
    assert got["stats"]["inflate_launches"] > 4
    assert_same_ints(got, want)


def test_num_records_limit_record_facets():
    """`-n`: pass 1 processes the first n records in file order (command.rs:312-315)."""
    bam, bai, _ = _synth(0, 50000)
    want = oracle_ints(bam, bai, n_records=12345, coverage=False)
    got = engine_ints(bam, n_records=12345, coverage=False)
    assert_same_ints(got, want, coverage=False)
