"""The streamed engine: waves of BGZF blocks through two recycled slots, the open record carried from one wave to the
next, the compressed staging ring, the device-side record counters — every shape of wave must give the oracle's
integers.  Also `-n` with Coverage (the reference's shared second-pass counter, src/qc/command.rs:350-397) and shards of
ONE file cut at BAI-derived virtual offsets (the multi-GPU path of the host driver, on one device)."""
import numpy as np
import pytest

from bamutil import as_u8, rec, write_bam
from helpers import assert_same_ints, engine_ints, merge_ints, oracle_ints

pytestmark = pytest.mark.gpu


def _synth(shape, n, level=6):
    from ngs_b200 import ffi
    return ffi.synth_bam(shape, n, level=level)


@pytest.mark.parametrize("shape,n,launch_blocks,chunk", [
    (1, 40000, 1, None), (1, 40000, 2, 70000), (1, 40000, 3, None), (1, 40000, 17, 300000), (0, 60000, 5, 1 << 20),
    (2, 300, 1, None), (2, 300, 2, 100000), (2, 300, 5, None),       # long reads: a record spans several blocks and several waves
    (3, 30000, 1, 40000), (3, 30000, 4, None)])
def test_any_wave_size_gives_the_oracles_integers(shape, n, launch_blocks, chunk):
    bam, bai, _ = _synth(shape, n)
    want = oracle_ints(bam, bai, gc_seed=5)
    got = engine_ints(bam, gc_seed=5, launch_blocks=launch_blocks, chunk_bytes=chunk)
    assert got["stats"]["records"] == n
    assert got["stats"]["waves"] > 3
    assert_same_ints(got, want)


def test_a_record_longer_than_the_carry_window_fails_cleanly():
    from ngs_b200 import ffi
    bam, bai, _ = _synth(2, 100)
    eng = ffi.Engine(launch_blocks=1, carry_bytes=4096)
    with pytest.raises(ffi.NgsqError) as ei:
        engine_ints(bam, engine=eng)
    assert ei.value.code == -6 and "carry_bytes" in str(ei.value)
    # the default window holds it
    assert engine_ints(bam, launch_blocks=1)["stats"]["records"] == 100


def test_quality_table_capacity_is_an_explicit_error_and_can_be_raised():
    from ngs_b200 import ffi
    bam, bai, _ = _synth(2, 200)
    want = oracle_ints(bam, bai, gc_seed=1)
    flags = ffi.NGSQ_F_RECORD_FACETS | ffi.NGSQ_F_COVERAGE | ffi.NGSQ_F_VERIFY_CRC
    eng = ffi.Engine(flags=flags, gc_seed=1, quality_positions=1000)
    with pytest.raises(ffi.NgsqError) as ei:
        engine_ints(bam, engine=eng)
    assert ei.value.code == -13 and "ngsq_set_quality_positions" in str(ei.value)
    eng.reset()
    eng.set_quality_positions(want["quality"].shape[0])
    got = engine_ints(bam, engine=eng)
    assert_same_ints(got, want)


def test_the_compressed_ring_is_recycled():
    """4 MB of staging for a 10 MB file: segments are reused as soon as their wave has been decoded."""
    from ngs_b200 import ffi
    bam, bai, _ = _synth(1, 80000, level=1)
    assert bam.size > 9 << 20
    want = oracle_ints(bam, bai, gc_seed=9)
    flags = ffi.NGSQ_F_RECORD_FACETS | ffi.NGSQ_F_COVERAGE | ffi.NGSQ_F_VERIFY_CRC
    eng = ffi.Engine(flags=flags, gc_seed=9, comp_ring_bytes=4 << 20, launch_blocks=8)
    got = engine_ints(bam, engine=eng, chunk_bytes=300000)
    assert_same_ints(got, want)
    # and again through the same engine (reset keeps the ring)
    eng.reset()
    assert_same_ints(engine_ints(bam, engine=eng, chunk_bytes=500000), want)


def test_progress_probe_never_blocks_and_ends_on_the_total():
    from ngs_b200 import ffi, formats
    bam, bai, _ = _synth(0, 50000)
    eng = ffi.Engine(launch_blocks=4)
    hdr = formats.read_header(eng, bam)
    eng.set_references([l for _, l in hdr.refs], [1] * len(hdr.refs))
    eng.set_range(hdr.first_voffset, 0)
    seen = [eng.progress()]
    data = np.ascontiguousarray(bam)
    o = 0
    while o < data.size:
        _, _, used = ffi.bgzf_walk(data[o:o + 200000], o)
        used = used or data.size - o
        eng.submit(data[o:o + used], o)
        o += used
        seen.append(eng.progress())
    eng.finish()
    seen.append(eng.progress())
    assert seen == sorted(seen) and seen[0] == 0 and seen[-1] == 50000


# ---------------------------------------------------------------- `-n`
@pytest.mark.parametrize("n_records", [1, 777, 12345, 30000, 10 ** 9])
@pytest.mark.parametrize("launch_blocks", [0, 3])
def test_num_records_applies_the_references_counters_to_both_passes(n_records, launch_blocks):
    """Pass 1: the first n records.  Pass 2: one counter over all contigs that only breaks the current contig's loop
    (command.rs:384-388) — the first n yielded records, then the first yielded record of every later contig."""
    bam, bai, _ = _synth(0, 50000)
    want = oracle_ints(bam, bai, n_records=n_records, gc_seed=3)
    got = engine_ints(bam, n_records=n_records, gc_seed=3, launch_blocks=launch_blocks)
    assert_same_ints(got, want)


def test_num_records_counts_records_of_unsupported_contigs_too():
    """chrM is not in the primary assembly: Coverage skips it, the second pass's counter does not (command.rs:377-388)."""
    refs = [("chr1", 5000), ("chrM", 3000), ("chr2", 4000)]
    r = [rec(name=f"a{i}", flag=0, ref=0, pos=10 + 3 * i, mapq=20, cigar="30M", seq="ACGTAC" * 5, qual=[30] * 30) for i in range(40)]
    r += [rec(name=f"m{i}", flag=0, ref=1, pos=5 + 2 * i, mapq=20, cigar="30M", seq="ACGTAC" * 5, qual=[30] * 30) for i in range(25)]
    r += [rec(name=f"b{i}", flag=0, ref=2, pos=7 + 5 * i, mapq=20, cigar="10M5D20M", seq="ACGTAC" * 5, qual=[30] * 30) for i in range(30)]
    bam, bai = write_bam(refs, r, block_payload=900)
    b, x = as_u8(bam), as_u8(bai)
    for n in (10, 40, 41, 50, 65, 66, 80, 1000):
        want = oracle_ints(b, x, n_records=n)
        got = engine_ints(b, n_records=n, launch_blocks=2)
        assert_same_ints(got, want)


# ---------------------------------------------------------------- shards of one file
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_bai_derived_shards_of_one_file_sum_to_the_whole_file(n_shards):
    """plan_shards cuts ONE BAM at the BAI's contig extents; every shard runs with its (first, end) virtual offsets —
    both usually inside a block — and the sum over shards must be the whole file's integers, GC window histogram
    included (the window offset is keyed by the record's virtual offset, which sharding does not change)."""
    from ngs_b200 import ffi, formats
    bam, bai, _ = _synth(1, 300000)
    want = oracle_ints(bam, bai, gc_seed=7)
    eng = ffi.Engine(gc_seed=7, flags=ffi.NGSQ_F_RECORD_FACETS | ffi.NGSQ_F_COVERAGE | ffi.NGSQ_F_VERIFY_CRC)
    hdr = formats.read_header(eng, bam)
    blocks, n_blocks, _ = ffi.bgzf_walk(bam)
    shards = formats.plan_shards(hdr, formats.parse_bai(bai.tobytes()), n_shards, bam.size)
    assert len(shards) == n_shards and sum(1 for s in shards if s.contigs) >= 2
    parts, records = [], 0
    for s in shards:
        if not s.contigs:
            continue
        lo, hi = formats.shard_byte_range(s, blocks, n_blocks, bam.size)
        eng.reset()
        got = engine_ints(bam, gc_seed=7, shard=(s.first_voffset, s.end_voffset, lo, hi, set(range(len(hdr.refs)))), engine=eng, launch_blocks=0)
        records += got["stats"]["records"]
        parts.append(got)
    assert records == 300000
    assert_same_ints(merge_ints(parts), want, gc_window=True)
    assert any(s.end_voffset & 0xFFFF for s in shards), "no cut fell inside a block: the case is not exercised"


def test_serial_stages_flag_changes_the_schedule_not_the_results():
    """NGSQ_F_SERIAL_STAGES (bench.py's per-kernel timing pass): every kernel of a wave on one stream."""
    from ngs_b200 import ffi
    bam, bai, _ = _synth(1, 60000, level=1)
    want = oracle_ints(bam, bai, gc_seed=4)
    base = ffi.NGSQ_F_RECORD_FACETS | ffi.NGSQ_F_COVERAGE | ffi.NGSQ_F_VERIFY_CRC
    for flags in (base, base | ffi.NGSQ_F_SERIAL_STAGES):
        eng = ffi.Engine(flags=flags, gc_seed=4, launch_blocks=16)
        got = engine_ints(bam, engine=eng, chunk_bytes=1 << 20)
        assert got["stats"]["waves"] >= 5
        assert_same_ints(got, want)
