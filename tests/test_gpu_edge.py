"""Edge shapes the reference's decoder accepts or rejects (SURVEY section 4 list): stored / fixed-Huffman /
empty / 64 KiB blocks, extra gzip subfields, records straddling tiny blocks, empty files, odd records;
and the error paths (corrupt DEFLATE, CRC, ISIZE, quality > 93, CIGAR op > 8, truncation)."""
import struct
import zlib

import numpy as np
import pytest

from bamutil import EOF_BLOCK, as_u8, bgzf_block, header_bytes, rec, record, write_bam
from helpers import assert_same_ints, engine_ints, oracle_ints

pytestmark = pytest.mark.gpu
REFS = [("chr1", 100000), ("chr2", 60000), ("chrM", 16569)]


def _records(n=400, seed=1, long_names=False):
    rng = np.random.default_rng(seed)
    out = []
    for ref in (0, 1, 2):
        pos = 0
        for i in range(n // 3):
            pos += int(rng.integers(0, 200))
            L = int(rng.integers(20, 160))
            cig = rng.choice([f"{L}M", f"5S{L-5}M", f"{L-10}M3I7M", f"10M5D{L-10}M", f"{L//2}M300N{L-L//2}M", f"4H{L}M", f"{L}="])
            flag = int(rng.choice([0x63, 0x93, 0x53, 0xA3, 0x400 | 0x63, 0x100 | 0x93, 0x800 | 0x63, 0x41, 0x0, 0x49]))
            seq = "".join(rng.choice(list("ACGTN"), size=L, p=[0.27, 0.22, 0.22, 0.27, 0.02]))
            qual = None if i % 37 == 0 else rng.integers(0, 94, size=L).tolist()
            nref = ref if rng.random() < 0.9 else int(rng.integers(0, 3))
            name = ("read_with_a_rather_long_name_" * 6 + str(i))[:200] if long_names and i % 5 == 0 else f"r{ref}_{i}"
            out.append(rec(name=name, flag=flag, ref=ref, pos=pos, mapq=int(rng.choice([0, 3, 5, 30, 60, 255])), cigar=cig,
                           next_ref=nref, next_pos=pos + 100, tlen=int(rng.integers(-1200, 1200)), seq=seq, qual=qual,
                           aux=b"NMC\x01" if i % 2 else b""))
    out += [rec(name=f"u{i}", flag=0x4 | 0x1 | 0x8 | 0x40, seq="ACGT" * 10, qual=[20] * 40) for i in range(7)]
    out += [rec(name="empty", flag=0x4)]
    return out


def _check(bam, bai, **kw):
    b, i = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, i, gc_seed=3)
    got = engine_ints(b, gc_seed=3, **kw)
    assert_same_ints(got, want)
    return got


@pytest.mark.parametrize("payload", [37, 100, 700, 4096, 0xFF00, 65536, [1, 36, 500, 65536]])
def test_records_straddle_blocks(payload):
    bam, bai = write_bam(REFS, _records(long_names=True), block_payload=payload)
    _check(bam, bai)


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY),
                                             (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (1, zlib.Z_FILTERED)])
def test_deflate_block_kinds(level, strategy):
    bam, bai = write_bam(REFS, _records(seed=2), level=level, strategy=strategy)
    _check(bam, bai)


def test_empty_blocks_extra_subfields_and_multi_member_streams():
    def blocker(i, payload):
        if i % 3 == 1:
            # an unrelated gzip subfield ahead of BC, and an empty block after the real one
            return bgzf_block(payload, extra_subfield=True) + bgzf_block(b"")
        if i % 3 == 2:
            # several deflate blocks inside one BGZF block (Z_FULL_FLUSH points), last one stored
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            third = len(payload) // 3
            comp = co.compress(payload[:third]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(payload[third:2 * third]) + co.flush(zlib.Z_SYNC_FLUSH)
            comp += co.compress(payload[2 * third:]) + co.flush()
            bsize = 18 + len(comp) + 8 - 1
            return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
                    struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))
        return None
    bam, bai = write_bam(REFS, _records(seed=5), block_payload=3000, blocker=blocker)
    _check(bam, bai)


def test_no_eof_marker_and_chunked_submission():
    bam, bai = write_bam(REFS, _records(seed=6), block_payload=2000, with_eof=False)
    _check(bam, bai, chunk_bytes=5000)


def test_header_only_file():
    bam, bai = write_bam(REFS, [])
    got = _check(bam, bai)
    assert got["general"][0] == 0 and got["coverage"] == {} and got["quality"].shape[0] == 0


def test_single_record_and_incompressible_payload():
    rng = np.random.default_rng(0)
    L = 30000
    seq = "".join(rng.choice(list("ACGT"), size=L))
    r = [rec(name="one", flag=0, ref=0, pos=10, cigar=f"{L}M", seq=seq, qual=rng.integers(0, 94, size=L).tolist(), aux=rng.bytes(5000))]
    bam, bai = write_bam(REFS, r)
    got = _check(bam, bai)
    assert got["quality"].shape[0] == L


def test_many_cigar_ops_and_wide_spans():
    ops = "".join(f"{3 + (i % 5)}{'MID=XMNM'[i % 8]}" for i in range(3000))
    ops = ops.replace("N", "M") + "2000N5M"
    from bamutil import parse_cigar
    cg = parse_cigar(ops)
    qlen = sum(l for l, k in cg if k in (0, 1, 4, 7, 8))
    r = [rec(name="cig", flag=0x40, ref=0, pos=100, cigar=ops, seq="A" * qlen, qual=[30] * qlen),
         rec(name="cig2", flag=0x80, ref=0, pos=99000, cigar="50M5000N50M", seq="C" * 100, qual=[11] * 100)]  # overhangs chr1
    bam, bai = write_bam(REFS, r)
    got = _check(bam, bai)
    assert got["nonsensical"] > 0


def _engine_error(bam_bytes, **kw):
    from ngs_b200 import ffi
    with pytest.raises(ffi.NgsqError) as ei:
        engine_ints(as_u8(bam_bytes), **kw)
    return ei.value


def test_corrupt_deflate_is_bad_block():
    bam, bai = write_bam(REFS, _records(seed=7), block_payload=5000)
    b = bytearray(bam)
    first = len(bgzf_block(header_bytes(REFS)))
    for k in range(first + 30, first + 60):
        b[k] ^= 0xA5
    e = _engine_error(bytes(b))
    assert e.code in (-4, -5)  # invalid stream / ISIZE mismatch, or a CRC failure if it still inflates


def test_crc_mismatch_is_reported():
    bam, bai = write_bam(REFS, _records(seed=8), block_payload=5000)
    b = bytearray(bam)
    first = len(bgzf_block(header_bytes(REFS)))
    total = struct.unpack_from("<H", b, first + 16)[0] + 1
    b[first + total - 8] ^= 0xFF  # CRC32 field of the first record block
    assert _engine_error(bytes(b)).code == -5
    # with the check disabled (peak-throughput mode) the same file runs clean
    engine_ints(as_u8(bytes(b)), crc=False)


def test_corrupt_payload_with_valid_deflate_is_a_crc_error_in_every_launch_shape():
    """Bytes changed inside a stored DEFLATE block still inflate, so only the CRC32 check (own stream, one
    launch per wave) can reject them, and it must win over the record-chain errors the garbage causes —
    with one inflate launch and with many (chunked submits, 64-block launches)."""
    payload = 4096
    bam, bai = write_bam(REFS, _records(seed=11), block_payload=payload, level=0)
    b = bytearray(bam)
    first = len(bgzf_block(header_bytes(REFS), 0))
    # third record block: skip 18 bytes of gzip header + 5 of the stored-block header, flip payload bytes
    off = first
    for _ in range(2):
        off += struct.unpack_from("<H", b, off + 16)[0] + 1
    for k in range(off + 18 + 5 + 40, off + 18 + 5 + 48):
        b[k] ^= 0x5A
    assert _engine_error(bytes(b)).code == -5
    assert _engine_error(bytes(b), chunk_bytes=20000, launch_blocks=64).code == -5
    assert _engine_error(bytes(b), chunk_bytes=9000, launch_blocks=2).code == -5


def test_isize_mismatch_is_bad_block():
    bam, bai = write_bam(REFS, _records(seed=9), block_payload=5000)
    b = bytearray(bam)
    first = len(bgzf_block(header_bytes(REFS)))
    total = struct.unpack_from("<H", b, first + 16)[0] + 1
    isize = struct.unpack_from("<I", b, first + total - 4)[0]
    struct.pack_into("<I", b, first + total - 4, isize - 1)
    assert _engine_error(bytes(b), crc=False).code in (-4, -3, -6, -8)


def test_quality_above_93_fails_the_run():
    r = [rec(name="bad", flag=0, ref=0, pos=1, cigar="4M", seq="ACGT", qual=[10, 94, 10, 10])]
    bam, bai = write_bam(REFS, r)
    assert _engine_error(bam).code == -7


def test_cigar_op_above_8_fails_the_run():
    raw = bytearray(record(name="bad", flag=0, ref=0, pos=1, cigar="4M", seq="ACGT", qual=[10] * 4))
    off = 4 + 32 + 4  # cigar word
    struct.pack_into("<I", raw, off, (4 << 4) | 9)
    bam, bai = write_bam(REFS, [(bytes(raw), dict(ref=0, pos=1, span=4, flag=0))])
    assert _engine_error(bam).code == -6


def test_mapped_pair_without_reference_ids_fails_like_the_reference_panics():
    r = [rec(name="x", flag=0x41, ref=0, pos=1, cigar="4M", next_ref=-1, seq="ACGT")]
    bam, bai = write_bam(REFS, r)
    assert _engine_error(bam).code == -6


def test_truncated_input():
    from ngs_b200 import ffi
    bam, bai = write_bam(REFS, _records(seed=10), block_payload=5000, with_eof=False)
    e = _engine_error(bam[:-100])
    assert e.code == -3


def _mixed_length_records(seed, n=260):
    """Reads on both sides of the facet kernel's shared-memory quality tables (152 positions) and of the tile width of
    the pass that tallies the rest (256): some without qualities, some whose only real bytes sit in the tail."""
    rng = np.random.default_rng(seed)
    out, pos = [], 0
    lengths = [100, 151, 152, 153, 200, 250, 407, 408, 409, 663, 664, 665, 1000, 3000, 7001]
    for i in range(n):
        L = lengths[i % len(lengths)] if i % 4 else int(rng.integers(1, 2500))
        pos += int(rng.integers(0, 300))
        seq = "".join(rng.choice(list("ACGT"), size=L))
        kind = i % 11
        if kind == 3:
            qual = None                                   # absent: every byte 0xFF
        elif kind == 7:
            qual = rng.integers(0, 3, size=L).tolist()    # piles on few counters
        else:
            qual = rng.integers(0, 94, size=L).tolist()
        out.append(rec(name=f"m{i}", flag=int(rng.choice([0x0, 0x10, 0x63, 0x93])), ref=0, pos=pos, mapq=60, cigar=f"{L}M",
                       next_ref=0, next_pos=pos + 50, tlen=300, seq=seq, qual=qual))
    return out


@pytest.mark.parametrize("payload,launch_blocks", [(0xFF00, 0), (20000, 3)])
def test_quality_positions_beyond_the_shared_memory_tables(payload, launch_blocks):
    refs = [("chr1", 2_000_000)]
    bam, bai = write_bam(refs, _mixed_length_records(21), block_payload=payload)
    b, i = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, i, gc_seed=3)
    got = engine_ints(b, gc_seed=3, launch_blocks=launch_blocks)
    assert want["quality"].shape[0] == 7001
    assert_same_ints(got, want)


@pytest.mark.parametrize("where", [10, 151, 152, 300, 407, 408, 2999])
def test_quality_above_93_anywhere_in_a_long_read_fails_the_run(where):
    q = [30] * 3000
    q[where] = 94
    r = [rec(name="ok", flag=0, ref=0, pos=1, cigar="200M", seq="A" * 200, qual=[20] * 200),
         rec(name="bad", flag=0, ref=0, pos=5, cigar="3000M", seq="C" * 3000, qual=q)]
    bam, bai = write_bam(REFS, r)
    assert _engine_error(bam).code == -7


@pytest.mark.parametrize("where", [0, 151, 152, 2999])
def test_a_single_real_byte_makes_a_long_quality_string_present_and_its_0xff_bytes_invalid(where):
    """quality_scores.rs:37-49 with noodles' presence rule: all-0xFF means absent; one real byte makes the
    string present, and then every 0xFF in it is a score of 255 > 93."""
    q = [0xFF] * 3000
    q[where] = 40
    r = [rec(name="odd", flag=0, ref=0, pos=5, cigar="3000M", seq="C" * 3000, qual=q)]
    bam, bai = write_bam(REFS, r)
    assert _engine_error(bam).code == -7
    # ... while the all-0xFF string is simply absent: the table stays empty beyond the other read
    r = [rec(name="short", flag=0, ref=0, pos=1, cigar="100M", seq="A" * 100, qual=[20] * 100),
         rec(name="absent", flag=0, ref=0, pos=5, cigar="3000M", seq="C" * 3000, qual=None)]
    bam, bai = write_bam(REFS, r)
    got = _check(bam, bai)
    assert got["quality"].shape[0] == 100


def _pileup_bam(n, span=600):
    """n identical records `1M{span-2}N1M` at one locus of a 5 kb contig (an amplicon seen through spliced two-base reads: the
    depth without the bytes), BGZF blocks and BAI written directly — 8 M records through write_bam's per-record loop take minutes."""
    refs = [("chr1", 5000)]
    one = record(name="a", flag=0, ref=0, pos=1000, mapq=60, cigar=f"1M{span - 2}N1M", seq="AC", qual=[30, 30])
    stream = np.tile(np.frombuffer(one, dtype=np.uint8), n).tobytes()
    out = bytearray(bgzf_block(header_bytes(refs)))
    first = len(out) << 16
    for o in range(0, len(stream), 0xFF00):
        out += bgzf_block(stream[o:o + 0xFF00], 1)
    end = len(out) << 16
    out += EOF_BLOCK
    from bamutil import reg2bin
    bai = bytearray(b"BAI\x01" + struct.pack("<i", 1))
    bai += struct.pack("<i", 2)
    bai += struct.pack("<Ii", reg2bin(1000, 1000 + span), 1) + struct.pack("<QQ", first, end)
    bai += struct.pack("<IiQQQQ", 37450, 2, first, end, n, 0)
    bai += struct.pack("<i", 1) + struct.pack("<Q", first)
    bai += struct.pack("<Q", 0)
    return bytes(out), bytes(bai)


def test_pileup_deeper_than_2_23_keeps_the_carry_of_the_bin_sums():
    """coverage.cuh reduces the per-lane bin sums of a warp's 512 positions as u64: two independent 32-bit halves would drop the
    carry once they pass 2^32 — 8.4 M reads deep over the whole window (amplicon / rRNA depth)."""
    n = 8_500_000
    bam, bai = _pileup_bam(n)
    b, i = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, i, gc_seed=3)
    assert int(want["coverage"][0]["bin_sums"].sum()) == n * 600 > 1 << 32
    got = engine_ints(b, gc_seed=3)
    assert_same_ints(got, want)
