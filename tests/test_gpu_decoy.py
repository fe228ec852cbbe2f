"""Record-boundary scan against a decoy: a BGZF block that BEGINS with bytes that look like three chained BAM records
(they sit inside another record's auxiliary data).  The speculative first-record search of recscan.cuh must take the
decoy, the closure check must catch it (the predecessor's walk lands elsewhere), and the serial fallback
(chain_serial_kernel) must give the oracle's integers."""
import os
import sys
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bamutil import as_u8, rec, record, write_bam  # noqa: E402
from helpers import assert_same_ints, engine_ints, oracle_ints  # noqa: E402

REFS = [("chr1", 100000), ("chr2", 50000)]


def _case():
    fake = b"".join(record(name="d", flag=0, ref=0, pos=5 + i, mapq=30, cigar="10M", seq="ACGTACGTAC", qual=[30] * 10) for i in range(3))
    before = [rec(name=f"a{i}", flag=0x43, ref=0, pos=100 + 7 * i, next_ref=0, next_pos=300, tlen=250, mapq=30, cigar="50M", seq="ACGTA" * 10, qual=[25] * 50) for i in range(40)]
    host = rec(name="host", flag=0x83, ref=0, pos=500, next_ref=0, next_pos=100, tlen=-450, mapq=30, cigar="50M", seq="TTGCA" * 10, qual=[35] * 50, aux=fake + b"\x00" * 7)
    after = [rec(name=f"b{i}", flag=0x43, ref=1, pos=10 + 9 * i, next_ref=0, next_pos=7, mapq=30, cigar="20M5D30M", seq="GGCAT" * 10, qual=[20] * 50) for i in range(60)]
    # the first record block ends exactly where the decoy starts inside `host`
    cut = sum(len(r[0]) for r in before) + len(host[0]) - len(fake) - 7
    assert cut < 0xFF00
    bam, bai = write_bam(REFS, before + [host] + after, block_payload=[cut, 0xFF00])
    return bam, bai, fake


def _blocks(raw):
    off, out = 0, []
    while off < len(raw):
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        out.append(zlib.decompress(raw[off + 18:off + bsize - 8], -15))
        off += bsize
    return out


def test_the_case_really_puts_the_decoy_at_a_block_start():
    bam, _, fake = _case()
    blocks = [b for b in _blocks(bam) if b]
    assert blocks[2].startswith(fake)          # block 0 = header, 1 = records up to the decoy, 2 starts with it


@pytest.mark.gpu
def test_decoy_at_a_block_start_is_caught_and_the_results_are_the_oracles():
    bam, bai, _ = _case()
    b, x = as_u8(bam), as_u8(bai)
    want = oracle_ints(b, x, gc_seed=3)
    got = engine_ints(b, gc_seed=3)
    assert got["stats"]["records"] == 101
    assert_same_ints(got, want)
