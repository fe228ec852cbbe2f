"""Host-side logic: genome table, BAM header / BAI parsing, shard planning (no GPU)."""
import gzip

import numpy as np

from bamutil import as_u8, rec, write_bam


def test_primary_assembly_table():  # grch38_no_alt.rs:296-311,332-341: 22 + 2 + 42 + 127 = 193
    from ngs_b200 import formats
    names = formats.grch38_no_alt_names()
    kinds = list(names.values())
    assert kinds.count("C") == 24 and kinds.count("L") == 42 and kinds.count("P") == 127 and kinds.count("M") == 1 and kinds.count("E") == 1
    assert sum(formats.is_primary(n) for n in names) == 193
    assert not formats.is_primary("chrM") and not formats.is_primary("chrEBV") and formats.is_primary("chrUn_KI270302v1")
    assert formats.is_known("chrM") and not formats.is_known("contigX")


def test_oracle_name_rule_equals_table():
    """The oracle decides primary/known by a name rule; it must agree with the table for every built-in name."""
    import ctypes as C
    from helpers import oracle_ints
    from ngs_b200 import formats
    names = formats.grch38_no_alt_names()
    refs = [(n, 1000) for n in names]
    recs = [rec(name=f"r{i}", flag=0, ref=i, pos=5, cigar="10M", seq="A" * 10) for i in range(len(refs))]
    bam, bai = write_bam(refs, recs)
    r = oracle_ints(as_u8(bam), as_u8(bai), records=False)
    touched = sorted(r["coverage"])
    assert touched == [i for i, n in enumerate(names) if formats.is_primary(n)]


def test_header_and_bai_roundtrip():
    from ngs_b200 import formats
    refs = [("chr1", 5000), ("chr2", 3000)]
    recs = [rec(name=f"a{i}", flag=0, ref=0, pos=10 * i, cigar="20M", seq="A" * 20) for i in range(50)]
    recs += [rec(name=f"b{i}", flag=0, ref=1, pos=7 * i, cigar="20M", seq="A" * 20) for i in range(30)]
    recs += [rec(name="u", flag=4)]
    bam, bai = write_bam(refs, recs, block_payload=700)
    plain = gzip.decompress(bam)
    text, prefs, hlen = formats.parse_bam_header(plain)
    assert prefs == refs
    idx = formats.parse_bai(bai)
    assert idx.n_no_coor == 1 and idx.refs[0].n_mapped == 50 and idx.refs[1].n_mapped == 30
    assert idx.refs[0].ref_end == idx.refs[1].ref_beg  # contigs are adjacent in the file
    try:
        formats.parse_bam_header(plain[:20])
        assert False
    except formats.NeedMore:
        pass


def test_plan_shards_contig_aligned():
    from ngs_b200 import ffi, formats
    bam, bai, info = ffi.synth_bam(1, 40000, level=1)
    plain_hdr = gzip.decompress(bam.tobytes()[: 1 << 16] if False else bam.tobytes())[: info["header_bytes"] + 64]
    text, refs, hlen = formats.parse_bam_header(plain_hdr)
    blocks, n, used = ffi.bgzf_walk(bam)
    first = None
    acc = 0
    for i in range(n):
        if hlen < acc + blocks[i].isize:
            first = (blocks[i].coffset << 16) | (hlen - acc)
            break
        acc += blocks[i].isize
    hdr = formats.BamHeader(text, refs, hlen, first)
    idx = formats.parse_bai(bai.tobytes())
    for k in (1, 2, 4, 8):
        shards = formats.plan_shards(hdr, idx, k, bam.size)
        assert len(shards) == k
        owned = sorted(c for s in shards for c in s.contigs)
        assert owned == sorted(set(owned)), "a contig has two owners"
        assert set(owned) >= {c for c, r in enumerate(idx.refs) if r.ref_beg is not None}
        live = [s for s in shards if s.contigs or s.first_voffset]
        assert live[0].first_voffset == hdr.first_voffset and live[-1].end_voffset == 0
        for a, b in zip(live[:-1], live[1:]):
            assert a.end_voffset == b.first_voffset  # contiguous, no record owned twice
            lo, hi = formats.shard_byte_range(a, blocks, n, bam.size)
            assert lo <= (a.end_voffset >> 16) <= hi
