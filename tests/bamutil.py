"""Pure-Python micro-BAM/BAI writer for hand-built test cases (edge shapes the synthetic
generator does not produce: stored / fixed-Huffman / empty blocks, 64 KiB blocks, odd records)."""
import struct
import zlib

import numpy as np

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
OPS = "MIDNSHP=X"
SEQ_CODES = "=ACMGRSVTWYHKDBN"


def bgzf_block(payload: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, extra_subfield=False) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    comp = co.compress(payload) + co.flush()
    extra = b""
    if extra_subfield:  # an unrelated subfield before BC: readers must walk the subfields
        extra = b"XY" + struct.pack("<H", 3) + b"abc"
    xlen = 6 + len(extra)
    bsize = 12 + xlen + len(comp) + 8 - 1
    assert bsize < 65536
    hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", xlen) + extra + b"BC" + struct.pack("<HH", 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def parse_cigar(s):
    if s in ("*", ""):
        return []
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num), OPS.index(ch)))
            num = ""
    return out


def ref_span(cigar):
    return sum(l for l, k in cigar if k in (0, 2, 3, 7, 8))


def record(name="r", flag=0, ref=-1, pos=-1, mapq=0, cigar="*", next_ref=-1, next_pos=-1, tlen=0, seq="", qual=None, aux=b""):
    """pos is 0-based (BAM).  qual: bytes / list of ints, None -> 0xFF * len(seq)."""
    cg = parse_cigar(cigar)
    l_seq = len(seq)
    nm = name.encode() + b"\x00"
    codes = [SEQ_CODES.index(c) for c in seq]
    if len(codes) & 1:
        codes.append(0)
    sq = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    if qual is None:
        ql = b"\xff" * l_seq
    else:
        ql = bytes(qual)
        assert len(ql) == l_seq
    span = ref_span(cg)
    b = reg2bin(pos, pos + (span if span else 1)) if pos >= 0 else 4680
    body = struct.pack("<iiBBHHHIiii", ref, pos, len(nm), mapq, b, len(cg), flag, l_seq, next_ref, next_pos, tlen)
    body += nm + b"".join(struct.pack("<I", (l << 4) | k) for l, k in cg) + sq + ql + aux
    return struct.pack("<I", len(body)) + body


def header_bytes(refs):
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    h = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        h += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    return h


def write_bam(refs, records, block_payload=0xFF00, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, blocker=None, with_eof=True):
    """records: list of (bytes, dict(ref,pos,span,flag)) from rec() below, coordinate-sorted by the caller.
    Returns (bam bytes, bai bytes).  blocker(i, payload) may return a custom BGZF block (bytes)."""
    hdr = header_bytes(refs)
    out = bytearray()
    out += bgzf_block(hdr, level)
    stream = b"".join(r[0] for r in records)
    starts = []
    o = 0
    for r in records:
        starts.append(o)
        o += len(r[0])
    # cut the record stream into payloads
    block_coffs, block_starts = [], []
    p = 0
    i = 0
    sizes = block_payload if isinstance(block_payload, list) else None
    while p < len(stream):
        n = sizes[i % len(sizes)] if sizes else block_payload
        payload = stream[p:p + n]
        block_coffs.append(len(out))
        block_starts.append(p)
        blk = blocker(i, payload) if blocker else None
        out += blk if blk is not None else bgzf_block(payload, level, strategy)
        p += len(payload)
        i += 1
    end_coff = len(out)
    if with_eof:
        out += EOF_BLOCK

    def voff(stream_off):
        if stream_off >= len(stream):
            return end_coff << 16
        k = max(j for j in range(len(block_starts)) if block_starts[j] <= stream_off)
        return (block_coffs[k] << 16) | (stream_off - block_starts[k])

    # BAI
    n_ref = len(refs)
    bins = [dict() for _ in range(n_ref)]
    lin = [dict() for _ in range(n_ref)]
    meta = [dict(beg=None, end=None, m=0, u=0) for _ in range(n_ref)]
    n_no_coor = 0
    for k, (raw, info) in enumerate(records):
        ref, pos, span, flag = info["ref"], info["pos"], info["span"], info["flag"]
        v0, v1 = voff(starts[k]), voff(starts[k] + len(raw))
        if ref < 0:
            n_no_coor += 1
            continue
        end = pos + (span if span else 1)
        b = reg2bin(pos, end)
        ch = bins[ref].setdefault(b, [])
        if ch and ch[-1][1] == v0:
            ch[-1] = (ch[-1][0], v1)
        else:
            ch.append((v0, v1))
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            lin[ref].setdefault(w, v0)
        m = meta[ref]
        if m["beg"] is None:
            m["beg"] = v0
        m["end"] = v1
        if flag & 4:
            m["u"] += 1
        else:
            m["m"] += 1
    bai = bytearray(b"BAI\x01" + struct.pack("<i", n_ref))
    for r in range(n_ref):
        if meta[r]["beg"] is None:
            bai += struct.pack("<ii", 0, 0)
            continue
        bai += struct.pack("<i", len(bins[r]) + 1)
        for b, ch in bins[r].items():
            bai += struct.pack("<Ii", b, len(ch))
            for c in ch:
                bai += struct.pack("<QQ", *c)
        bai += struct.pack("<IiQQQQ", 37450, 2, meta[r]["beg"], meta[r]["end"], meta[r]["m"], meta[r]["u"])
        nw = max(lin[r]) + 1
        vals, last = [], 0
        for w in range(nw):
            last = lin[r].get(w, last)
            vals.append(last)
        bai += struct.pack("<i", nw) + b"".join(struct.pack("<Q", v) for v in vals)
    bai += struct.pack("<Q", n_no_coor)
    return bytes(out), bytes(bai)


def rec(**kw):
    raw = record(**kw)
    cg = parse_cigar(kw.get("cigar", "*"))
    return raw, dict(ref=kw.get("ref", -1), pos=kw.get("pos", -1), span=ref_span(cg), flag=kw.get("flag", 0))


def as_u8(b: bytes) -> np.ndarray:
    return np.frombuffer(b, dtype=np.uint8).copy()
