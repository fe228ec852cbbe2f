"""GPU parity for the Genomic Features facet: the nine counters of `ngsq_get_features` must equal
the oracle's on the same BAM + GFF, for distinct and coinciding feature names."""
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

from bamutil import as_u8, write_bam  # noqa: E402
from test_features_model import NAME_SETS  # noqa: E402
from test_oracle_features import KEYS, REFS, _bam_records, features, gff_line, one  # noqa: E402


def engine_features(bam: bytes, gff: str, names, n_records=0):
    """The caller's side of the facet: parse the GFF, file every record under its class, hand the arrays to the engine."""
    from ngs_b200 import ffi, formats
    b = as_u8(bam)
    eng = ffi.Engine(flags=ffi.NGSQ_F_RECORD_FACETS | ffi.NGSQ_F_VERIFY_CRC | ffi.NGSQ_F_FEATURES, max_records=n_records)
    hdr = formats.read_header(eng, b)
    ref_names = [n for n, _ in hdr.refs]
    primary = [1 if formats.is_primary(n) else 0 for n in ref_names]
    eng.set_references([l for _, l in hdr.refs], [0] * len(ref_names))
    slot_class = eng.set_feature_model(names, primary)
    per_ref = {}
    for ln in gff.splitlines():
        if not ln or ln.startswith("#"):
            continue
        f = ln.split("\t")
        if f[0] not in ref_names or not primary[ref_names.index(f[0])] or f[2] not in names:
            continue
        cls = int(slot_class[names.index(f[2])])
        per_ref.setdefault(ref_names.index(f[0]), []).append((int(f[3]), int(f[4]), cls))
    for c, rows in per_ref.items():
        a, s, k = zip(*rows)
        eng.set_features(c, a, s, k)
    eng.set_range(hdr.first_voffset, 0)
    eng.submit(np.ascontiguousarray(b), 0)
    eng.finish()
    return dict(zip(KEYS, (int(x) for x in eng.features())))


@pytest.mark.parametrize("names", NAME_SETS, ids=["-".join(n) for n in NAME_SETS])
def test_features_match_oracle(names):
    from ngs_b200 import ffi
    bam, _, _ = ffi.synth_bam(3, 20000, level=1)
    raw = bam.tobytes()
    rng = random.Random(11)
    spots = [(refs[ref], pos) for refs, ref, pos, flag, _ in _bam_records(raw) if ref >= 0][::5]
    kinds = sorted(set(names)) + ["transcript"]
    lines = []
    for seq, pos in spots:
        for _ in range(4):
            a = max(1, pos + rng.randrange(-3000, 3000))
            lines.append(gff_line(seq, rng.choice(kinds), a, a + rng.choice([0, 1, 2, 50, 400, 5000, 60000]), rng.choice("+-")))
    gff = "##gff-version 3\n" + "".join(lines)
    want = features(raw, gff, names=names)
    assert engine_features(raw, gff, list(names)) == {k: want[k] for k in KEYS}
    want = features(raw, gff, names=names, n_records=777)
    assert engine_features(raw, gff, list(names), n_records=777) == {k: want[k] for k in KEYS}


def test_engine_fails_where_the_oracle_aborts():
    from ngs_b200 import ffi
    gff = gff_line("chr1", "gene", 1, 90000)
    bam, _ = write_bam(REFS, [one(10), one(20, name="*")])
    with pytest.raises(RuntimeError, match="read name"):
        features(bam, gff)
    with pytest.raises(ffi.NgsqError) as ei:
        engine_features(bam, gff, ["five_prime_UTR", "three_prime_UTR", "CDS", "exon", "gene"])
    assert ei.value.code == -12  # NGSQ_E_FEATURES
