"""Host side of the two "next" facets on CPU (their device paths are written but not yet verified on a GPU, so the
driver still refuses `--features-gff` / `--reference-fasta`): the GFF filing rules of features.rs:287-345 and the
FASTA reader (ngs_b200/host/gene_model.hpp), and the `features` / `edits` blocks of the results JSON with the
reference's summary arithmetic (facets.hpp, results.hpp), fed through the test double of the C ABI."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["five_prime_UTR", "three_prime_UTR", "CDS", "exon", "gene"]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("next") / "host_next_rows")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-o", out, os.path.join(ROOT, "tests", "cpp", "host_next_rows.cpp"),
                    os.path.join(ROOT, "tests", "cpp", "fake_engine.cpp"), "-lz"], check=True)
    return out


def run(exe, *args, env=None):
    r = subprocess.run([exe, *args], capture_output=True, text=True, env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stderr
    return r.stdout


def line(seq, ty, a, b, strand="+"):
    return f"{seq}\tsrc\t{ty}\t{a}\t{b}\t.\t{strand}\t.\tID=x\n"


def test_gff_filing_rules(exe, tmp_path):
    gff = ("##gff-version 3\n#comment\n" + line("chr1", "gene", 100, 900) + line("chr1", "exon", 100, 200, "-") + line("chr1", "transcript", 100, 900)
           + line("chrM", "gene", 1, 10, ".")            # not in the primary assembly: skipped before its strand is parsed
           + line("chrUn_KI270302v1", "CDS", 5, 50) + line("chr2", "five_prime_UTR", 7, 7) + "##FASTA\n>chr1\nthis is not a record\n")
    p = tmp_path / "m.gff"
    p.write_text(gff)
    out = run(exe, "gff", str(p), *NAMES).splitlines()
    assert out[0] == "slot_class 0 1 2 3 4"
    assert sorted(out[1:]) == sorted(["chr1 100 900 4", "chr1 100 200 3", "chrUn_KI270302v1 5 50 2", "chr2 7 7 0"])
    # coinciding names: a feature is filed under the FIRST slot with its name (features.rs:322-331)
    out = run(exe, "gff", str(p), "gene", "three_prime_UTR", "exon", "exon", "gene").splitlines()
    assert out[0] == "slot_class 0 1 2 2 0"
    assert sorted(out[1:]) == sorted(["chr1 100 900 0", "chr1 100 200 2"])
    # gzip input (formats/gff.rs:24-27)
    gz = tmp_path / "m.gff.gz"
    gz.write_bytes(gzip.compress(gff.encode()))
    assert run(exe, "gff", str(gz), *NAMES).splitlines()[0] == "slot_class 0 1 2 3 4"


@pytest.mark.parametrize("text,msg", [
    (line("chr1", "gene", 1, 9, "."), "attempted to parse strand from value: ."),      # features/utils.rs:36-42
    (line("chr1", "transcript", 1, 9, "?"), "attempted to parse strand"),               # ... whatever the type is
    ("chr1\tsrc\tgene\t0\t9\t.\t+\t.\tID=x\n", "invalid GFF record"),                   # positions are 1-based
    ("chr1 gene 1 9\n", "invalid GFF record"),
])
def test_gff_errors_abort(exe, tmp_path, text, msg):
    p = tmp_path / "bad.gff"
    p.write_text(text)
    assert msg in run(exe, "gff", str(p), *NAMES)
    assert "error: opening GFF file" in run(exe, "gff", str(tmp_path / "missing.gff"), *NAMES)


def test_fasta_reader(exe, tmp_path):
    fa = ">chr1 first sequence\nACGTACGTAC\nGTAC\r\nGG\n>chr2\n\nnnnnACGT\n>empty\n"
    p = tmp_path / "r.fa"
    p.write_text(fa, newline="")
    assert run(exe, "fasta", str(p)).splitlines() == ["chr1 16 ACGTACGTACGTACGG", "chr2 8 nnnnACGT", "empty 0 "]
    gz = tmp_path / "r.fa.gz"
    gz.write_bytes(gzip.compress(fa.encode()))
    assert run(exe, "fasta", str(gz)).splitlines()[0] == "chr1 16 ACGTACGTACGTACGG"


def test_features_and_edits_json_blocks(exe, tmp_path):
    rng = np.random.default_rng(4)
    counts = np.array([11, 7, 30, 1000, 250, 400, 1650, 37, 5], dtype=np.uint64)
    one, two, vaf = np.zeros(513, np.uint64), np.zeros(513, np.uint64), np.zeros(101, np.uint64)
    one[:9] = rng.integers(0, 5000, 9)
    one[512] = 2
    two[:6] = rng.integers(0, 5000, 6)
    vaf[[0, 1, 33, 50, 100]] = [90000, 40, 7, 12, 300]
    ints = tmp_path / "next.u64"
    np.concatenate([counts, one, two, vaf, np.array([int(one.sum() + two.sum())], np.uint64)]).tofile(ints)
    doc = json.loads(run(exe, "json", env={"NGSQ_FAKE_NEXT": str(ints)}))
    f = doc["features"]
    assert list(f) == ["exonic_translation_regions", "gene_regions", "records", "summary"]   # struct declaration order
    assert f["exonic_translation_regions"] == {"utr_five_prime_count": 11, "utr_three_prime_count": 7, "coding_sequence_count": 30}
    assert f["gene_regions"] == {"intergenic_count": 1000, "exonic_count": 250, "intronic_count": 400}
    assert f["records"] == {"processed": 1650, "ignored_flags": 37, "ignored_nonprimary_chromosome": 5}
    total = float(37 + 5 + 1650)
    assert f["summary"] == {"ignored_flags_pct": (37.0 / total) * 100.0, "ignored_nonprimary_chromosome_pct": (5.0 / total) * 100.0}
    e = doc["edits"]
    assert list(e) == ["read_one_edits", "read_two_edits", "vaf_histogram", "summary"]
    assert e["read_one_edits"] == {"values": [int(x) for x in one], "range_start": 0, "range_stop": 512}
    assert e["vaf_histogram"] == {"values": [int(x) for x in vaf], "range_start": 0, "range_stop": 100}

    def mean(h):  # Histogram::mean (histogram.rs:258-269): running f64 sums in bin order
        s = d = 0.0
        for i, v in enumerate(h):
            d += float(v)
            s += float(int(v) * i)
        return s / d
    assert e["summary"] == {"mean_edits_read_one": mean(one), "mean_edits_read_two": mean(two)}
    assert doc["general"] is None and doc["coverage"] is None


def test_vaf_file_lines_follow_the_reference(exe, tmp_path):
    """edits.rs:317-340: header line, then `name<TAB>1-based position<TAB>{f32}` for every position with refs + alts > 0, sequences in
    header order; the f32 is printed like Rust's `{}` (shortest round-trip, positional, no trailing `.0`)."""
    fa = tmp_path / "r.fa"
    fa.write_text(">chr1\nACGT\n>chr2\nACGT\n")
    n1, n2 = 11, 6
    refs1, alts1 = np.zeros(n1, np.uint32), np.zeros(n1, np.uint32)
    refs2, alts2 = np.zeros(n2, np.uint32), np.zeros(n2, np.uint32)
    refs1[[1, 2, 3, 4, 5, 10]] = [1, 2, 1, 5, 999999, 0]
    alts1[[1, 2, 3, 4, 5, 10]] = [0, 1, 1, 2, 1, 7]
    refs2[[2, 5]] = [3, 70000]
    alts2[[2, 5]] = [4, 30000]
    with open(tmp_path / "pos.u32", "wb") as f:
        for r, a in ((refs1, alts1), (refs2, alts2)):
            f.write(np.array([r.size], np.uint32).tobytes() + r.tobytes() + a.tobytes())
    one = np.zeros(9 + 513 + 513 + 101 + 1, np.uint64)
    one.tofile(tmp_path / "next.u64")
    out = tmp_path / "x.vaf.tsv"
    env = {"NGSQ_FAKE_VAF": str(tmp_path / "pos.u32"), "NGSQ_FAKE_NEXT": str(tmp_path / "next.u64")}
    assert run(exe, "vaf", str(fa), str(out), "chr1:10", "chr2:5", env=env) == "ok\n"
    want = ["Sequence\tPosition\tVAF",
            "chr1\t1\t0", "chr1\t2\t0.33333334", "chr1\t3\t0.5", "chr1\t4\t0.2857143", "chr1\t5\t0.000001", "chr1\t10\t1",
            "chr2\t2\t0.5714286", "chr2\t5\t0.3"]
    assert out.read_text().splitlines() == want
    # the same strings from numpy's shortest round-trip formatting of the f32 quotient
    for ln in want[1:]:
        name, pos, v = ln.split("\t")
        r, a = (refs1, alts1) if name == "chr1" else (refs2, alts2)
        q = np.float32(a[int(pos)]) / np.float32(int(r[int(pos)]) + int(a[int(pos)]))
        assert v == np.format_float_positional(q, unique=True, trim="-")
    # an existing file is never overwritten (edits.rs:137-143)
    assert "refusing to overwrite existing VAF file" in run(exe, "vaf", str(fa), str(out), "chr1:10", env=env)
