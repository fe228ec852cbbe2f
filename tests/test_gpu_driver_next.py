"""The host driver end to end with the two optional facets: the `edits` and
`features` blocks of the results JSON must carry the oracle's integers and the reference's summary arithmetic."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_oracle_edits import _python_edits, make_edits_case, oracle_edits  # noqa: E402
from test_oracle_features import KEYS, features, gff_line  # noqa: E402

EXE = os.path.join(ROOT, "ngs_b200", "ngs-cuda-qc")


def test_results_json_carries_the_optional_facets(tmp_path):
    bam, bai, fa, _, _ = make_edits_case(23)
    (tmp_path / "x.bam").write_bytes(bam)
    (tmp_path / "x.bam.bai").write_bytes(bai)
    (tmp_path / "ref.fa").write_bytes(fa)
    gff = ("##gff-version 3\n" + gff_line("chr1", "gene", 200, 3000) + gff_line("chr1", "exon", 200, 700) + gff_line("chr1", "exon", 1500, 1900, "-")
           + gff_line("chr1", "five_prime_UTR", 200, 260) + gff_line("chr1", "CDS", 261, 700) + gff_line("chr2", "gene", 100, 2500)
           + gff_line("chr2", "three_prime_UTR", 2000, 2500) + gff_line("chrM", "gene", 1, 800))
    (tmp_path / "m.gff").write_text(gff)
    r = subprocess.run([EXE, "qc", str(tmp_path / "x.bam"), "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "out",
                        "--reference-fasta", str(tmp_path / "ref.fa"), "--features-gff", str(tmp_path / "m.gff")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    doc = json.load(open(tmp_path / "out.results.json"))
    one, two, vaf, n, means = oracle_edits(bam, bai, fa)
    e = doc["edits"]
    assert e["read_one_edits"]["values"] == [int(x) for x in one] and e["read_two_edits"]["values"] == [int(x) for x in two]
    assert e["vaf_histogram"]["values"] == [int(x) for x in vaf]
    assert e["summary"] == {"mean_edits_read_one": means[0], "mean_edits_read_two": means[1]}
    want = features(bam, gff)
    f = doc["features"]
    got = [f["exonic_translation_regions"]["utr_five_prime_count"], f["exonic_translation_regions"]["utr_three_prime_count"],
           f["exonic_translation_regions"]["coding_sequence_count"], f["gene_regions"]["intergenic_count"], f["gene_regions"]["exonic_count"],
           f["gene_regions"]["intronic_count"], f["records"]["processed"], f["records"]["ignored_flags"], f["records"]["ignored_nonprimary_chromosome"]]
    assert got == [want[k] for k in KEYS]
    assert (f["summary"]["ignored_flags_pct"], f["summary"]["ignored_nonprimary_chromosome_pct"]) == want["pct"]
    assert doc["general"]["records"]["total"] > 0 and doc["coverage"] is not None   # the default facets ran beside them


def test_vaf_file_equals_the_restated_teardown(tmp_path):
    """`--vaf-file` (edits.rs:108-113, 317-340): every position an `M` of a counted record covered, in header order, with the f32
    VAF printed like Rust's `{}`; the per-position counters come off the device (ngsq_get_edit_positions)."""
    bam, bai, fa, refseqs, recs = make_edits_case(31)
    (tmp_path / "x.bam").write_bytes(bam)
    (tmp_path / "x.bam.bai").write_bytes(bai)
    (tmp_path / "ref.fa").write_bytes(fa)
    vaf_path = tmp_path / "x.vaf.tsv"
    cmd = [EXE, "qc", str(tmp_path / "x.bam"), "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "out",
           "--reference-fasta", str(tmp_path / "ref.fa"), "--vaf-file", str(vaf_path)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = []
    one, two, vaf, n = _python_edits(refseqs, recs, vaf_lines=lines)
    got = vaf_path.read_text().splitlines()
    assert got[0] == "Sequence\tPosition\tVAF" and len(lines) > 1000
    assert got[1:] == lines
    doc = json.load(open(tmp_path / "out.results.json"))
    assert doc["edits"]["vaf_histogram"]["values"] == [int(x) for x in vaf]
    assert sum(doc["edits"]["vaf_histogram"]["values"]) == len(lines)   # one histogram entry per line
    # the reference refuses to overwrite (edits.rs:137-143)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode != 0 and "refusing to overwrite existing VAF file" in r.stderr


def test_num_records_with_reference_fasta(tmp_path):
    """`-n` through the driver with Edits beside the default facets: pass 1 stops after n records, pass 2 follows its own counter."""
    bam, bai, fa, _, _ = make_edits_case(11)
    (tmp_path / "x.bam").write_bytes(bam)
    (tmp_path / "x.bam.bai").write_bytes(bai)
    (tmp_path / "ref.fa").write_bytes(fa)
    r = subprocess.run([EXE, "qc", str(tmp_path / "x.bam"), "GRCh38_no_alt_AnalysisSet", "-o", str(tmp_path), "-p", "out", "-n", "161",
                        "--reference-fasta", str(tmp_path / "ref.fa")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    doc = json.load(open(tmp_path / "out.results.json"))
    one, two, vaf, n, means = oracle_edits(bam, bai, fa, n_records=161)
    e = doc["edits"]
    assert e["read_one_edits"]["values"] == [int(x) for x in one] and e["read_two_edits"]["values"] == [int(x) for x in two]
    assert e["vaf_histogram"]["values"] == [int(x) for x in vaf]
    assert doc["general"]["records"]["total"] == 161
