"""The reference's own known-answer vectors against the PRODUCT's host-side C++ — the
Histogram every derived float comes from (ngs_b200/host/histogram.hpp) and the facet registry
(facets.hpp::get_qc_facets) — not only against the oracle.  No GPU needed: nothing here creates an engine."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kat(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hostkat") / "host_kat")
    lib_dir = os.path.join(ROOT, "ngs_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_kat.cpp"),
                    "-L" + lib_dir, "-lngs_cuda", "-Wl,-rpath," + lib_dir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return {ln.split(" ", 1)[0]: ln.split(" ", 1)[1] for ln in out.splitlines()}


def test_histogram_statistics(kat):  # histogram.rs:414-431
    assert [float(kat[k]) for k in ("mean", "q1", "median", "q3", "iqr")] == [80.0, 75.0, 87.5, 100.0, 25.0]


def test_histogram_medians(kat):  # histogram.rs:434-463
    assert kat["empty_median"] == "None"
    assert [float(kat[f"tie_median_{i}"]) for i in (1, 2, 3)] == [100.0, 150.0, 200.0]


def test_histogram_bounds_and_default(kat):  # histogram.rs:466-481
    assert kat["out_of_bounds"] == "1"
    assert kat["default_range"] == "0 512 513"


def test_histogram_values_and_cumulative_counts(kat):  # histogram.rs:484-523
    assert kat["values"] == "0 1 1 3"
    assert [float(x) for x in kat["normalized"].split()] == [0.0, 0.2, 0.2, 0.6]
    assert kat["bottom"] == "5 8 14 14"
    assert kat["top"] == "0 6 9 14"


def test_facet_registry(kat):  # qc.rs:238-271
    assert kat["default_facets"] == "4 1"
    assert kat["only_gc"] == "1 0 GC Content"
    assert kat["only_coverage_case_insensitive"] == "0 1"
    assert kat["only_unknown"] == "rejected"


def test_gc_histogram_range(kat):  # gc_content.rs:145-151
    assert kat["gc_hist_range"] == "0 100 101"
