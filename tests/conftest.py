import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once (nvcc cross-compiles without a GPU)."""
    need = [os.path.join(ROOT, "ngs_b200", "libngs_cuda.so"), os.path.join(ROOT, "ngs_b200", "libngs_synth.so"),
            os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "ngs_b200", "ngs-cuda-qc")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()
    yield
