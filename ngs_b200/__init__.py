"""ngs_b200 — B200-native engine for the `ngs qc` BAM hot path.

The product is libngs_cuda.so (C ABI, include/ngs_cuda.h) plus the C++ host driver under
ngs_b200/host; this package only holds the ctypes binding used by tests and bench.py and the
small host-side format helpers (BAM header, BAI, shard planning).
"""
from . import ffi  # noqa: F401

__all__ = ["ffi"]
