"""desamba_b200 -- B200-native (CUDA sm_100a) implementation of deSAMBA's classification hot path.

The product is the C-ABI shared library ``desamba_b200/lib/libdesamba_b200.so`` (kernels + ``include/desamba_b200.h``)
and the C driver ``desamba_b200/bin/deSAMBA-b200`` (drop-in for ``deSAMBA classify``).  This Python package is a thin
ctypes binding over that ABI for the test-suite and ``bench.py``; it contains no compute and no CPU fallback: importing
``desamba_b200.api`` fails loudly when the library has not been built.
"""
from .api import (  # noqa: F401
    Index, Context, BatchResult, DsbError, lib, lib_path, HIT_DTYPE, RR_DTYPE, SEED_DTYPE, KERNEL_NAMES, gather_bench, set_sync_mode,
)
