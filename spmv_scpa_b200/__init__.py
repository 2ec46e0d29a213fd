"""spmv_scpa_b200 -- B200-native FP64 SpMV (CSR, HLL) behind the reference's C API.

The product is two in-tree C libraries (see _lib.py):
  lib/libspmv_b200.so   hand-written sm_100a kernels + the C ABI
  lib/libspmv_host.so   C host layer (loader, packer, CSV logger, generators)
and bin/spmv, the reference-compatible CLI.  This Python package is a thin
ctypes mirror of that interface used by tests/ and bench.py.
"""
from .api import *  # noqa: F401,F403
from .api import (CSR_KERNEL_NAMES, HLL_KERNEL_NAMES, CsrDevice, CsrMatrix, HllDevice, HllMatrix)
from . import _lib, structs  # noqa: F401

__version__ = "0.1"
