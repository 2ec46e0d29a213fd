"""Shared fixtures.  `-m gpu` tests need a B200; everything else runs on CPU."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0]
                      for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                      if os.path.basename(p) != "load_errors.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sp():
    import spmv_scpa_b200
    return spmv_scpa_b200


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.port()
    return oracle


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_mtx(name):
    return os.path.join(GOLDEN, name + ".mtx")


@pytest.fixture(scope="session")
def have_ref(O):
    return O.ref_available()


def random_csr(rng, M, N, max_len, empty_frac=0.2, sort=False):
    """Random CSR with ragged rows (unsorted, duplicates allowed like the loader keeps them)."""
    lens = rng.integers(0, max_len + 1, size=M)
    lens[rng.random(M) < empty_frac] = 0
    IRP = np.zeros(M + 1, np.int32)
    IRP[1:] = np.cumsum(lens)
    NZ = int(IRP[-1])
    JA = rng.integers(0, N, size=NZ).astype(np.int32)
    if sort:
        for r in range(M):
            JA[IRP[r]:IRP[r + 1]].sort()
    AS = rng.uniform(-1, 1, size=NZ)
    return IRP, JA, AS
