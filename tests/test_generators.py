"""Synthetic generators (include/spmv_gen.h): the in-memory CSR equals what io_load_csr
produces from the .mtx written by gen_write_mtx, and the shapes are the ones BASELINE.json names."""
import numpy as np
import pytest


def same_as_roundtrip(sp, A, tmp_path):
    p = str(tmp_path / (A.name + ".mtx"))
    sp.gen_write_mtx(A, p)
    B = sp.io_load_csr(p)
    assert (A.M, A.N, A.NZ, A.name) == (B.M, B.N, B.NZ, B.name)
    assert np.array_equal(A.IRP, B.IRP) and np.array_equal(A.JA, B.JA)
    assert np.array_equal(A.AS.view(np.uint64), B.AS.view(np.uint64))


def test_poisson2d(sp, tmp_path):
    A = sp.gen_poisson2d(7, 5)
    assert A.NZ == 5 * 35 - 2 * 7 - 2 * 5
    same_as_roundtrip(sp, A, tmp_path)
    y = np.add.reduceat(A.AS, A.IRP[:-1])
    assert y.min() >= 0 and y[2 * 7 + 3] == 0      # interior rows sum to zero
    big = sp.gen_poisson2d(1000, 1000)             # BASELINE configs[0]
    assert (big.M, big.NZ) == (1000000, 4996000)


def test_stencil27(sp, tmp_path):
    for dims in [(4, 3, 5), (1, 1, 1), (2, 2, 2), (6, 1, 3)]:
        A = sp.gen_stencil27(*dims)
        nnz = 1
        for n in dims:
            nnz *= 3 * n - 2
        assert A.NZ == nnz
        for r in range(A.M):
            cols = A.JA[A.IRP[r]:A.IRP[r + 1]]
            assert (np.diff(cols) > 0).all()
        same_as_roundtrip(sp, A, tmp_path)
    # rows of a slab == the same rows of the full matrix (columns stay global)
    full = sp.gen_stencil27(5, 4, 6)
    part = sp.gen_stencil27_rows(5, 4, 6, 40, 100)
    k0, k1 = full.IRP[40], full.IRP[100]
    assert np.array_equal(part.IRP, full.IRP[40:101] - k0)
    assert np.array_equal(part.JA, full.JA[k0:k1]) and np.array_equal(part.AS, full.AS[k0:k1])
    A = sp.gen_stencil27(128, 128, 128)            # BASELINE configs[1]
    assert (A.M, A.NZ) == (2097152, 382 ** 3)


def test_uniform_random(sp, tmp_path):
    A = sp.gen_uniform_random(500, 32, 42)
    assert A.NZ == 500 * 32
    ja = A.JA.reshape(500, 32)
    assert (np.diff(ja, axis=1) > 0).all()          # 32 DISTINCT sorted columns per row
    assert ja.min() >= 0 and ja.max() < 500
    assert (np.abs(A.AS) < 1).all()
    same_as_roundtrip(sp, A, tmp_path)
    B = sp.gen_uniform_random(500, 32, 42)
    assert np.array_equal(A.JA, B.JA) and np.array_equal(A.AS, B.AS)   # deterministic
    C = sp.gen_uniform_random(500, 32, 43)
    assert not np.array_equal(A.JA, C.JA)


def test_rmat(sp, tmp_path):
    A = sp.gen_rmat(10, 16)
    assert (A.M, A.NZ) == (1024, 16384)
    lens = np.diff(A.IRP)
    assert lens.max() > 20 * lens.mean()            # skewed
    assert (lens == 0).any()
    same_as_roundtrip(sp, A, tmp_path)


def test_ragged_covers_all_bins(sp, tmp_path):
    A = sp.gen_ragged(4000, 300)
    lens = np.diff(A.IRP)
    for lo, hi in [(0, 0), (1, 4), (5, 8), (9, 16), (17, 32), (33, 64), (65, 300)]:
        assert ((lens >= lo) & (lens <= hi)).any()
    same_as_roundtrip(sp, A, tmp_path)


def test_too_large_is_refused(sp):
    with pytest.raises(MemoryError):
        sp.gen_stencil27(512, 512, 512)             # nnz > INT_MAX: not representable as sparse_csr


def test_oracle_side_generators_equal_the_product_generators(sp, O):
    """bench.py --impl reference builds its inputs with the generators restated in oracle/oracle.c
    (it must not map a product library): both sides produce identical arrays."""
    pairs = [(lambda: O.gen_stencil27(7, 5, 6), lambda: sp.gen_stencil27(7, 5, 6)),
             (lambda: O.gen_stencil27_rows(9, 4, 8, 72, 180), lambda: sp.gen_stencil27_rows(9, 4, 8, 72, 180)),
             (lambda: O.gen_poisson2d(13, 7), lambda: sp.gen_poisson2d(13, 7)),
             (lambda: O.gen_uniform_rows(900, 9, 42, 0, 900), lambda: sp.gen_uniform_random(900, 9, 42)),
             (lambda: O.gen_rmat(10, 8), lambda: sp.gen_rmat(10, 8))]
    for mk_o, mk_p in pairs:
        M, N, IRP, JA, AS = mk_o()
        A = mk_p()
        assert (M, N) == (A.M, A.N)
        assert np.array_equal(IRP, A.IRP) and np.array_equal(JA, A.JA) and np.array_equal(AS, A.AS)
    # a row range of the uniform matrix regenerates independently
    M, N, IRP, JA, AS = O.gen_uniform_rows(900, 9, 42, 300, 500)
    A = sp.gen_uniform_random(900, 9, 42)
    assert np.array_equal(JA, A.JA[300 * 9:500 * 9]) and np.array_equal(AS, A.AS[300 * 9:500 * 9])
