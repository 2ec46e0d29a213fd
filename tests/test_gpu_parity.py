"""GPU parity tests (run with -m gpu on the B200 box).

Every kernel behind every reference entry point (include/cuda_csr.h,
include/cuda_hll.h) is called through the C ABI with HOST pointers -- exactly
how the reference's csr.c / hll.c call it -- and its y is compared with the
oracle: the reference's own serial CSR result stored in tests/golden (or run
live from oracle/_ref when present), else the strict-IEEE port.  Tolerance is
the north-star one: |y_i - y_ref_i| <= 1e-12 * sum_j |a_ij x_j| per row.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden, golden_mtx, random_csr

pytestmark = pytest.mark.gpu
TOL = 1e-12
WPB = (2, 4, 8)


def oracle_y(O, M, N, IRP, JA, AS, x):
    """Reference serial CSR when the compiled reference is here, else the port."""
    if O.ref_available():
        _, y = O.ref_csr_serial(O.RefCsr(M, N, IRP, JA, AS), x)
        return y
    return O.csr_spmv(M, IRP, JA, AS, x)


def assert_parity(O, y, M, N, IRP, JA, AS, x, what, y_ref=None):
    if y_ref is None:
        y_ref = oracle_y(O, M, N, IRP, JA, AS, x)
    bound = O.csr_abs_bound(M, IRP, JA, AS, x)
    assert np.isfinite(y).all(), what
    ok, worst = O.check_tolerance(y, y_ref, bound, TOL)
    assert ok, f"{what}: worst |dy|/(1e-12*sum|ax|) = {worst:.3g}"


CSR_BENCHES = ["bench_csr_cuda_thread_row", "bench_csr_cuda_warp_row", "bench_csr_cuda_halfwarp_row",
               "bench_csr_cuda_block_row", "bench_csr_cuda_halfwarp_row_text"]
HLL_BENCHES = [("bench_hll_cuda_threads_row_major", False), ("bench_hll_cuda_threads_col_major", True),
               ("bench_hll_cuda_warp_block", True), ("bench_hll_cuda_halfwarp_row", False)]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_all_entry_points(sp, O, case):
    """Reference call sequence (src/main.c:78-102, :258-359) on the golden inputs."""
    g = golden(case)
    A = sp.io_load_csr(golden_mtx(case))
    x = g["x"]
    args = (int(g["M"]), int(g["N"]), g["IRP"], g["JA"], g["AS"], x)
    for fname in CSR_BENCHES:
        for wpb in WPB:
            ms, gflops, y = getattr(sp, fname)(A, x, wpb)
            assert ms > 0 and y.shape == (A.M,)
            assert_parity(O, y, *args, f"{case}/{fname}/wpb{wpb}", y_ref=g["y"])
    H = {False: sp.csr_to_hll(A, False), True: sp.csr_to_hll(A, True)}
    for fname, cm in HLL_BENCHES:
        for wpb in WPB:
            ms, gflops, y = getattr(sp, fname)(H[cm], x, wpb)
            assert ms > 0 and y.shape == (A.M,)
            assert_parity(O, y, *args, f"{case}/{fname}/wpb{wpb}", y_ref=g["y"])
    sp.release_all()


def _shapes():
    return [
        ("ragged_bins", lambda sp: sp.gen_ragged(5000, 300)),          # every sub-warp bin + warp bin
        ("ragged_long", lambda sp: sp.gen_ragged(3000, 3000)),         # warp + block bins
        ("stencil27", lambda sp: sp.gen_stencil27(24, 20, 18)),
        ("poisson2d", lambda sp: sp.gen_poisson2d(97, 61)),
        ("uniform32", lambda sp: sp.gen_uniform_random(20000, 32)),
        ("uniform_even6", lambda sp: sp.gen_uniform_random(9000, 6)),  # even row length: rotated walk
        ("rmat", lambda sp: sp.gen_rmat(13, 16)),                      # skewed, many empty rows
    ]


@pytest.mark.parametrize("name,make", _shapes())
def test_generated_all_kernels(sp, O, name, make):
    A = make(sp)
    rng = np.random.default_rng(len(name))
    x = rng.uniform(-1, 1, A.N)
    args = (A.M, A.N, A.IRP.copy(), A.JA.copy(), A.AS.copy(), x)
    y_ref = oracle_y(O, *args)
    for fname in CSR_BENCHES:
        for wpb in WPB:
            _, _, y = getattr(sp, fname)(A, x, wpb)
            assert_parity(O, y, *args, f"{name}/{fname}/wpb{wpb}", y_ref=y_ref)
    if name != "ragged_long":  # HLL padding of a 3000-wide ragged matrix is pointless
        H = {False: sp.csr_to_hll(A, False), True: sp.csr_to_hll(A, True)}
        for fname, cm in HLL_BENCHES:
            for wpb in WPB:
                _, _, y = getattr(sp, fname)(H[cm], x, wpb)
                assert_parity(O, y, *args, f"{name}/{fname}/wpb{wpb}", y_ref=y_ref)
    sp.release_all()


def test_huge_rows_split_path(sp, O):
    """Rows beyond 65536 entries take the split + combine kernels; one of 300k."""
    rng = np.random.default_rng(9)
    M, N = 70, 400000
    lens = np.zeros(M, np.int64)
    lens[3], lens[17], lens[40], lens[69] = 300000, 70000, 5000, 1
    IRP = np.zeros(M + 1, np.int32)
    IRP[1:] = np.cumsum(lens)
    JA = rng.integers(0, N, int(IRP[-1])).astype(np.int32)
    AS = rng.uniform(-1, 1, int(IRP[-1]))
    A = sp.csr_from_arrays("huge", M, N, IRP, JA, AS)
    x = rng.uniform(-1, 1, N)
    for fname in ("bench_csr_cuda_halfwarp_row", "bench_csr_cuda_halfwarp_row_text",
                  "bench_csr_cuda_warp_row", "bench_csr_cuda_block_row"):
        _, _, y = getattr(sp, fname)(A, x, 4)
        assert_parity(O, y, M, N, IRP, JA, AS, x, f"huge/{fname}")
    sp.release_all()


def test_random_ragged_matrices(sp, O):
    """Seeded random shapes: unsorted columns, duplicates, empty rows, M % 32 != 0."""
    rng = np.random.default_rng(123)
    for trial in range(10):
        M = int(rng.integers(1, 3000))
        N = int(rng.integers(1, 3000))
        IRP, JA, AS = random_csr(rng, M, N, int(rng.integers(0, 120)), empty_frac=float(rng.random()) * 0.5)
        A = sp.csr_from_arrays(f"r{trial}", M, N, IRP, JA, AS)
        x = rng.uniform(-1, 1, N)
        y_ref = oracle_y(O, M, N, IRP, JA, AS, x)
        for fname in CSR_BENCHES:
            _, _, y = getattr(sp, fname)(A, x, int(rng.choice(WPB)))
            assert_parity(O, y, M, N, IRP, JA, AS, x, f"rand{trial}/{fname}", y_ref=y_ref)
        H = {False: sp.csr_to_hll(A, False), True: sp.csr_to_hll(A, True)}
        for fname, cm in HLL_BENCHES:
            _, _, y = getattr(sp, fname)(H[cm], x, int(rng.choice(WPB)))
            assert_parity(O, y, M, N, IRP, JA, AS, x, f"rand{trial}/{fname}", y_ref=y_ref)
        sp.release_all()


def test_y_is_fully_overwritten_and_cache_sees_new_values(sp, O):
    """The entry points must not depend on y being pre-zeroed, and a matrix that is edited in
    place (same pointers, same shape) must never be served stale from the device cache: with the
    default policy ("hash": the caller's arrays are re-hashed in full on every call) a change of
    ONE value or ONE column index anywhere in the arrays is seen."""
    A = sp.gen_stencil27(40, 40, 40)          # 1.7 M entries: far more than any sampled probe
    x = np.linspace(-1, 1, A.N)
    args = lambda: (A.M, A.N, A.IRP.copy(), A.JA.copy(), A.AS.copy(), x)
    H = sp.csr_to_hll(A, True)
    for fname, mat in (("bench_csr_cuda_halfwarp_row", A), ("bench_csr_cuda_halfwarp_row_text", A)):
        _, _, y0 = getattr(sp, fname)(mat, x, 4)
        assert_parity(O, y0, *args(), "fresh")
        k = 1234567
        row = int(np.searchsorted(A.IRP, k, side="right") - 1)
        old = A.AS[k]
        A.AS[k] = 123.456                     # ONE value
        _, _, y1 = getattr(sp, fname)(mat, x, 4)
        assert_parity(O, y1, *args(), "one value edited in place")
        assert y1[row] != y0[row] and np.array_equal(np.delete(y1, row), np.delete(y0, row))
        oldc = A.JA[k]
        A.JA[k] = (oldc + 777) % A.N          # ONE column index
        _, _, y2 = getattr(sp, fname)(mat, x, 4)
        assert_parity(O, y2, *args(), "one column index edited in place")
        A.AS[k], A.JA[k] = old, oldc
    # HLL: one value inside one hack
    _, _, z0 = sp.bench_hll_cuda_warp_block(H, x, 4)
    blk = H.block(777)
    blk[5][3] = 9.75
    _, _, z1 = sp.bench_hll_cuda_warp_block(H, x, 4)
    assert not np.array_equal(z0, z1)
    Hr, Hw, Hn, Ho, Hj, Ha = H.flat()
    y_h = O.hll_spmv(A.M, Hr, Hw, Ho, True, O.hll_patch_pads(Hr, Hw, Ho, True, Hj), Ha, x)
    bound = O.csr_abs_bound(*args()[:1], *args()[2:5], x) + 10 * np.abs(x).max()
    ok, worst = O.check_tolerance(z1, y_h, bound, TOL)
    assert ok, worst
    sp.release_all()


def test_cache_policies(sp, O):
    """"trust": pointers + shape are the key, spmv_b200_invalidate() after an in-place edit;
    "off": a fresh upload per call (the reference's behaviour)."""
    A = sp.gen_poisson2d(60, 60)
    x = np.linspace(-1, 1, A.N)
    try:
        sp.set_cache_policy("trust")
        _, _, y0 = sp.bench_csr_cuda_halfwarp_row(A, x, 4)
        A.AS[:] *= 3.0
        _, _, y_stale = sp.bench_csr_cuda_halfwarp_row(A, x, 4)
        assert np.array_equal(y_stale, y0)            # by contract: the caller did not invalidate
        sp.invalidate(A)
        _, _, y1 = sp.bench_csr_cuda_halfwarp_row(A, x, 4)
        assert_parity(O, y1, A.M, A.N, A.IRP.copy(), A.JA.copy(), A.AS.copy(), x, "after invalidate")
        sp.set_cache_policy("off")
        A.AS[:] *= 0.5
        c0 = sp.counters()["h2d_bytes"]
        _, _, y2 = sp.bench_csr_cuda_halfwarp_row(A, x, 4)
        assert_parity(O, y2, A.M, A.N, A.IRP.copy(), A.JA.copy(), A.AS.copy(), x, "policy off")
        assert sp.counters()["h2d_bytes"] - c0 >= 12 * A.NZ   # the matrix went up again
    finally:
        sp.set_cache_policy("hash")
        sp.release_all()


def test_timer_c_api(sp):
    """include/cuda_timer.h: init / start / stop / destroy around a device memset."""
    import ctypes as C
    t = sp.structs.cuda_timer()
    L = sp._lib.b200
    assert L.timer_init(C.byref(t)) == 0
    buf = L.spmv_b200_dmalloc(1 << 26)
    L.timer_start(C.byref(t), None)
    L.spmv_b200_dmemset(buf, 1, 1 << 26, None)
    ms = L.timer_stop(C.byref(t), None)
    assert 0.0 < ms < 1000.0
    L.timer_destroy(C.byref(t))
    L.spmv_b200_dfree(buf)


def test_host_buffer_pipeline(sp, O):
    """Host x in / host y out.  Banded matrix: x goes up in column order, row chunks run as their
    columns arrive, y chunks travel back meanwhile; page-locked caller buffers are used in place,
    pageable ones (what the reference's compute_benchmark_csr hands over: posix_memalign vectors,
    src/vector.c:11-20) go through the library's bounce buffers.  Same results as the plain path,
    CSR and HLL entry points, with and without separately timed repetitions; a non-banded matrix
    takes the unchunked pass."""
    import ctypes as C
    L = sp._lib.b200
    dp = C.POINTER(C.c_double)

    for make, banded in ((lambda: sp.gen_stencil27(64, 64, 64), True),
                         (lambda: sp.gen_uniform_random(300000, 16, 5), False)):
        A = make()
        Hc, Hr = sp.csr_to_hll(A, True), sp.csr_to_hll(A, False)
        x_src = np.random.default_rng(8).uniform(-1, 1, A.N)
        IRP, JA, AS = A.IRP.copy(), A.JA.copy(), A.AS.copy()
        y_ref = oracle_y(O, A.M, A.N, IRP, JA, AS, x_src)
        bound = O.csr_abs_bound(A.M, IRP, JA, AS, x_src)
        for pinned in (True, False):
            if pinned:
                xh, yh = sp.pinned_empty(A.N), sp.pinned_empty(A.M)
            else:
                xh, yh = sp.aligned_array(A.N), sp.aligned_array(A.M)
            xh[:] = x_src
            calls = [(L.csr_spmv_cuda_halfwarp_row, A), (L.csr_spmv_cuda_halfwarp_row_text, A),
                     (L.csr_spmv_cuda_warp_row, A), (L.hll_spmv_cuda_warp_block, Hc),
                     (L.hll_spmv_cuda_threads_col_major, Hc), (L.hll_spmv_cuda_halfwarp_row, Hr)]
            for fn, mat in calls:
                for reps in (0, 2):
                    sp.set_timing(1, reps)
                    L.set_csr_warps_per_block(4)
                    L.set_hll_warps_per_block(4)
                    yh[:] = np.nan
                    c0 = sp.counters()
                    ms = fn(mat.ptr, xh.ctypes.data_as(dp), yh.ctypes.data_as(dp), None)
                    c1 = sp.counters()
                    assert ms > 0, sp._lib.last_error()
                    ok, worst = O.check_tolerance(yh, y_ref, bound, TOL)
                    assert ok, (banded, pinned, fn.__name__, reps, worst)
                    assert c1["h2d_bytes"] - c0["h2d_bytes"] >= 8 * A.N     # x went up (plus plans on first use)
                    assert c1["d2h_bytes"] - c0["d2h_bytes"] >= 8 * A.M
            if pinned:
                sp.pinned_free(xh)
                sp.pinned_free(yh)
        sp.set_timing(1, 3)
        sp.release_all()
