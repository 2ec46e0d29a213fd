"""GPU tests of the resident-matrix API (include/spmv_b200.h): device layouts,
device-side generators and converters, row-range launches, the fused push
epilogue, and full-size (BASELINE configs[1]) parity + size-independent
properties."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden, golden_mtx

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float64)).cuda()


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_hll_device_layout_bit_exact(sp, O, torch, case):
    """What sits in HBM is the reference's device convention (src/cuda_hll.cu:173-195):
    column-major, stride 32, pads = previous column / 0 -- from either host layout and
    from the on-GPU CSR->HLL conversion."""
    g = golden(case)
    A = sp.io_load_csr(golden_mtx(case))
    want_hoff, want_ja, want_as = O.hll_device_layout(A.M, g["IRP"], g["JA"], g["AS"])
    built = [sp.HllDevice.from_host(sp.csr_to_hll(A, False)), sp.HllDevice.from_host(sp.csr_to_hll(A, True)),
             sp.CsrDevice.from_host(A).to_hll()]
    for h in built:
        hoff, ja, as_ = h.download()
        assert np.array_equal(hoff, want_hoff)
        assert np.array_equal(ja, want_ja)
        assert np.array_equal(as_.view(np.uint64), want_as.view(np.uint64))
        h.close()


def test_device_stencil_generator_matches_host(sp, torch):
    for (nx, ny, nz, z0, z1) in [(7, 5, 6, 0, 6), (9, 4, 8, 2, 5), (1, 1, 3, 0, 3), (5, 1, 1, 0, 1),
                                 (2, 2, 2, 0, 2), (16, 16, 9, 8, 9)]:
        plane = nx * ny
        want = sp.gen_stencil27_rows(nx, ny, nz, z0 * plane, z1 * plane)
        h = sp.CsrDevice.stencil27(nx, ny, nz, z0, z1)
        irp, ja, as_ = h.download()
        assert np.array_equal(irp, want.IRP.astype(np.int64))
        assert np.array_equal(ja, want.JA)
        assert np.array_equal(as_, want.AS)
        h.close()
    # column offset (shard with a halo): indices shift, nothing else
    nx, ny, nz = 6, 5, 7
    plane = nx * ny
    want = sp.gen_stencil27_rows(nx, ny, nz, 2 * plane, 5 * plane)
    h = sp.CsrDevice.stencil27(nx, ny, nz, 2, 5, col_offset=plane, n_local=5 * plane)
    _, ja, _ = h.download()
    assert np.array_equal(ja, want.JA - plane)
    h.close()


def test_row_range_launch_and_push(sp, O, torch):
    """spmv_rows on declared cut points == the same rows of a full spmv; the push epilogue
    mirrors the chosen rows into a second buffer (stand-in for a peer GPU's halo)."""
    A = sp.gen_stencil27(12, 10, 9)
    plane = 120
    cuts = [plane, A.M - plane]
    h = sp.CsrDevice.from_arrays(A.M, A.N, A.IRP, A.JA, A.AS, cuts=cuts)
    x = dev(torch, np.random.default_rng(0).uniform(-1, 1, A.N))
    y_full = torch.zeros(A.M, dtype=torch.float64, device="cuda")
    for kernel in (2, 4, 0, 1):
        h.spmv(x, y_full, kernel=kernel)
        y = torch.full((A.M,), float("nan"), dtype=torch.float64, device="cuda")
        halo_lo = torch.zeros(plane, dtype=torch.float64, device="cuda")
        halo_hi = torch.zeros(plane, dtype=torch.float64, device="cuda")
        h.spmv(x, y, kernel=kernel, rows=(0, plane), push=[(0, plane, halo_lo.data_ptr())])
        h.spmv(x, y, kernel=kernel, rows=(A.M - plane, A.M),
               push=[(A.M - plane, A.M, halo_hi.data_ptr())])
        h.spmv(x, y, kernel=kernel, rows=(plane, A.M - plane))
        torch.cuda.synchronize()
        assert torch.equal(y, y_full)
        assert torch.equal(halo_lo, y_full[:plane]) and torch.equal(halo_hi, y_full[-plane:])
    with pytest.raises(RuntimeError, match="cut points"):
        h.spmv(x, y_full, rows=(5, 17))
    h.close()


def test_wide_offsets_path(sp, O, torch):
    """64-bit row offsets (shards beyond 2^31 entries) exercised on a small matrix."""
    A = sp.gen_ragged(4000, 200)
    irp64 = A.IRP.astype(np.int64)
    x = np.random.default_rng(1).uniform(-1, 1, A.N)
    y_ref = O.csr_spmv(A.M, A.IRP, A.JA, A.AS, x)
    bound = O.csr_abs_bound(A.M, A.IRP, A.JA, A.AS, x)
    sp.set_knob("force_wide", 1)
    try:
        h = sp.CsrDevice.from_arrays(A.M, A.N, irp64, A.JA, A.AS)
        hh = h.to_hll()
    finally:
        sp.set_knob("force_wide", 0)
    xd = dev(torch, x)
    y = torch.zeros(A.M, dtype=torch.float64, device="cuda")
    for kernel in range(5):
        y.fill_(float("nan"))
        h.spmv(xd, y, kernel=kernel)
        ok, worst = O.check_tolerance(y.cpu().numpy(), y_ref, bound, TOL)
        assert ok, (kernel, worst)
    y.fill_(float("nan"))
    hh.spmv(xd, y, kernel=2)
    ok, worst = O.check_tolerance(y.cpu().numpy(), y_ref, bound, TOL)
    assert ok, ("hll from wide csr", worst)
    irp, ja, as_ = h.download()
    assert np.array_equal(irp, irp64) and np.array_equal(ja, A.JA)
    h.close()
    hh.close()


@pytest.fixture(scope="module")
def c2(sp):
    """BASELINE configs[1]: 3D 27-point stencil 128^3."""
    return sp.gen_stencil27(128, 128, 128)


def test_c2_full_size_parity(sp, O, torch, c2):
    A = c2
    assert (A.M, A.NZ) == (2097152, 55742968)
    x = np.random.default_rng(2).uniform(0, 1, A.N)
    y_ref = O.csr_spmv(A.M, A.IRP, A.JA, A.AS, x)
    if O.ref_available():
        _, y_refref = O.ref_csr_serial(O.RefCsr(A.M, A.N, A.IRP, A.JA, A.AS), x)
    bound = O.csr_abs_bound(A.M, A.IRP, A.JA, A.AS, x)
    h = sp.CsrDevice.from_host(A)
    hh = h.to_hll()
    xd = dev(torch, x)
    y = torch.zeros(A.M, dtype=torch.float64, device="cuda")
    for kernel in (0, 1, 2, 4):
        y.fill_(float("nan"))
        h.spmv(xd, y, kernel=kernel, warps_per_block=8)
        ok, worst = O.check_tolerance(y.cpu().numpy(), y_ref, bound, TOL)
        assert ok, ("csr", kernel, worst)
        if O.ref_available():
            ok, worst = O.check_tolerance(y.cpu().numpy(), y_refref, bound, TOL)
            assert ok, ("csr vs reference binary", kernel, worst)
    for kernel in (1, 2, 3):
        y.fill_(float("nan"))
        hh.spmv(xd, y, kernel=kernel, warps_per_block=4)
        ok, worst = O.check_tolerance(y.cpu().numpy(), y_ref, bound, TOL)
        assert ok, ("hll", kernel, worst)
    # size-independent properties
    ones = torch.ones(A.N, dtype=torch.float64, device="cuda")
    h.spmv(ones, y, kernel=2)
    # interior rows of diag 26 / 26 neighbours -1 sum to 0; every row sum is 26 - (nnz_row - 1)
    rowsum = 27.0 - np.diff(A.IRP)
    assert np.array_equal(y.cpu().numpy(), rowsum)
    # linearity: A(2x + e) == 2 A x + A e, within tolerance
    y1 = torch.zeros_like(y)
    h.spmv(xd, y1, kernel=4)
    y2 = torch.zeros_like(y)
    h.spmv(2 * xd + ones, y2, kernel=4)
    resid = (y2 - (2 * y1 + y)).abs().cpu().numpy()
    assert (resid <= 4e-12 * (2 * bound + 52)).all()
    # CSR and HLL agree bit-for-bit in structure: same x, same y within tolerance
    y3 = torch.zeros_like(y)
    hh.spmv(xd, y3, kernel=2)
    ok, _ = O.check_tolerance(y3.cpu().numpy(), y1.cpu().numpy(), bound, 2 * TOL)
    assert ok
    h.close()
    hh.close()


def test_shard_with_column_offset_from_host_arrays(sp, O, torch):
    """A row slab uploaded with col_offset (how a rank holds its part of A): indices are stored
    relative to the local x slice [c0, c1), the result equals the same rows of the full product."""
    nx, ny, nz = 10, 9, 12
    plane = nx * ny
    full = sp.gen_stencil27(nx, ny, nz)
    x = np.random.default_rng(4).uniform(-1, 1, full.N)
    y_full = O.csr_spmv(full.M, full.IRP, full.JA, full.AS, x)
    bound = O.csr_abs_bound(full.M, full.IRP, full.JA, full.AS, x)
    z0, z1 = 4, 9
    r0, r1 = z0 * plane, z1 * plane
    c0, c1 = (z0 - 1) * plane, (z1 + 1) * plane
    part = sp.gen_stencil27_rows(nx, ny, nz, r0, r1)          # global column indices
    for wide in (0, 1):
        sp.set_knob("force_wide", wide)
        try:
            h = sp.CsrDevice.from_arrays(part.M, c1 - c0, part.IRP, part.JA, part.AS, col_offset=c0,
                                         cuts=[plane, part.M - plane])
        finally:
            sp.set_knob("force_wide", 0)
        _, ja_dev, _ = h.download()
        assert np.array_equal(ja_dev, part.JA - c0)
        xl = dev(torch, x[c0:c1])
        y = torch.full((part.M,), float("nan"), dtype=torch.float64, device="cuda")
        for kernel in (2, 4, 0, 1, 3):
            y.fill_(float("nan"))
            h.spmv(xl, y, kernel=kernel)
            ok, worst = O.check_tolerance(y.cpu().numpy(), y_full[r0:r1], bound[r0:r1], TOL)
            assert ok, (wide, kernel, worst)
        h.close()
