"""Pins the oracle port (oracle/oracle.c) to the reference.

(a) against the golden vectors in tests/golden/*.npz, which were produced by
    the reference's own code (tests/golden/make_golden.py);
(b) against the reference compiled in oracle/_ref on fresh random inputs, when
    that library is present (always in the build container, prebuilt on the
    GPU box).
Integer / index / packing work is compared bit for bit; y within the
north-star tolerance |dy_i| <= 1e-12 * sum_j |a_ij x_j| (the reference binary is
-ffast-math, the port is strict IEEE).
"""
import errno
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden, golden_mtx, random_csr

TOL = 1e-12


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_port_loader_matches_golden(O, case):
    g = golden(case)
    M, N, IRP, JA, AS = O.load_mtx(golden_mtx(case))
    assert (M, N) == (int(g["M"]), int(g["N"]))
    assert np.array_equal(IRP, g["IRP"])
    assert np.array_equal(JA, g["JA"])
    assert np.array_equal(AS.view(np.uint64), g["AS"].view(np.uint64))  # bit-exact FP64


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("layout", ["rm", "cm"])
def test_port_packer_matches_golden(O, case, layout):
    g = golden(case)
    rows, width, nz, off, ja, as_ = O.csr_to_hll(int(g["M"]), g["IRP"], g["JA"], g["AS"], layout == "cm")
    assert np.array_equal(rows, g[f"{layout}_rows"])
    assert np.array_equal(width, g[f"{layout}_width"])
    assert np.array_equal(nz, g[f"{layout}_nz"])
    assert np.array_equal(off, g[f"{layout}_off"])
    assert np.array_equal(ja, g[f"{layout}_JA"])
    assert np.array_equal(as_.view(np.uint64), g[f"{layout}_AS"].view(np.uint64))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_port_spmv_matches_golden(O, case):
    g = golden(case)
    M = int(g["M"])
    y = O.csr_spmv(M, g["IRP"], g["JA"], g["AS"], g["x"])
    bound = O.csr_abs_bound(M, g["IRP"], g["JA"], g["AS"], g["x"])
    ok, worst = O.check_tolerance(y, g["y"], bound, TOL)
    assert ok, f"worst ratio {worst}"
    # HLL serial paths of the port agree with CSR too (pads skipped by JA == -1)
    for layout in ("rm", "cm"):
        yh = O.hll_spmv(M, g[f"{layout}_rows"], g[f"{layout}_width"], g[f"{layout}_off"],
                        layout == "cm", g[f"{layout}_JA"], g[f"{layout}_AS"], g["x"])
        ok, worst = O.check_tolerance(yh, g["y"], bound, TOL)
        assert ok, f"{layout}: worst ratio {worst}"


def test_port_loader_errors_match_golden(O):
    t = np.load(os.path.join(os.path.dirname(golden_mtx("x")), "load_errors.npz"))
    for name, want in zip(t["names"], t["errnos"]):
        with pytest.raises(OSError) as ei:
            O.load_mtx(golden_mtx(str(name)))
        assert ei.value.errno == int(want), name
    with pytest.raises(OSError) as ei:
        O.load_mtx("/nonexistent/dir/m.mtx")
    assert ei.value.errno == errno.ENOENT


def test_tolerance_gate_rejects_wrong_answers(O):
    """The checker itself must fail when y is wrong (guards against a vacuous gate)."""
    g = golden("rect_general")
    M = int(g["M"])
    bound = O.csr_abs_bound(M, g["IRP"], g["JA"], g["AS"], g["x"])
    y = g["y"].copy()
    r = int(np.argmax(bound))
    y[r] += 1e-9 * bound[r]
    ok, worst = O.check_tolerance(y, g["y"], bound, TOL)
    assert not ok and worst > 100
    # empty rows: bound 0 -> any nonzero y is an error
    e = golden("tricky_sym")
    be = O.csr_abs_bound(int(e["M"]), e["IRP"], e["JA"], e["AS"], e["x"])
    ye = e["y"].copy()
    ye[int(np.argmin(be))] = 1e-300
    assert not O.check_tolerance(ye, e["y"], be, TOL)[0]


def test_device_pad_convention(O):
    """Pads become the previous column of the row, 0 for an empty row
    (reference src/cuda_hll.cu:173-195)."""
    g = golden("tricky_sym")
    for layout in ("rm", "cm"):
        cm = layout == "cm"
        ja = O.hll_patch_pads(g[f"{layout}_rows"], g[f"{layout}_width"], g[f"{layout}_off"], cm,
                              g[f"{layout}_JA"])
        assert (ja >= 0).all()
        src = g[f"{layout}_JA"]
        assert np.array_equal(ja[src != -1], src[src != -1])
        # y is unchanged by the patch because padded values are 0.0
        rows, width, off = g[f"{layout}_rows"], g[f"{layout}_width"], g[f"{layout}_off"]
        y0 = O.hll_spmv(int(g["M"]), rows, width, off, cm, src, g[f"{layout}_AS"], g["x"])
        y1 = O.hll_spmv(int(g["M"]), rows, width, off, cm, ja, g[f"{layout}_AS"], g["x"])
        assert np.array_equal(y0, y1)
    # row-major: a pad repeats the last real column of ITS row
    rm = O.hll_patch_pads(g["rm_rows"], g["rm_width"], g["rm_off"], False, g["rm_JA"])
    w0 = int(g["rm_width"][0])
    row0 = rm[:w0]
    n0 = int(g["IRP"][1] - g["IRP"][0])
    assert (row0[n0:] == row0[n0 - 1]).all()


def test_partition_rows_rule(O):
    """Greedy nnz-balanced split (reference src/csr.c:218-276)."""
    IRP = np.array([0, 5, 5, 6, 20, 21, 30], np.int32)  # M = 6, nnz = 30
    cut = O.partition_rows(6, IRP, 3)  # target 10: [0,4) has 20>=10 at row 3 -> cut after row 3
    assert cut[0] == 0 and cut[-1] == 6
    assert list(cut) == [0, 4, 6] or list(cut) == [0, 4, 6, 6]
    cut1 = O.partition_rows(6, IRP, 1)
    assert list(cut1) == [0, 6]


# ---------------------------------------------------------------- live vs _ref --
def test_port_vs_reference_live(O, have_ref, tmp_path):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for trial in range(12):
        M = int(rng.integers(1, 150))
        N = int(rng.integers(1, 150))
        IRP, JA, AS = random_csr(rng, M, N, int(rng.integers(0, 40)))
        A = O.RefCsr(M, N, IRP, JA, AS)
        for cm in (False, True):
            want = O.ref_csr_to_hll(A, cm)
            got = O.csr_to_hll(M, IRP, JA, AS, cm)
            for w, g_ in zip(want[:6], got):
                assert np.array_equal(w, g_)
        x = rng.uniform(-1, 1, N)
        _, y_ref = O.ref_csr_serial(A, x)
        y = O.csr_spmv(M, IRP, JA, AS, x)
        ok, worst = O.check_tolerance(y, y_ref, O.csr_abs_bound(M, IRP, JA, AS, x), TOL)
        assert ok, worst


def test_port_loader_vs_reference_live(O, have_ref, tmp_path):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    for trial, (field, sym) in enumerate([("real", "general"), ("real", "symmetric"),
                                          ("pattern", "general"), ("pattern", "symmetric")] * 2):
        M = int(rng.integers(1, 60))
        N = M if sym == "symmetric" else int(rng.integers(1, 60))
        nnz = int(rng.integers(0, 200))
        p = tmp_path / f"t{trial}.mtx"
        with open(p, "w") as f:
            f.write(f"%%MatrixMarket matrix coordinate {field} {sym}\n% c\n{M} {N} {nnz}\n")
            for _ in range(nnz):
                i, j = int(rng.integers(1, M + 1)), int(rng.integers(1, N + 1))
                if sym == "symmetric" and j > i:
                    i, j = j, i
                f.write(f"{i} {j}" + ("" if field == "pattern" else f" {rng.normal():.17g}") + "\n")
        want = O.ref_load_mtx(str(p))
        got = O.load_mtx(str(p))
        assert want[0] == got[0] and want[1] == got[1]
        for w, g_ in zip(want[2:5], got[2:5]):
            assert np.array_equal(w, g_)


def test_golden_rand_x_matches_reference(O, have_ref):
    """x = rand()/RAND_MAX of a fresh process (reference src/vector.c:36-41)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    import subprocess
    import sys
    from conftest import ROOT, GOLDEN
    code = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; import numpy as np; "
            "sys.stdout.write(O.ref_rand_x(64).tobytes().hex())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, check=True).stdout
    got = np.frombuffer(bytes.fromhex(out.decode()), np.float64)
    assert np.array_equal(got, np.load(os.path.join(GOLDEN, "rand_x64.npy")))


def test_reference_hll_cpu_paths_agree_with_its_csr(O, have_ref):
    """The reference's serial HLL SpMV (src/hll.c:127-150) on its own packing equals its serial
    CSR result on the golden inputs (both are the reference; this pins the HLL leg of the CPU
    baseline that bench.py reports)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    for case in ("rect_general", "tricky_sym", "long_row_mixedcase"):
        g = golden(case)
        A = O.RefCsr(int(g["M"]), int(g["N"]), g["IRP"], g["JA"], g["AS"])
        ser_ms, omp_ms, y = O.ref_hll_bench(A, g["x"], 2)
        bound = O.csr_abs_bound(int(g["M"]), g["IRP"], g["JA"], g["AS"], g["x"])
        ok, worst = O.check_tolerance(y, g["y"], bound, TOL)
        assert ok, (case, worst)
        assert ser_ms >= 0 and omp_ms >= 0
