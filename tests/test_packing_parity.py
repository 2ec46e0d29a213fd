"""PRODUCT host layer vs the reference: loader and packer must be bit-exact.

Checked against the committed reference-run fixtures (tests/golden) and, when
oracle/_ref is present, against the reference's own code on random inputs.
Reads like the reference's own call sequence (src/main.c:78-88).
"""
import errno
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, golden, golden_mtx, random_csr


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_io_load_csr_bit_exact(sp, case):
    g = golden(case)
    A = sp.io_load_csr(golden_mtx(case))
    assert (A.M, A.N, A.NZ) == (int(g["M"]), int(g["N"]), len(g["JA"]))
    assert A.name == str(g["name"])
    assert np.array_equal(A.IRP, g["IRP"])
    assert np.array_equal(A.JA, g["JA"])
    assert np.array_equal(bits(A.AS), bits(g["AS"]))
    for arr in (A.IRP, A.JA, A.AS):
        if arr.size:
            assert arr.ctypes.data % 64 == 0  # ALIGNMENT contract (include/utils.h)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("layout", ["rm", "cm"])
def test_csr_to_hll_bit_exact(sp, case, layout):
    g = golden(case)
    A = sp.io_load_csr(golden_mtx(case))
    H = sp.csr_to_hll(A, layout == "cm")
    assert (H.M, H.N, H.NZ, H.hack_size) == (A.M, A.N, A.NZ, 32)
    assert H.num_blocks == (A.M + 31) // 32 == len(g[f"{layout}_rows"])
    rows, width, nz, off, JA, AS = H.flat()
    assert np.array_equal(rows, g[f"{layout}_rows"])
    assert np.array_equal(width, g[f"{layout}_width"])
    assert np.array_equal(nz, g[f"{layout}_nz"])
    assert np.array_equal(JA, g[f"{layout}_JA"])
    assert np.array_equal(bits(AS), bits(g[f"{layout}_AS"]))
    for b in range(H.num_blocks):
        m, n, _, w, ja, as_ = H.block(b)
        assert n == A.N
        if m * w:
            assert ja.ctypes.data % 64 == 0 and as_.ctypes.data % 64 == 0


def test_loader_error_codes(sp):
    t = np.load(os.path.join(GOLDEN, "load_errors.npz"))
    for name, want in zip(t["names"], t["errnos"]):
        with pytest.raises(OSError) as ei:
            sp.io_load_csr(golden_mtx(str(name)))
        assert ei.value.errno == int(want), name
    with pytest.raises(OSError) as ei:
        sp.io_load_csr("/nonexistent/dir/m.mtx")
    assert ei.value.errno == errno.ENOENT


def test_matrix_name_rule(sp, tmp_path):
    """basename minus '.mtx', at most 63 chars (reference src/csr.c:18-30)."""
    body = "%%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 2.0\n"
    for fname, want in [("abc.mtx", "abc"), ("noext", "noext"), (".mtx", ".mtx"),
                        ("x" * 80 + ".mtx", "x" * 63), ("a.b.mtx", "a.b")]:
        p = tmp_path / fname
        p.write_text(body)
        assert sp.io_load_csr(str(p)).name == want


def test_vec_fill_random_matches_reference_stream(sp):
    """First 64 draws of a fresh process equal the reference's (golden/rand_x64.npy)."""
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys; sys.path.insert(0, %r); import spmv_scpa_b200 as sp; "
            "sys.stdout.write(sp.vec_fill_random(64).tobytes().hex())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, check=True).stdout
    got = np.frombuffer(bytes.fromhex(out.decode()), np.float64)
    assert np.array_equal(got, np.load(os.path.join(GOLDEN, "rand_x64.npy")))


def test_validation_gate(sp):
    a = np.zeros(10)
    b = np.zeros(10)
    b[3] = 0.09
    assert sp.validation_vec_result(a, b) == 0
    b[3] = 0.11
    assert sp.validation_vec_result(a, b) == -1
    assert sp.validation_vec_result(a, np.zeros(9)) == -1


# ---------------------------------------------------------------- live vs _ref --
def test_product_packer_vs_reference_live(sp, O, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for trial in range(25):
        M = int(rng.integers(1, 200))
        N = int(rng.integers(1, 100))
        IRP, JA, AS = random_csr(rng, M, N, int(rng.integers(0, 50)), empty_frac=float(rng.random()))
        A = sp.csr_from_arrays("t", M, N, IRP, JA, AS)
        R = O.RefCsr(M, N, IRP, JA, AS)
        for cm in (False, True):
            want = O.ref_csr_to_hll(R, cm)
            got = sp.csr_to_hll(A, cm).flat()
            for w, g_ in zip(want[:5], got[:5]):
                assert np.array_equal(w, g_)
            assert np.array_equal(bits(want[5]), bits(got[5]))


def test_product_loader_vs_reference_live(sp, O, have_ref, tmp_path):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(17)
    fmts = ["%.17g", "%e", "%.3f", "%g"]
    for trial, (field, sym) in enumerate([("real", "general"), ("real", "symmetric"),
                                          ("pattern", "general"), ("pattern", "symmetric"),
                                          ("real", "skew-symmetric"), ("real", "hermitian")] * 3):
        M = int(rng.integers(1, 80))
        N = M if sym != "general" else int(rng.integers(1, 80))
        nnz = int(rng.integers(0, 300))
        p = tmp_path / f"m{trial}.mtx"
        fmt = fmts[trial % len(fmts)]
        with open(p, "w") as f:
            f.write(f"%%MatrixMarket matrix coordinate {field} {sym}\n%c1\n%c2\n{M} {N} {nnz}\n")
            for _ in range(nnz):
                i, j = int(rng.integers(1, M + 1)), int(rng.integers(1, N + 1))
                if sym != "general" and j > i:
                    i, j = j, i
                sep = "\t" if trial % 2 else " "
                f.write(f"{i}{sep}{j}" + ("" if field == "pattern" else sep + fmt % rng.normal(0, 1e3)) + "\n")
        want = O.ref_load_mtx(str(p))
        A = sp.io_load_csr(str(p))
        assert (A.M, A.N) == (want[0], want[1])
        assert np.array_equal(A.IRP, want[2])
        assert np.array_equal(A.JA, want[3])
        assert np.array_equal(bits(A.AS), bits(want[4]))
        assert A.name == want[5]


# ------------------------------------------------------- large files: parallel parse path --
def _write_big(path, field, sym, M, N, nnz, rng, layout="lines"):
    i = rng.integers(1, M + 1, nnz)
    j = rng.integers(1, N + 1, nnz)
    if sym != "general":
        i, j = np.maximum(i, j), np.minimum(i, j)
    v = rng.normal(0, 1e3, nnz)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {sym}\n% big\n{M} {N} {nnz}\n")
        if layout == "lines":          # one entry per line: the parallel path
            if field == "pattern":
                f.write("".join(f"{a} {b}\n" for a, b in zip(i, j)))
            else:
                f.write("".join(f"{a} {b} {c:.17g}\n" for a, b, c in zip(i, j, v)))
        elif layout == "token_per_line":  # every token on its own line: chunks not entry-aligned
            f.write("".join(f"{a}\n{b}\n{c:.17g}\n" for a, b, c in zip(i, j, v)))
        elif layout == "two_per_line":
            it = iter(zip(i, j, v))
            f.write("".join(f"{a} {b} {c:.17g} {d} {e} {g:.17g}\n"
                            for (a, b, c), (d, e, g) in zip(it, it)))
    return path


@pytest.mark.parametrize("field,sym,layout", [("real", "general", "lines"), ("real", "symmetric", "lines"),
                                              ("pattern", "general", "lines"), ("pattern", "symmetric", "lines"),
                                              ("real", "general", "token_per_line"),
                                              ("real", "general", "two_per_line")])
def test_large_file_parallel_parse_bit_exact(sp, O, tmp_path, field, sym, layout):
    """> 1 MB of entries takes the chunked parallel parser (or its sequential fallback for
    layouts a chunk cannot be aligned to); either way the arrays equal the fscanf walk of the
    oracle port, which is pinned to the reference."""
    rng = np.random.default_rng(99)
    M, N, nnz = 5000, (5000 if sym != "general" else 3000), 120000
    p = _write_big(str(tmp_path / "big.mtx"), field, sym, M, N, nnz, rng, layout)
    assert os.path.getsize(p) > (1 << 20) or field == "pattern"
    want = O.load_mtx(p)
    if O.ref_available():
        ref = O.ref_load_mtx(p)
        for a, b in zip(want[2:5], ref[2:5]):
            assert np.array_equal(a, b)
    A = sp.io_load_csr(p)
    assert (A.M, A.N) == (want[0], want[1])
    assert np.array_equal(A.IRP, want[2])
    assert np.array_equal(A.JA, want[3])
    assert np.array_equal(bits(A.AS), bits(want[4]))


def test_large_file_errors_keep_reference_order(sp, O, tmp_path):
    """An out-of-range index deep inside a big file, and a truncated big file: same errno as
    the sequential reference walk."""
    rng = np.random.default_rng(5)
    p = _write_big(str(tmp_path / "e.mtx"), "real", "general", 4000, 4000, 100000, rng)
    lines = open(p).read().splitlines()
    bad = list(lines)
    bad[70000] = "4001 1 1.0"           # ERANGE at entry ~70000
    bad[90000] = "x y z"                # a later EIO must not win
    q = str(tmp_path / "erange.mtx")
    open(q, "w").write("\n".join(bad) + "\n")
    for loader in (sp.io_load_csr, O.load_mtx):
        with pytest.raises(OSError) as ei:
            loader(q)
        assert ei.value.errno == errno.ERANGE
    q2 = str(tmp_path / "short.mtx")
    open(q2, "w").write("\n".join(lines[:50000]) + "\n")
    for loader in (sp.io_load_csr, O.load_mtx):
        with pytest.raises(OSError) as ei:
            loader(q2)
        assert ei.value.errno == errno.EIO


# ------------------------------------------------------------- property-based cases --
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402


@st.composite
def mtx_files(draw):
    field = draw(st.sampled_from(["real", "pattern"]))
    sym = draw(st.sampled_from(["general", "symmetric", "skew-symmetric"]))
    M = draw(st.integers(1, 70))
    N = M if sym != "general" else draw(st.integers(1, 70))
    nnz = draw(st.integers(0, 120))
    ent = []
    for _ in range(nnz):
        i, j = draw(st.integers(1, M)), draw(st.integers(1, N))
        if sym != "general" and j > i:
            i, j = j, i
        v = draw(st.floats(allow_nan=False, allow_infinity=False, width=64))
        ent.append((i, j, v))
    sep = draw(st.sampled_from([" ", "\t", "  ", " \t "]))
    comments = draw(st.lists(st.sampled_from(["% a comment", "%", "%% odd"]), max_size=3))
    blank = draw(st.booleans())
    case = draw(st.sampled_from([str.lower, str.upper, str.title]))
    lines = [f"%%MatrixMarket {case('matrix')} {case('coordinate')} {case(field)} {case(sym)}"] + comments
    if blank:
        lines.append("")
    lines.append(f"{M}{sep}{N}{sep}{nnz}")
    for i, j, v in ent:
        lines.append(f"{i}{sep}{j}" + ("" if field == "pattern" else f"{sep}{v!r}"))
    return "\n".join(lines) + "\n"


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=mtx_files())
def test_loader_and_packer_properties(sp, O, tmp_path, text):
    """Any well-formed coordinate file: product loader == oracle port (== reference when built),
    and unpacking either HLL layout gives back the CSR rows exactly."""
    p = tmp_path / "h.mtx"
    p.write_text(text)
    want = O.load_mtx(str(p))
    if O.ref_available():
        ref = O.ref_load_mtx(str(p))
        assert np.array_equal(ref[2], want[2]) and np.array_equal(ref[3], want[3])
        assert np.array_equal(bits(ref[4]), bits(want[4]))
    A = sp.io_load_csr(str(p))
    assert (A.M, A.N) == (want[0], want[1])
    assert np.array_equal(A.IRP, want[2]) and np.array_equal(A.JA, want[3])
    assert np.array_equal(bits(A.AS), bits(want[4]))
    for cm in (False, True):
        H = sp.csr_to_hll(A, cm)
        for b in range(H.num_blocks):
            m, n, nzb, w, ja, as_ = H.block(b)
            assert n == A.N and m == min(32, A.M - 32 * b)
            got = 0
            for i in range(m):
                r = 32 * b + i
                k0, k1 = A.IRP[r], A.IRP[r + 1]
                idx = [(j * m + i) if cm else (i * w + j) for j in range(w)]
                row_ja, row_as = ja[idx], as_[idx]
                ln = k1 - k0
                assert np.array_equal(row_ja[:ln], A.JA[k0:k1])
                assert np.array_equal(bits(row_as[:ln]), bits(A.AS[k0:k1]))
                assert (row_ja[ln:] == -1).all() and (row_as[ln:] == 0.0).all()
                got += ln
            assert got == nzb
            assert w == max([A.IRP[32 * b + i + 1] - A.IRP[32 * b + i] for i in range(m)] + [0])


def test_hll_free_handles_both_ownership_layouts(sp, O):
    """hll_free() releases an HLL built here (two slabs) and one built the reference's way (one
    allocation per hack, src/hll.c:60-61) -- valgrind-free by construction: run both in a loop
    and watch the resident set."""
    import ctypes as C
    import resource
    L = sp._lib.host
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    A = sp.gen_stencil27(20, 20, 20)
    S = sp.structs

    def build_reference_style():
        """per-hack allocations through the C allocator, as the reference packer does"""
        src = sp.csr_to_hll(A, True)
        nb = src.num_blocks
        H = C.cast(libc.malloc(C.sizeof(S.sparse_hll)), C.POINTER(S.sparse_hll))
        C.memmove(H, src.ptr, C.sizeof(S.sparse_hll))
        blocks = C.cast(libc.malloc(nb * C.sizeof(S.ellpack_block)), C.POINTER(S.ellpack_block))
        for b in range(nb):
            blk = src.struct.blocks[b]
            n = blk.M * blk.max_NZ
            ja = C.cast(libc.malloc(max(n, 1) * 4), C.POINTER(C.c_int))
            as_ = C.cast(libc.malloc(max(n, 1) * 8), C.POINTER(C.c_double))
            C.memmove(ja, blk.JA, n * 4)
            C.memmove(as_, blk.AS, n * 8)
            blocks[b] = S.ellpack_block(blk.M, blk.N, blk.NZ, blk.max_NZ, ja, as_)
        H.contents.blocks = blocks
        src.free()
        return H

    def rss_mb():
        return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0

    for _ in range(3):
        L.hll_free(C.cast(build_reference_style(), C.c_void_p))
        sp.csr_to_hll(A, False).free()
    base = rss_mb()
    for _ in range(40):                                  # 40 x ~3.5 MB would show as > 100 MB
        L.hll_free(C.cast(build_reference_style(), C.c_void_p))
        sp.csr_to_hll(A, False).free()
    assert rss_mb() - base < 40.0
    # an empty matrix still owns (and releases) its slabs
    E = sp.csr_from_arrays("empty", 0, 5, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    sp.csr_to_hll(E, True).free()
