"""The C-ABI libraries load on a CPU-only box and export every symbol include/*.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

HEADERS_B200 = ["cuda_csr.h", "cuda_hll.h", "cuda_timer.h", "spmv_b200.h"]
HEADERS_HOST = ["csr.h", "hll.h", "vector.h", "utils.h", "logger.h", "mmio.h", "spmv_gen.h", "spmv_errptr.h"]

def declared_functions(header):
    """Functions a header declares itself (macro-generated prototypes included): run the C
    preprocessor and keep the text that the line markers attribute to that header."""
    import subprocess
    path = os.path.join(ROOT, "include", header)
    out = subprocess.run(["/usr/bin/gcc", "-E", "-I" + os.path.join(ROOT, "include"), path],
                         capture_output=True, text=True, check=True).stdout
    keep, mine = [], False
    for line in out.splitlines():
        m = re.match(r'# \d+ "([^"]+)"', line)
        if m:
            mine = os.path.abspath(m.group(1)) == os.path.abspath(path)
            continue
        if mine:
            keep.append(line)
    text = "\n".join(keep)
    text = re.sub(r"static\s+inline[^{;]*\{.*?\n\}", "", text, flags=re.S)  # inline bodies
    names = set()
    for m in re.finditer(r"([\w\*\s]+?)\b(\w+)\s*\(([^;{()]*)\)\s*(?:__attribute__\s*\(\(.*?\)\))?\s*;", text, flags=re.S):
        if m.group(2) not in ("while", "if", "for", "sizeof", "return", "__attribute__", "format"):
            names.add(m.group(2))
    return sorted(names)


def test_struct_sizes(sp):
    S = sp.structs
    # reference ABI (sizeof(sparse_csr)=104, ellpack_block=32, sparse_hll=96)
    assert C.sizeof(S.sparse_csr) == 104
    assert C.sizeof(S.ellpack_block) == 32
    assert C.sizeof(S.sparse_hll) == 96
    assert C.sizeof(S.vec) == 16
    assert C.sizeof(S.bench) == 32
    assert C.sizeof(S.bench_cuda) == 40
    assert C.sizeof(S.cuda_timer) == 16


@pytest.mark.parametrize("header", HEADERS_B200)
def test_libspmv_b200_exports(sp, header):
    names = declared_functions(header)
    assert names, header
    # csr.h/hll.h bench_* live in the host library, not in libspmv_b200
    for n in names:
        if n.startswith(("bench_", "io_load", "csr_free", "csr_to_hll", "hll_free", "init_")):
            continue
        assert hasattr(sp._lib.b200, n), f"libspmv_b200.so does not export {n} ({header})"


@pytest.mark.parametrize("header", HEADERS_HOST)
def test_libspmv_host_exports(sp, header):
    for n in declared_functions(header):
        if n.startswith(("init_", "spmv_err_ptr", "spmv_ptr_err", "spmv_is_err", "now", "compute_gflops")):
            continue  # static inline
        assert hasattr(sp._lib.host, n), f"libspmv_host.so does not export {n} ({header})"


def test_reference_boundary_symbols(sp):
    """Exactly the 11 names the reference's csr.c / hll.c bind (reference include/cuda_csr.h:10-25,
    include/cuda_hll.h:10-22)."""
    want = ["set_csr_warps_per_block", "set_hll_warps_per_block"] + list(
        sp._lib.CSR_ENTRY_POINTS) + list(sp._lib.HLL_ENTRY_POINTS)
    assert len(want) == 11
    for n in want:
        assert hasattr(sp._lib.b200, n)


def test_no_gpu_fails_loudly(sp):
    """On a box without a GPU the product must refuse, not fall back to a CPU path."""
    import numpy as np
    if sp._lib.b200.spmv_b200_device_count() > 0:
        pytest.skip("GPU present")
    A = sp.gen_poisson2d(8, 8)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        sp.bench_csr_cuda_halfwarp_row(A, np.ones(A.N))
    with pytest.raises(RuntimeError):
        sp.CsrDevice.from_host(A)


def test_product_never_touches_oracle():
    """Nothing under the product package may import, link or open oracle/."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "spmv_scpa_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                for line in txt.splitlines():
                    s = line.strip()
                    if ("import" in s or "#include" in s or "CDLL" in s or "dlopen" in s) and "oracle" in s:
                        bad.append((f, s))
    assert not bad, bad
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "liboracle" not in mk.replace("$(MAKE) -C oracle", "")


def test_host_copy_pool(sp):
    """The copy-thread pool behind the bounce buffers of the host-pointer calls: every size from a
    few bytes to several pieces, odd offsets, back-to-back jobs of shrinking and growing size."""
    import ctypes as C
    import time
    import numpy as np
    L = sp._lib.b200
    rng = np.random.default_rng(0)
    src = rng.integers(0, 255, 48 << 20, dtype=np.uint8)
    dst = np.zeros_like(src)
    sizes = [0, 1, 7, 4096, (256 << 10) - 1, (256 << 10) * 2 + 1, (256 << 10) * 3, 5 << 20, (8 << 20) + 13, 40 << 20,
             (256 << 10) * 3 + 5, 17 << 20, 1 << 20]
    for rep in range(3):
        for n in sizes:
            o_s, o_d = int(rng.integers(0, 4096)), int(rng.integers(0, 4096))
            dst[:] = 0
            assert L.spmv_b200_host_copy(C.c_void_p(dst.ctypes.data + o_d), C.c_void_p(src.ctypes.data + o_s), n) == 0
            assert np.array_equal(dst[o_d:o_d + n], src[o_s:o_s + n])
            assert not dst[:o_d].any() and not dst[o_d + n:].any()
    t0 = time.perf_counter()
    for _ in range(5):
        L.spmv_b200_host_copy(C.c_void_p(dst.ctypes.data), C.c_void_p(src.ctypes.data), src.nbytes)
    gbs = 5 * src.nbytes / (time.perf_counter() - t0) / 1e9
    assert gbs > 1.0


def test_host_copy_pool_with_a_pinned_caller():
    """OMP_PROC_BIND (set by bench.py for the CPU baseline) binds the initial thread to one CPU and
    threads created later inherit that mask; the copy team takes the CPUs the container allows
    instead, so a pinned caller must not put the whole team on one core (measured on the bench box
    before the fix: 3.5 s per 2 GiB pass)."""
    import subprocess
    import sys
    code = r'''
import os, ctypes as C, time, numpy as np
os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
import spmv_scpa_b200 as sp
L = sp._lib.b200
src = np.random.default_rng(0).integers(0, 255, 64 << 20, dtype=np.uint8); dst = np.zeros_like(src)
L.spmv_b200_host_copy(C.c_void_p(dst.ctypes.data), C.c_void_p(src.ctypes.data), src.nbytes)
t0 = time.perf_counter()
for _ in range(4):
    L.spmv_b200_host_copy(C.c_void_p(dst.ctypes.data), C.c_void_p(src.ctypes.data), src.nbytes)
gbs = 4 * src.nbytes / (time.perf_counter() - t0) / 1e9
assert np.array_equal(src, dst)
print("GBS", gbs)
'''
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    gbs = float(out.stdout.split("GBS")[1])
    assert gbs > 1.0, gbs
