"""The `spmv -m <mtx> -o <outdir> [-d]` driver: same flags, same CSV files, same row formats as
the reference driver (src/main.c:28-109, src/logger.c:31-41,:89-153), consumed unmodified by the
reference's scripts/results.py / plots.py."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, golden_mtx

SPMV = os.path.join(ROOT, "bin", "spmv")
REF_ON_B200 = os.path.join(ROOT, "oracle", "_ref", "spmv_ref_on_b200")

HDR = {
    "serial.csv": "matrix,format,rows,cols,nnz,num_blocks,duration_ms,gflops",
    "omp.csv": "matrix,format,bench,rows,cols,nnz,num_blocks,num_threads,duration_ms,gflops",
    "cuda.csv": "matrix,format,kernel,warps_per_block,rows,cols,nnz,num_blocks,duration_ms,gflops",
}
F = r"\d+\.\d{6}"


def run(exe, args, **env):
    e = dict(os.environ, OMP_NUM_THREADS="40", OMP_WAIT_POLICY="passive", **env)
    return subprocess.run([exe] + args, capture_output=True, text=True, env=e, timeout=600)


def read(d, name):
    with open(os.path.join(d, name)) as f:
        return f.read().splitlines()


def check_cpu_csvs(d, name, M, N, NZ, nb):
    s = read(d, "serial.csv")
    assert s[0] == HDR["serial.csv"]
    assert re.fullmatch(rf"{name},CSR,{M},{N},{NZ},,{F},{F}", s[1])
    assert re.fullmatch(rf"{name},HLL,{M},{N},{NZ},{nb},{F},{F}", s[2])
    o = read(d, "omp.csv")
    assert o[0] == HDR["omp.csv"] and len(o) == 1 + 18
    kinds = [l.split(",")[1:3] for l in o[1:]]
    assert kinds == [["CSR", "omp_nnz"]] * 6 + [["CSR", "omp_guided"]] * 6 + [["HLL", "omp_guided"]] * 6
    for l in o[1:7]:
        assert re.fullmatch(rf"{name},CSR,omp_nnz,{M},{N},{NZ},,\d+,{F},{F}", l)
    assert [int(l.split(",")[7]) for l in o[7:13]] == [2, 4, 8, 16, 32, 40]
    for l in o[13:]:
        assert re.fullmatch(rf"{name},HLL,omp_guided,{M},{N},{NZ},{nb},\d+,{F},{F}", l)


def test_usage_and_errors(tmp_path):
    assert run(SPMV, ["-h"]).returncode == 0
    r = run(SPMV, [])
    assert r.returncode == 1 and "Usage:" in r.stderr
    assert run(SPMV, ["-m", golden_mtx("one_hack")]).returncode == 1          # -o missing
    assert run(SPMV, ["-m", golden_mtx("one_hack"), "-o", str(tmp_path), "-b", "x"]).returncode == 1
    # a bad matrix is reported, not a crash (the reference segfaults here: src/main.c:78-85)
    r = run(SPMV, ["-m", golden_mtx("err_short_file"), "-o", str(tmp_path)])
    assert r.returncode == 1 and "Failed to load matrix" in r.stderr and "-5" in r.stderr
    r = run(SPMV, ["-m", "/nonexistent.mtx", "-o", str(tmp_path)])
    assert r.returncode == 1
    assert run(SPMV, ["-m", golden_mtx("one_hack"), "-o", "/nonexistent_dir_xyz"]).returncode == 1


def test_cpu_rows_and_append_mode(tmp_path):
    """serial.csv / omp.csv are complete before the first GPU variant runs; on a box without a
    GPU the run then stops with an error (no CPU fallback for the cuda.csv rows)."""
    import spmv_scpa_b200 as sp
    has_gpu = sp._lib.b200.spmv_b200_device_count() > 0
    out = str(tmp_path)
    r = run(SPMV, ["-m", golden_mtx("rect_general"), "-o", out, "-d"])
    assert r.returncode == (0 if has_gpu else 1), r.stderr
    check_cpu_csvs(out, "rect_general", 70, 45, 400, 3)
    c = read(out, "cuda.csv")
    assert c[0] == HDR["cuda.csv"]
    if not has_gpu:
        assert len(c) == 1 and "no CUDA device" in r.stderr
    # append mode: a second run adds rows, not headers (reference src/logger.c:19-54)
    run(SPMV, ["-m", golden_mtx("rect_general"), "-o", out])
    s = read(out, "serial.csv")
    assert s.count(HDR["serial.csv"]) == 1 and len(s) == 5


@pytest.mark.gpu
def test_full_run_with_validation(tmp_path):
    """All 15 CSR + 12 HLL GPU variants, each validated against the serial CSR result (-d)."""
    out = str(tmp_path)
    for case, (M, N, NZ, nb) in {"rect_general": (70, 45, 400, 3), "poisson6_sym": (36, 36, 156, 2)}.items():
        r = run(SPMV, ["-m", golden_mtx(case), "-o", out, "-d"])
        assert r.returncode == 0, r.stderr
    c = read(out, "cuda.csv")
    assert c[0] == HDR["cuda.csv"] and len(c) == 1 + 2 * 27
    rows = [l.split(",") for l in c[1:28]]
    assert [(r[1], int(r[2]), int(r[3])) for r in rows] == \
        [("CSR", k, w) for k in range(5) for w in (2, 4, 8)] + [("HLL", k, w) for k in range(4) for w in (2, 4, 8)]
    for r in rows:
        assert r[0] == "rect_general" and float(r[8]) > 0 and float(r[9]) >= 0
        assert (r[7] == "") == (r[1] == "CSR")


@pytest.mark.gpu
def test_reference_driver_runs_on_libspmv_b200(tmp_path):
    """Drop-in proof: the reference's own unmodified main.c/csr.c/hll.c, linked against
    libspmv_b200.so instead of its CUDA files (oracle/Makefile), passes its own -d validation."""
    if not os.path.exists(REF_ON_B200):
        pytest.skip("oracle/_ref/spmv_ref_on_b200 not built")
    out = str(tmp_path)
    r = run(REF_ON_B200, ["-m", golden_mtx("long_row_mixedcase"), "-o", out, "-d"])
    assert r.returncode == 0, r.stderr
    c = read(out, "cuda.csv")
    assert c[0] == HDR["cuda.csv"] and len(c) == 28
    check_cpu_csvs(out, "long_row_mixedcase", 100, 100, 220, 4)
    # and our driver writes the same keys for the same matrix
    out2 = str(tmp_path / "ours")
    os.makedirs(out2)
    assert run(SPMV, ["-m", golden_mtx("long_row_mixedcase"), "-o", out2, "-d"]).returncode == 0
    key = lambda l: l.split(",")[:8]
    assert [key(l) for l in read(out2, "cuda.csv")] == [key(l) for l in c]
    assert [l.split(",")[:6] for l in read(out2, "serial.csv")] == [l.split(",")[:6] for l in read(out, "serial.csv")]


@pytest.mark.gpu
def test_multi_gpu_rows_from_the_cli(tmp_path):
    """SPMV_B200_GPUS=N: the iterated SpMV runs behind the C ABI from the C driver (no Python in
    the loop), is validated by -d against the serial CSR result, and logs to its own CSV."""
    import spmv_scpa_b200 as sp
    n = min(sp._lib.b200.spmv_b200_device_count(), 2)
    A = sp.gen_stencil27(20, 18, 16)
    mtx = str(tmp_path / "stencil.mtx")
    sp.gen_write_mtx(A, mtx)
    out = str(tmp_path)
    env = dict(os.environ, SPMV_B200_GPUS=str(n), SPMV_B200_STEPS="6", SPMV_B200_SKIP_CPU="0")
    r = subprocess.run([SPMV, "-m", mtx, "-o", out, "-d"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    rows = read(out, "b200_dist.csv")
    assert rows[0] == "matrix,gpus,exchange,steps,rows,cols,nnz,ms_per_step,gflops" and len(rows) == 2
    f = rows[1].split(",")
    assert f[0] == "stencil" and int(f[1]) == n and f[2] in ("push", "nccl") and int(f[3]) == 6
    assert (int(f[4]), int(f[6])) == (A.M, A.NZ) and float(f[7]) > 0
    assert len(read(out, "cuda.csv")) == 28          # the reference schedule is untouched
