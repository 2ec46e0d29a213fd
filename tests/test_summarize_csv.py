"""tools/summarize_csv.py aggregates the benchmark CSVs exactly as the reference's
scripts/plots.py does (:21-53) -- checked against the reference's own functions where
/root/reference exists (matplotlib, which plots.py imports and this image lacks, is stubbed)."""
import importlib.util
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from conftest import ROOT

pd = pytest.importorskip("pandas")
REF_PLOTS = "/root/reference/scripts/plots.py"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_csvs(d):
    rng = np.random.default_rng(0)
    rows = ["matrix,format,kernel,warps_per_block,rows,cols,nnz,num_blocks,duration_ms,gflops"]
    for run in range(5):                       # results.py -i 5: repeated runs, append mode
        for mat, (M, N, NZ, nb) in {"a": (100, 100, 500, 4), "b": (70, 45, 400, 3)}.items():
            for fmt, nk in (("CSR", 5), ("HLL", 4)):
                for k in range(nk):
                    for w in (2, 4, 8):
                        ms = float(rng.uniform(0.01, 0.2))
                        rows.append(f"{mat},{fmt},{k},{w},{M},{N},{NZ},{'' if fmt == 'CSR' else nb},{ms:.6f},"
                                    f"{2 * NZ / (ms * 1e6):.6f}")
    open(os.path.join(d, "cuda.csv"), "w").write("\n".join(rows) + "\n")
    rows = ["matrix,format,bench,rows,cols,nnz,num_blocks,num_threads,duration_ms,gflops"]
    for run in range(3):
        for t in (2, 4, 7, 16, 29, 40):          # 7 and 29: the nnz split logged fewer threads than asked
            ms = float(rng.uniform(0.1, 1.0))
            rows.append(f"a,CSR,omp_nnz,100,100,500,,{t},{ms:.6f},{1.0 / ms:.6f}")
            rows.append(f"a,HLL,omp_guided,100,100,500,4,{t},{ms:.6f},{1.0 / ms:.6f}")
    open(os.path.join(d, "omp.csv"), "w").write("\n".join(rows) + "\n")
    rows = ["matrix,format,rows,cols,nnz,num_blocks,duration_ms,gflops"]
    for run in range(3):
        rows += [f"a,CSR,100,100,500,,{0.5 + run:.6f},1.000000", f"a,HLL,100,100,500,4,{0.7 + run:.6f},1.000000"]
    open(os.path.join(d, "serial.csv"), "w").write("\n".join(rows) + "\n")


def test_same_aggregation_as_the_reference_plots(tmp_path):
    if not os.path.exists(REF_PLOTS):
        pytest.skip("reference not mounted")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ref = load(REF_PLOTS, "ref_plots")
    ours = load(os.path.join(ROOT, "tools", "summarize_csv.py"), "summarize_csv")
    make_csvs(str(tmp_path))
    cuda, omp, ser = (pd.read_csv(tmp_path / f) for f in ("cuda.csv", "omp.csv", "serial.csv"))
    for a, b, keys in ((ours.aggregate_cuda(cuda), ref.aggregate_cuda(cuda), ["matrix", "format", "kernel", "warps_per_block"]),
                       (ours.aggregate_openmp(omp), ref.aggregate_openmp(omp), ["matrix", "format", "bench", "num_threads"]),
                       (ours.aggregate_serial(ser), ref.aggregate_serial(ser), ["matrix", "format"])):
        a, b = a.sort_values(keys).reset_index(drop=True), b.sort_values(keys).reset_index(drop=True)
        assert len(a) == len(b)
        for c in keys + ["duration_ms", "gflops"]:
            assert (a[c].to_numpy() == b[c].to_numpy()).all(), c
    assert sorted(set(ours.aggregate_openmp(omp)["num_threads"])) == [2, 4, 8, 16, 32, 40]


def test_cli_prints_every_table(tmp_path):
    make_csvs(str(tmp_path))
    open(tmp_path / "b200_dist.csv", "w").write(
        "matrix,gpus,exchange,steps,rows,cols,nnz,ms_per_step,gflops\na,2,push,10,100,100,500,0.010000,0.100000\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_csv.py"), str(tmp_path), "--peak", "6560"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for title in ("== cuda.csv", "== best GPU variant", "== serial.csv ==", "== omp.csv ==", "== b200_dist.csv"):
        assert title in r.stdout
    assert "stream_tma" in r.stdout and "pct_of_peak" in r.stdout
