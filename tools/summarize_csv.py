#!/usr/bin/env python
"""Matplotlib-free summary of the benchmark CSVs (serial.csv / omp.csv / cuda.csv).

Same aggregation as the reference's scripts/plots.py (:21-53: median over repeated runs per
(matrix, format, kernel, warps_per_block) / (matrix, format, bench, num_threads)), printed as
tables, plus the columns the reference does not have: achieved GB/s on the minimum-traffic byte
count B_min = 12 nnz + 4 (rows+1) + 8 rows + 8 cols and its fraction of the HBM peak.

    python tools/summarize_csv.py <results-dir> [--peak GB/s]
"""
import argparse
import json
import os

import pandas as pd

CSR_KERNELS = ["thread_row", "warp_row", "adaptive", "block_row", "stream_tma"]
HLL_KERNELS = ["thread_row_rm", "thread_row", "warp_hack", "stream_tma"]


VALID_THREADS = [2, 4, 8, 16, 32, 40]       # reference scripts/plots.py:10


def bmin(df):
    return 12 * df["nnz"] + 4 * (df["rows"] + 1) + 8 * df["rows"] + 8 * df["cols"]


def aggregate_cuda(df):
    """scripts/plots.py:32-40: median per (matrix, format, kernel, warps_per_block)."""
    return df.groupby(["matrix", "format", "kernel", "warps_per_block"], as_index=False).agg(
        rows=("rows", "first"), cols=("cols", "first"), nnz=("nnz", "first"),
        duration_ms=("duration_ms", "median"), gflops=("gflops", "median"), runs=("gflops", "size"))


def aggregate_serial(df):
    """scripts/plots.py:21-29."""
    return df.groupby(["matrix", "format"], as_index=False).agg(duration_ms=("duration_ms", "median"),
                                                                gflops=("gflops", "median"))


def aggregate_openmp(df):
    """scripts/plots.py:43-53: thread counts are first rounded UP to the next value of the
    reference's fixed list (the nnz-balanced split may log fewer threads than it was given)."""
    df = df.copy()
    df["num_threads"] = [next((t for t in VALID_THREADS if n <= t), VALID_THREADS[-1]) for n in df["num_threads"]]
    return df.groupby(["matrix", "format", "bench", "num_threads"], as_index=False).agg(
        duration_ms=("duration_ms", "median"), gflops=("gflops", "median"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--peak", type=float, default=None, help="HBM GB/s (default: MEASURED_PEAKS.json or 6650)")
    a = ap.parse_args()
    peak = a.peak
    if peak is None:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            peak = 6650.0
    pd.set_option("display.width", 200)
    p = os.path.join(a.dir, "cuda.csv")
    if os.path.exists(p):
        df = pd.read_csv(p)
        g = aggregate_cuda(df)
        g["kernel_name"] = [(CSR_KERNELS if f == "CSR" else HLL_KERNELS)[int(k)] for f, k in zip(g["format"], g["kernel"])]
        g["GBs"] = bmin(g) / (g["duration_ms"] * 1e6)
        g["pct_of_peak"] = 100 * g["GBs"] / peak
        print(f"== cuda.csv (median per variant; HBM peak {peak:.0f} GB/s) ==")
        print(g[["matrix", "format", "kernel", "kernel_name", "warps_per_block", "runs", "duration_ms", "gflops",
                 "GBs", "pct_of_peak"]].to_string(index=False, float_format=lambda v: f"{v:.4g}"))
        best = g.loc[g.groupby(["matrix", "format"])["gflops"].idxmax()]
        print("\n== best GPU variant per matrix and format ==")
        print(best[["matrix", "format", "kernel_name", "warps_per_block", "gflops", "GBs", "pct_of_peak"]].to_string(
            index=False, float_format=lambda v: f"{v:.4g}"))
    p = os.path.join(a.dir, "b200_dist.csv")
    if os.path.exists(p):
        df = pd.read_csv(p)
        g = df.groupby(["matrix", "gpus", "exchange"], as_index=False).agg(
            ms_per_step=("ms_per_step", "median"), gflops=("gflops", "median"), runs=("gflops", "size"))
        print("\n== b200_dist.csv (SPMV_B200_GPUS rows: iterated SpMV over several GPUs) ==")
        print(g.to_string(index=False, float_format=lambda v: f"{v:.4g}"))
    p = os.path.join(a.dir, "serial.csv")
    if os.path.exists(p):
        df = pd.read_csv(p)
        g = aggregate_serial(df)
        print("\n== serial.csv ==")
        print(g.to_string(index=False, float_format=lambda v: f"{v:.4g}"))
    p = os.path.join(a.dir, "omp.csv")
    if os.path.exists(p):
        df = pd.read_csv(p)
        g = aggregate_openmp(df)
        print("\n== omp.csv ==")
        print(g.to_string(index=False, float_format=lambda v: f"{v:.4g}"))


if __name__ == "__main__":
    main()
