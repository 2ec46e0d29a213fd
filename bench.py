#!/usr/bin/env python
"""bench.py -- FP64 SpMV (CSR, HLL) throughput on B200 against the HBM roofline.

Contract (one JSON line on stdout, printed by rank 0):
  python bench.py --gpus N --steps K --warmup W [--impl reference]
                  [--workload c1|c2|c3|c4] [--format csr|hll] [--kernel ID] [--wpb W]

  metric   GFLOP/s = 2*nnz / t  (BASELINE.json), with achieved HBM GB/s on the
           minimum-traffic byte count B_min = 12 nnz + 4 (M+1) + 8 M + 8 N.
  step     one pass y = A*x of the hot path over the resident matrix.
  workload N=1: BASELINE configs[1], 3D 27-point stencil 128^3 (n=2 097 152,
           nnz=55 742 968, B_min = 710 858 660 B > L2, so every step streams
           the matrix from HBM: "inputs larger than L2").
           N>1: weak scaling -- rank r owns a 128^3 slab (planes
           [128 r, 128 r+128)) of a 128 x 128 x 128N stencil, x_{k+1} = A x_k
           with a halo exchange of one plane per neighbour each step.
           --workload c5: BASELINE configs[4], the 512^3 stencil split into N
           z-slabs (strong scaling; 46 GB and 64-bit row offsets at N=1).
  value    whole-job GFLOP/s with inputs resident in HBM.
  e2e      same metric through the reference-facing C ABI entry point
           (csr_spmv_cuda_halfwarp_row / hll_spmv_cuda_warp_block: HOST x in,
           HOST y out), x H2D + kernel + y D2H inside the timed region.
  roofline dominant kernel: B_min / mean CUDA-event launch time vs the measured
           HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the reference's own CPU code (oracle/_ref,
           compiled from /root/reference/src) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# libgomp reads these when it is first loaded: idle OpenMP workers must sleep, not spin
# (the reference's OpenMP rows collapse otherwise on shared cores)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic(workload, kernel_symbol):
    """Per-launch dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the
    committed ncu summary (profiles/traffic.json, written by tools/summarize_profile.py), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            allt = json.load(f)
        keys = sorted(k for k in allt if k.endswith("_" + workload))
        for k in reversed(keys):
            for name, v in allt[k].items():
                if name.startswith(kernel_symbol):
                    return v
    except Exception:
        pass
    return None


WORKLOADS = {
    "c1": ("2D 5-point Poisson 1000x1000", lambda sp: sp.gen_poisson2d(1000, 1000)),
    "c2": ("3D 27-point stencil 128^3", lambda sp: sp.gen_stencil27(128, 128, 128)),
    "c3": ("uniform random n=16M, 32 nnz/row", lambda sp: sp.gen_uniform_random(16000000, 32, 42)),
    "c4": ("R-MAT scale 24, degree 16", lambda sp: sp.gen_rmat(24, 16)),
    "tiny": ("3D 27-point stencil 32^3 (self-test)", lambda sp: sp.gen_stencil27(32, 32, 32)),
}


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [l for (t, l) in self.lines if t0 - 0.1 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        for l in rows:
            f = [s.strip() for s in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------- reference (CPU) --
def cpu_reference_run(O, A_arrays, x, steps, warmup, threads=None):
    """Times the reference's own CPU CSR SpMV on this box.  Returns a dict with the best
    variant (GFLOP/s) and every variant measured.  oracle/_ref when present ("reference"),
    else the oracle port ("port")."""
    M, N, IRP, JA, AS = A_arrays
    nnz = len(JA)
    nthreads = threads or O.max_threads()
    variants = {}
    if O.ref_available():
        kind = "reference"
        R = O.RefCsr(M, N, IRP, JA, AS)

        def run(name):
            if name == "serial":
                return O.ref_csr_serial(R, x)[0]
            sched = "guided" if name == "omp_guided" else "nnz"
            return O.ref_csr_omp(R, x, nthreads, sched)[0]

        names = ["serial", "omp_guided", "omp_nnz"]
    else:
        kind = "port"
        IRPc, JAc, ASc = (np.ascontiguousarray(IRP, np.int32), np.ascontiguousarray(JA, np.int32),
                          np.ascontiguousarray(AS, np.float64))

        def run(name):
            return O.csr_spmv_timed(M, IRPc, JAc, ASc, x, 1 if name == "serial" else nthreads)[0]

        names = ["serial", "omp_guided"]
    for name in names:
        for _ in range(warmup):
            run(name)
        ms = [run(name) for _ in range(steps)]
        variants[name] = {"ms_per_step": statistics.median(ms),
                          "gflops": 2.0 * nnz / (statistics.median(ms) * 1e6),
                          "cores": 1 if name == "serial" else nthreads}
    best = max(variants, key=lambda k: variants[k]["gflops"])
    if kind == "reference" and nnz <= 600_000_000:
        # the reference's HLL CPU paths too (row-major hacks, its own packer); reported beside
        # the CSR ones, never the headline: the B200 arm's value is CSR
        try:
            ser, omp = [], []
            for _ in range(max(1, min(steps, 3))):
                s_ms, o_ms, _ = O.ref_hll_bench(R, x, nthreads)
                ser.append(s_ms)
                omp.append(o_ms)
            for name, ms, cores in (("hll_serial", statistics.median(ser), 1),
                                    ("hll_omp_guided", statistics.median(omp), nthreads)):
                variants[name] = {"ms_per_step": ms, "gflops": 2.0 * nnz / (ms * 1e6), "cores": cores}
        except Exception as e:  # a missing HLL row must not cost the CSR baseline
            variants["hll_error"] = repr(e)[:120]
    return kind, best, variants, nthreads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import spmv_scpa_b200 as sp  # host generators only; no GPU work on this arm
    from oracle import oracle as O
    # same config as the B200 arm at this N (see run_single / bench_dist): c2 at N=1, the
    # 128 x 128 x 128N stencil for N>1; for the 512^3 case a bounded sample (its first 64 planes,
    # 451 M entries: the whole matrix does not fit int32-indexed host CSR)
    sample = None
    if args.workload == "c5":
        desc = "3D 27-point stencil 512^3"
        A = sp.gen_stencil27_rows(512, 512, 512, 0, 64 * 512 * 512)
        sample = "rows of the first 64 of 512 planes (451 M of 3 610 M entries)"
    elif args.gpus > 1 and args.workload == "c2":
        desc = f"3D 27-point stencil 128x128x{128 * args.gpus} (one 128^3 slab per GPU in the B200 arm)"
        A = sp.gen_stencil27(128, 128, 128 * args.gpus)
    else:
        desc, make = WORKLOADS[args.workload]
        A = make(sp)
    x = np.random.default_rng(0).uniform(0, 1, A.N)
    arrays = (A.M, A.N, A.IRP, A.JA, A.AS)
    t0 = time.time()
    kind, best, variants, nthreads = cpu_reference_run(O, arrays, x, args.steps, max(args.warmup, 1))
    v = variants[best]
    bmin = sp.roofline_bytes(A.M, A.N, A.NZ)
    line = {
        "impl": "reference", "metric": "fp64_spmv_gflops", "value": v["gflops"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"{args.workload}: {desc}", "format": "csr", "rows": A.M, "nnz": A.NZ,
                   "variant": best, "B_min_bytes": bmin},
        "cpu_baseline": {"value": v["gflops"], "unit": "GFLOP/s", "cores": v["cores"], "kind": kind,
                         "sample": (sample or f"full matrix ({desc})") + f", {args.steps} SpMV passes per variant, median",
                         "variants": variants, "host_threads": nthreads, "build": O.ref_kind() if kind == "reference" else "port"},
        "e2e": {"value": v["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------- B200 (N = 1) --
def run_single(args):
    import torch
    import spmv_scpa_b200 as sp
    from oracle import oracle as O

    torch.cuda.set_device(0)
    desc, make = WORKLOADS[args.workload]
    A = make(sp)
    nnz, M, N = A.NZ, A.M, A.N
    bmin = sp.roofline_bytes(M, N, nnz)
    peak, peak_src = measured_peak()
    rng = np.random.default_rng(0)
    x_host = rng.uniform(0, 1, N)

    hcsr = sp.CsrDevice.from_host(A)
    hhll = hcsr.to_hll() if args.workload != "c4" else None
    x = torch.from_numpy(x_host).cuda()
    y = torch.zeros(M, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    ck = args.kernel if args.kernel is not None else 4
    hk = args.kernel if args.kernel is not None else 2
    KERNEL_SYMBOL = {("csr", 0): "csr_vec_kernel<1,", ("csr", 1): "csr_vec_kernel<32,",
                     ("csr", 2): "csr_vec_kernel", ("csr", 3): "csr_block_row_kernel",
                     ("csr", 4): "csr_stream_kernel", ("hll", 0): "hll_warp_kernel<1, 0",
                     ("hll", 1): "hll_warp_kernel<1, 0", ("hll", 2): "hll_warp_kernel<1, 1",
                     ("hll", 3): "hll_stream_kernel"}
    fmt = {"csr": (hcsr, ck, sp.CSR_KERNEL_NAMES[ck]),
           "hll": (hhll, hk, sp.HLL_KERNEL_NAMES[hk])}

    import ctypes
    small = bmin <= 256e6  # fits (or nearly fits) the 126 MB L2

    def timed_region(handle, kernel, steps, warmup):
        """K steps bracketed by sync on both sides; every launch also carries its own
        CUDA-event pair (on the launching stream) for the roofline."""
        for _ in range(warmup):
            handle.spmv(x, y, kernel=kernel, warps_per_block=args.wpb)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        c0 = sp.counters()["launches"]
        torch.cuda.synchronize()
        t0 = time.time()
        s_all = torch.cuda.Event(enable_timing=True)
        e_all = torch.cuda.Event(enable_timing=True)
        s_all.record(stream)
        for a, b in ev:
            if small:  # matrix fits L2: evict it before every timed launch (outside the events)
                sp._lib.b200.spmv_b200_flush_l2(ctypes.c_void_p(stream.cuda_stream))
            a.record(stream)
            handle.spmv(x, y, kernel=kernel, warps_per_block=args.wpb)
            b.record(stream)
        e_all.record(stream)
        torch.cuda.synchronize()
        t1 = time.time()
        per = [a.elapsed_time(b) for a, b in ev]
        # with flushes in between, the step time is the sum of the launch intervals
        total_ms = sum(per) if small else s_all.elapsed_time(e_all)
        return total_ms, per, sp.counters()["launches"] - c0, t0, t1

    # parity gate on this very input before anything is timed
    y_ref = O.csr_spmv(M, A.IRP, A.JA, A.AS, x_host)
    bound = O.csr_abs_bound(M, A.IRP, A.JA, A.AS, x_host)
    parity = {}
    for name, (h, k, _) in fmt.items():
        if h is None:
            continue
        y.fill_(float("nan"))
        h.spmv(x, y, kernel=k, warps_per_block=args.wpb)
        ok, worst = O.check_tolerance(y.cpu().numpy(), y_ref, bound, 1e-12)
        parity[name] = {"ok": ok, "worst_ratio": worst}
        if not ok:
            raise SystemExit(f"parity failed for {name} kernel {k}: worst ratio {worst}")

    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    results = {}
    t_first, t_last = None, None
    for name, (h, k, kname) in fmt.items():
        if h is None:
            continue
        total_ms, per, launches, t0, t1 = timed_region(h, k, args.steps, args.warmup)
        t_first = t0 if t_first is None else t_first
        t_last = t1
        ms_step = total_ms / args.steps
        kern_ms = statistics.mean(per)
        results[name] = {
            "kernel": kname, "kernel_id": k, "ms_per_step": ms_step,
            "gflops": 2.0 * nnz / (ms_step * 1e6), "gbs": bmin / (ms_step * 1e6),
            "launches": launches,
            "roofline": {"bound": "hbm", "achieved": bmin / (kern_ms * 1e6), "peak": peak,
                         "unit": "GB/s", "frac": bmin / (kern_ms * 1e6) / peak,
                         "traffic": ncu_traffic(args.workload, KERNEL_SYMBOL[(name, k)]),
                         "peak_source": peak_src, "kernel_ms_mean": kern_ms,
                         "kernel_ms_min": min(per), "algorithmic_bytes": bmin},
        }

    # e2e: the reference-facing C ABI call, host buffers (pinned), matrix resident after
    # the first call (device cache); x H2D + kernel + y D2H every step
    sp.set_timing(0, 0)  # one pass per call: no separately timed launches inside the e2e region
    import ctypes as C
    L = sp._lib.b200
    px = L.spmv_b200_host_alloc(N * 8)
    py = L.spmv_b200_host_alloc(M * 8)
    xh = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_double)), shape=(N,))
    yh = np.ctypeslib.as_array(C.cast(py, C.POINTER(C.c_double)), shape=(M,))
    xh[:] = x_host
    e2e = {}
    Hc = sp.csr_to_hll(A, True) if hhll is not None else None
    entry = {"csr": (L.csr_spmv_cuda_halfwarp_row_text if ck == 4 else L.csr_spmv_cuda_halfwarp_row, A),
             "hll": (L.hll_spmv_cuda_warp_block, Hc)}
    L.set_csr_warps_per_block(args.wpb)
    L.set_hll_warps_per_block(args.wpb)
    for name, (fn, mat) in entry.items():
        if mat is None:
            continue
        xp, yp = C.cast(px, C.POINTER(C.c_double)), C.cast(py, C.POINTER(C.c_double))
        for _ in range(max(args.warmup, 1)):
            if fn(mat.ptr, xp, yp, None) <= 0:
                raise SystemExit("e2e entry point failed: " + sp._lib.last_error())
        c0 = sp.counters()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(mat.ptr, xp, yp, None)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        c1 = sp.counters()
        ok, worst = O.check_tolerance(yh, y_ref, bound, 1e-12)
        if not ok:
            raise SystemExit(f"e2e parity failed for {name}: {worst}")
        e2e[name] = {"value": 2.0 * nnz / (dt * 1e9), "unit": "GFLOP/s", "ms_per_step": dt * 1e3,
                     "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) // args.steps,
                     "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) // args.steps}
    clocks = sampler.stop(t_first, time.time())
    sp.release_all()

    # CPU baseline on this box (bounded: the full C2 matrix is ~60 ms per serial pass)
    cpu = None
    if not args.no_cpu:
        kind, best, variants, nthreads = cpu_reference_run(
            O, (M, N, A.IRP, A.JA, A.AS), x_host, steps=5, warmup=1)
        cpu = {"value": variants[best]["gflops"], "unit": "GFLOP/s", "cores": variants[best]["cores"],
               "kind": kind, "variant": best,
               "sample": f"full {args.workload} matrix, 5 SpMV passes per variant (serial + OpenMP), median",
               "variants": variants, "host_threads": nthreads}

    head = results[args.format]
    line = {
        "metric": "fp64_spmv_gflops", "value": head["gflops"], "unit": "GFLOP/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "format": args.format, "kernel": head["kernel"],
                   "warps_per_block": args.wpb, "rows": M, "cols": N, "nnz": nnz, "B_min_bytes": bmin,
                   "l2_policy": "inputs larger than L2 (matrix streams 710 MB per step; no flush needed)"
                   if not small else "L2 flushed (512 MB memset) before every timed launch; ms_per_step = mean launch interval",
                   "e2e_matrix": "resident after first call (device cache keyed on host pointers + fingerprint)",
                   "e2e_path": "pinned host x/y; banded matrix -> x upload, row-chunk kernels and y download pipelined on 3 streams",
                   "parity": parity},
        "hbm_gbs": head["gbs"], "roofline": head["roofline"], "cpu_baseline": cpu,
        "e2e": e2e.get(args.format), "gpu_launches": head["launches"], "clocks": clocks,
        "formats": {k: {kk: vv for kk, vv in v.items()} for k, v in results.items()},
        "e2e_formats": e2e,
        "device": sp.device_info()["name"],
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"],
                    help="c5 = 3D 27-point stencil 512^3 (BASELINE configs[4]), z-slabs over --gpus ranks, strong scaling")
    ap.add_argument("--format", default="csr", choices=["csr", "hll"])
    ap.add_argument("--kernel", type=int, default=None)
    ap.add_argument("--wpb", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1 or args.workload == "c5":
        if "RANK" not in os.environ:  # plain `python bench.py --workload c5`: a 1-rank group
            os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                              MASTER_PORT=os.environ.get("MASTER_PORT", "29531"))
        import bench_dist
        return bench_dist.run(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
