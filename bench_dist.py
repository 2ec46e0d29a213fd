"""bench.py's multi-GPU leg (launched by torchrun, one rank per GPU).  Part of the
benchmark harness, not of the product package: it is allowed to use oracle/ as checker.

Workload: weak scaling of BASELINE configs[1] towards configs[4] -- rank r owns
the 128^3 slab (planes [128 r, 128 r + 128)) of a 128 x 128 x 128*N 27-point
stencil generated directly in HBM, and the job iterates x_{k+1} = A x_k with a
one-plane halo exchange per neighbour per step (see dist.py).  `--workload c5`
switches to the strong-scaling 512^3 case of configs[4] (planes split evenly).
"""
import json
import os
import statistics
import time

import numpy as np

import spmv_scpa_b200 as sp
from spmv_scpa_b200 import dist as D


def x0_slice(g0, g1):
    """x_0[g] in (0,1), a pure function of the global index (so every rank can
    regenerate any part of it)."""
    g = np.arange(g0, g1, dtype=np.uint64)
    z = g + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0


def slab_geometry(args, rank, world):
    if args.workload == "c5":
        nx = ny = nz = 512
    else:
        nx = ny = 128
        nz = 128 * world
    per = nz // world
    z0 = rank * per
    z1 = nz if rank == world - 1 else z0 + per
    return nx, ny, nz, z0, z1


def run(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)

    nx, ny, nz, z0, z1 = slab_geometry(args, rank, world)
    plane = nx * ny
    r0, r1 = z0 * plane, z1 * plane
    c0, c1 = max(0, z0 - 1) * plane, min(nz, z1 + 1) * plane
    table = D.gather_table(dist, r0, r1, c0, c1, device=device)
    plan = D.ExchangePlan(rank, table)

    t_build = time.time()
    shard = sp.CsrDevice.stencil27(nx, ny, nz, z0, z1, col_offset=c0, n_local=c1 - c0, cuts=plan.cuts)
    torch.cuda.synchronize()
    t_build = time.time() - t_build
    kernel = args.kernel if args.kernel is not None else 4
    # Exchange mode.  "push" (default): the boundary-row kernel stores the halo into the
    # neighbour's HBM from its epilogue, signal/wait kernels order the steps -- no collective on
    # the step path.  "nccl": batched isend/irecv on a high-priority side stream overlapping
    # the interior rows.  Measured ms per step at N = 2 / 4 / 8 (N=1: 0.1142):
    #   push 0.1186 / 0.1232 / 0.1238     nccl 0.1181 / 0.1214 / 0.1251     (profiles/r1_bench_N*)
    mode = os.environ.get("SPMV_B200_EXCHANGE", "push")

    x0 = torch.from_numpy(x0_slice(r0, r1)).to(device)
    it = D.DistSpMV(dist, shard, plan, x0, device, mode=mode, kernel=kernel, wpb=args.wpb)

    # ---- parity of step 1 on this rank's rows against the oracle (bounded: own slab) ----
    from oracle import oracle as O
    it.step()
    torch.cuda.synchronize()
    y1 = it.result_own().cpu().numpy()
    plane = nx * ny
    # checked rows: the whole slab when it is small, else its first and last two planes
    # (the rows that depend on the halo) -- keeps the host-side oracle bounded for 512^3
    M_own = r1 - r0
    spans = [(0, M_own)] if M_own <= 4 * 128 * 128 * 128 else [(0, 2 * plane), (M_own - 2 * plane, M_own)]
    xl = x0_slice(c0, c1)  # local view of x0 over [c0, c1), columns shifted like the shard's
    ok, worst = True, 0.0
    for a, b in spans:
        A_part = sp.gen_stencil27_rows(nx, ny, nz, r0 + a, r0 + b)
        ja_local = (A_part.JA.astype(np.int64) - c0).astype(np.int32)
        y_ref = O.csr_spmv(A_part.M, A_part.IRP, ja_local, A_part.AS, xl)
        bound = O.csr_abs_bound(A_part.M, A_part.IRP, ja_local, A_part.AS, xl)
        ok_i, worst_i = O.check_tolerance(y1[a:b], y_ref, bound, 1e-12)
        ok, worst = ok and ok_i, max(worst, worst_i)
        del A_part
    flag = torch.tensor([0 if ok else 1], device=device)
    dist.all_reduce(flag)
    if int(flag.item()) != 0:
        raise SystemExit(f"rank {rank}: multi-GPU parity failed after step 1 (worst ratio {worst})")

    # ---- timed region ----
    nnz_local = shard.NZ
    nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device=device)
    dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())
    n_total = nx * ny * nz
    bmin_total = sp.roofline_bytes(n_total, n_total, nnz_total)

    def reset():
        it.X[0].zero_()
        it.X[1].zero_()
        it.own(0).copy_(x0)
        it.step_no = 0
        it._initial_exchange()

    reset()
    use_graph = os.environ.get("SPMV_B200_GRAPH", "1") == "1"
    for _ in range(2):
        it.step()
    graph_err = None
    if use_graph:
        try:
            it.build_graph(2)
        except Exception as e:  # stay measurable if capture is refused
            graph_err = repr(e)[:200]
            it.graph = None
    reset()
    sampler = None
    if rank == 0:
        from bench import ClockSampler  # bench.py is on sys.path
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.25)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t_region0 = time.time()
    c_before = sp.counters()["launches"]
    # The W warm-up steps run on the same stream immediately before the K timed steps, with
    # no host synchronisation in between: the ranks leave the host barrier up to a few ms
    # apart, and the first steps absorb that skew on the device (each rank's wait kernel
    # paces it to its neighbours) instead of charging it to the timed region.
    # push mode lets a rank drift one step ahead per hop, so the skew needs ~world steps to drain
    warm = max(args.warmup, world + 4)
    warm += warm % 2
    it.run(warm)
    # device-side rendezvous right before the start event: the first graph launch costs a
    # different number of milliseconds on every rank, and in push mode a rank may run one step
    # ahead per hop, so without it the early ranks' timed region would include the late ranks'
    # start-up.  (Stream-ordered: the compute stream waits for the all-reduce, not the host.)
    sync_word = torch.zeros(1, device=device)
    dist.all_reduce(sync_word)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    it.run(args.steps)
    end.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop(t_region0, time.time()) if sampler else None
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = sp.counters()["launches"] - c_before
    if it.graph is not None:
        # kernels replayed from the graph are not seen by the library's launch counter
        launches = args.steps * (len(plan.segments) + (2 if mode == "push" else 0))
    it.check_errors()
    finite = bool(torch.isfinite(it.result_own()).all().item())

    diag = None
    if os.environ.get("SPMV_B200_DIAG") == "1":
        # where does a step's time go?  each piece alone, 20 back-to-back launches
        def timed(fn, n=20):
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n * 1e3
        segs = plan.segments
        xb, yb = it.X[0], it.own(1)
        diag = {}
        for (s0, s1, isb) in segs:
            diag[f"rows[{s0},{s1}){'B' if isb else 'I'}_us"] = timed(
                lambda: shard.spmv(xb, yb, kernel=kernel, warps_per_block=args.wpb, rows=(s0, s1)))
        diag["all_rows_us"] = timed(lambda: shard.spmv(xb, yb, kernel=kernel, warps_per_block=args.wpb))
        diag["eager_step_us"] = timed(it.step)
        if it.graph is not None:
            diag["graph_2steps_us"] = timed(lambda: it.graph.replay())
            # per-replay times on every rank (does one rank pace the other?)
            torch.cuda.synchronize()
            dist.barrier()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(26)]
            evs[0].record()
            for i in range(25):
                it.graph.replay()
                evs[i + 1].record()
            torch.cuda.synchronize()
            mine = torch.tensor([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(25)],
                                dtype=torch.float64, device=device)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            diag["per_replay_us_by_rank"] = [[round(v, 1) for v in t.cpu().tolist()[:6]] for t in allr]

            def variant(name, body, n=25):
                torch.cuda.synchronize()
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                a.record()
                for _ in range(n):
                    body()
                b.record()
                cpu_us = (time.perf_counter() - t0) / n * 1e6
                torch.cuda.synchronize()
                diag[name] = {"gpu_us": a.elapsed_time(b) / n * 1e3, "cpu_issue_us": cpu_us}
            variant("v1_back_to_back", lambda: it.graph.replay())
            scratch = torch.cuda.Event()

            def with_event():
                it.graph.replay()
                scratch.record()
            variant("v2_event_after_each", with_event)
            variant("v4_eager_step", it.step, n=50)
        it.step_no = 0

    # ---- kernel-only roofline on this rank: all rows, no exchange ----
    xs = it.X[0]
    ys = torch.zeros(shard.M, dtype=torch.float64, device=device)
    per = shard.time(xs, ys, kernel=kernel, warps_per_block=args.wpb, warmup=3, reps=20)
    bmin_local = sp.roofline_bytes(shard.M, shard.N, shard.NZ)
    kern_ms = statistics.mean(per)

    # ---- e2e: host x slice in, host y slice out, every step ----
    xh = torch.from_numpy(x0_slice(r0, r1)).pin_memory()
    yh = torch.empty(shard.M, dtype=torch.float64).pin_memory()

    def e2e_step():
        it.own(0).copy_(xh, non_blocking=True)
        it._exchange_nccl(0)
        shard.spmv(it.X[0], it.own(1), kernel=kernel, warps_per_block=args.wpb)
        yh.copy_(it.own(1), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dist.barrier()
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=device)
    dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)

    if rank == 0:
        from bench import measured_peak  # bench.py is on sys.path
        peak, peak_src = measured_peak()
        ms_step = total_ms / args.steps
        line = {
            "metric": "fp64_spmv_gflops", "value": 2.0 * nnz_total / (ms_step * 1e6), "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if args.workload == "c5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"c5: 3D 27-point stencil 512^3, {world} z-slabs" if args.workload == "c5"
                                    else f"c2-slabs: 3D 27-point stencil 128x128x{nz}, one 128^3 slab per GPU"),
                       "format": "csr", "kernel": sp.CSR_KERNEL_NAMES[kernel], "warps_per_block": args.wpb,
                       "rows": n_total, "nnz": nnz_total, "B_min_bytes": bmin_total,
                       "iteration": "x_{k+1} = A x_k, halo exchange of one plane per neighbour per step",
                       "exchange": mode, "cuda_graph": it.graph is not None, "graph_error": graph_err,
                       "halo_bytes_per_rank_per_step": plan.halo_bytes(),
                       "l2_policy": "inputs larger than L2 (>= 0.7 GB streamed per GPU per step)",
                       "shard_build_s": t_build, "finite": finite, "diag": diag},
            "hbm_gbs": bmin_total / (ms_step * 1e6),
            "roofline": {"bound": "hbm", "achieved": bmin_local / (kern_ms * 1e6), "peak": peak, "unit": "GB/s",
                         "frac": bmin_local / (kern_ms * 1e6) / peak, "traffic": None, "peak_source": peak_src,
                         "kernel_ms_mean": kern_ms, "algorithmic_bytes": bmin_local,
                         "note": "rank 0 shard, all rows in one call, no exchange"},
            "cpu_baseline": None,
            "e2e": {"value": 2.0 * nnz_total / (float(e2e_dt.item()) * 1e9), "unit": "GFLOP/s",
                    "h2d_bytes_per_step": 8 * shard.M * world, "d2h_bytes_per_step": 8 * shard.M * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": {"step1_vs_oracle": True, "worst_ratio_rank0": worst},
        }
        print(json.dumps(line))
    it.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    return 0
