"""Multi-GPU iterated SpMV (needs >= 2 GPUs; skipped on a 1-GPU box), three ways:
  * one process per GPU (the deployment bench.py uses): ranks wired through CUDA IPC by
    spmv_b200_dist_connect, blobs moved with torch.distributed -- stencil (push and NCCL
    exchange), a structurally NON-symmetric banded matrix with one rank deliberately slowed
    down (the write-after-read hazard of round 1), and a general matrix (all-gather plan);
  * one process driving both GPUs (spmv_b200_dist_group_*) from Python;
  * bin/dist_check: the same through the C ABI only, no Python in the loop.
Every result is compared with the single-process oracle after several steps."""
import os
import socket
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def x0_slice(g0, g1):
    g = np.arange(g0, g1, dtype=np.uint64)
    z = g + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0


def make_case(sp, kind):
    """(M, IRP, JA, AS) of the global matrix, scaled so that |x_k| stays O(1)."""
    if kind == "stencil":
        A = sp.gen_stencil27(24, 20, 16)
        return A.M, A.IRP.copy(), A.JA.copy(), A.AS / 52.0
    if kind == "upper_banded":
        n, w = 40000, 300
        lens = np.minimum(w, n - np.arange(n))
        IRP = np.zeros(n + 1, np.int32)
        IRP[1:] = np.cumsum(lens)
        JA = (np.repeat(np.arange(n), lens) + (np.arange(IRP[-1]) - np.repeat(IRP[:-1], lens))).astype(np.int32)
        AS = (0.5 + x0_slice(0, int(IRP[-1]))) / (2.0 * w)
        return n, IRP, JA, AS
    A = sp.gen_uniform_random(30000, 12, 5)
    return A.M, A.IRP.copy(), A.JA.copy(), A.AS / 12.0


def _worker(rank, world, port, kind, mode, steps, slow_rank, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import spmv_scpa_b200 as sp
    from spmv_scpa_b200 import dist as D
    from oracle import oracle as O
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    M, IRP, JA, AS = make_case(sp, kind)
    cuts = D.balanced_row_cuts(IRP, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    k0, k1 = int(IRP[r0]), int(IRP[r1])
    desc = D.shard_scan(r0, r1, IRP[r0:r1 + 1], JA[k0:k1])
    table = D.gather_table(dist, desc, device=dev)
    plan = D.Plan(rank, table, mode=mode)
    shard = sp.CsrDevice.from_arrays(r1 - r0, plan.c1 - plan.c0, IRP[r0:r1 + 1] - k0, JA[k0:k1], AS[k0:k1],
                                     col_offset=plan.c0, cuts=plan.cuts)
    it = D.DistSpMV(dist, shard, plan, kernel=2 if kind == "uniform" else 4, wpb=4, device=dev)
    x0 = x0_slice(0, M)
    it.set_x(x0[r0:r1])
    if rank == slow_rank:
        # hold this rank's stream back: its neighbours get one step ahead and push into its halo
        with torch.cuda.stream(it.stream):
            torch.cuda._sleep(int(4e8))
    it.iterate(steps)
    mine = it.result_own()
    graph = it.has_graph
    # the same again after a reset (set_x on a live job), this time through the captured graph
    it.set_x(x0[r0:r1])
    it.iterate(steps)
    again = it.result_own()
    graph2 = it.has_graph
    it.close()
    x, bound = x0.copy(), None
    for _ in range(steps):
        bound = O.csr_abs_bound(M, IRP, JA, AS, np.abs(x))
        x = O.csr_spmv(M, IRP, JA, AS, x)
    ok, worst = O.check_tolerance(mine, x[r0:r1], bound[r0:r1] * steps + 1e-300, 1e-12)
    ok2, worst2 = O.check_tolerance(again, x[r0:r1], bound[r0:r1] * steps + 1e-300, 1e-12)
    q.put((rank, ok and ok2, max(worst, worst2), it.mode, graph or graph2, plan.segments))
    dist.barrier()
    dist.destroy_process_group()


CASES = [("stencil", "push", -1), ("stencil", "nccl", -1), ("upper_banded", "push", 0),
         ("upper_banded", "push", 1), ("uniform", "auto", -1), ("uniform", "nccl", -1)]


@pytest.mark.parametrize("kind,mode,slow_rank", CASES)
def test_two_gpu_iterated_spmv(kind, mode, slow_rank):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, mode, 6, slow_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=150) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, worst, used_mode, graph, segs in out:
        assert ok, (kind, mode, rank, worst)
        if mode != "auto":
            assert used_mode == mode
        if used_mode == "push":
            assert graph, "the push step should have been captured in a CUDA graph"
        if kind == "upper_banded" and rank == 0:
            assert segs[0][2] and segs[0][1] == segs[-1][1] + (segs[0][1] - segs[0][0])  # tail rows are boundary


@pytest.mark.parametrize("kind", ["stencil", "upper_banded", "uniform"])
def test_single_process_group(sp, O, kind):
    """spmv_b200_dist_group_*: one process, both GPUs, host x in / host x out."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from spmv_scpa_b200 import dist as D
    M, IRP, JA, AS = make_case(sp, kind)
    A = sp.csr_from_arrays(kind, M, M, IRP, JA, AS)
    g = D.DistGroup.from_csr(A, 2, kernel=2 if kind == "uniform" else 4)
    x0 = x0_slice(0, M)
    steps = 5
    x, bound = x0.copy(), None
    for _ in range(steps):
        bound = O.csr_abs_bound(M, IRP, JA, AS, np.abs(x))
        x = O.csr_spmv(M, IRP, JA, AS, x)
    for _ in range(2):       # second round: after a reset, through the captured graph
        g.set_x(x0)
        ms = g.iterate(steps)
        assert ms > 0
        ok, worst = O.check_tolerance(g.get_x(), x, bound * steps + 1e-300, 1e-12)
        assert ok, (kind, worst)
    g.close()
    # generated-in-HBM stencil shards
    if kind == "stencil":
        nx, ny, nz = 20, 18, 14
        g = D.DistGroup.stencil27(nx, ny, nz, 2)
        S = sp.gen_stencil27(nx, ny, nz)
        x0 = x0_slice(0, S.M)
        g.set_x(x0)
        g.iterate(3)
        x = x0.copy()
        for _ in range(3):
            bound = O.csr_abs_bound(S.M, S.IRP, S.JA, S.AS, np.abs(x))
            x = O.csr_spmv(S.M, S.IRP, S.JA, S.AS, x)
        ok, worst = O.check_tolerance(g.get_x(), x, bound * 3, 1e-12)
        assert ok, worst
        assert g.mode() == "push"
        g.close()
    sp.release_all()


@pytest.mark.parametrize("spec,mode,knobs", [
    ("stencil:48:40:64", "auto", ()), ("upper:60000:200", "push", ()), ("uniform:200000:16", "auto", ()),
    ("stencil:32:32:40", "nccl", ()), ("uniform:200000:16", "nccl", ()),
    # general matrix through column panels: the last panel's epilogue adds, stores and pushes the
    # finished sums to every peer (the all-gather fused into the SpMV: EPI_ACC_PUSH)
    ("uniform:200000:16", "push", ("sell_panels=2",)), ("uniform:300000:24", "auto", ("sell_panels=3",)),
    # power-law rows: every peer needs every slice, the binned kernels push
    ("rmat:16:8", "auto", ()), ("rmat:16:8", "nccl", ())])
def test_dist_check_binary(spec, mode, knobs):
    """The C program: N GPUs vs one GPU through the C ABI only."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "bin", "dist_check")
    extra = [a for k in knobs for a in ("--knob", k)]
    r = subprocess.run([exe, "--gpus", str(n), "--steps", "6", "--matrix", spec, "--mode", mode] + extra,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
