"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): partition rule, exchange plan,
and the iterated x_{k+1} = A x_k loop of spmv_scpa_b200/dist.py with a CPU stand-in for
the GPU shard (the oracle's CSR loop -- test infrastructure only)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_balanced_row_cuts_rule(sp, O):
    from spmv_scpa_b200 import dist as D
    A = sp.gen_ragged(5000, 90)
    for parts in (1, 2, 3, 4, 8):
        cuts = D.balanced_row_cuts(A.IRP, parts)
        assert cuts[0] == 0 and cuts[-1] == A.M and len(cuts) == parts + 1
        assert (np.diff(cuts) >= 0).all()
        assert all(c % 32 == 0 for c in cuts[1:-1])
        # same greedy rule as the reference's partition (src/csr.c:218-276), cuts only rounded up
        ref = O.partition_rows(A.M, A.IRP, parts)
        nnz = np.diff(A.IRP.astype(np.int64)[cuts])
        if parts > 1:
            assert nnz.max() <= A.NZ / parts + 32 * 90 + 90  # one hack of slack
        assert len(ref) <= parts + 1
    # stencil: equal slabs
    S = sp.gen_stencil27(8, 8, 64)
    cuts = D.balanced_row_cuts(S.IRP, 4)
    assert np.abs(np.diff(cuts) - S.M / 4).max() <= 64 + 32


def test_exchange_plan_shapes():
    from spmv_scpa_b200 import dist as D
    plane = 100
    table = [(0, 400, 0, 500), (400, 800, 300, 900), (800, 1200, 700, 1200)]
    p1 = D.ExchangePlan(1, table)
    assert sorted(p1.recv) == [(0, 300, 400), (2, 800, 900)]
    assert sorted(p1.send) == [(0, 400, 500), (2, 700, 800)]
    assert (p1.boundary_lo, p1.boundary_hi) == (100, 300)
    assert p1.segments == [(0, 100, True), (300, 400, True), (100, 300, False)]
    assert p1.halo_bytes() == 8 * 2 * plane
    p0 = D.ExchangePlan(0, table)
    assert p0.send == [(1, 300, 400)] and p0.recv == [(1, 400, 500)]
    assert p0.segments == [(300, 400, True), (0, 300, False)]
    assert p1.max_push_targets() == 1 and p0.max_push_targets() == 1
    # all-gather case: everybody needs everything -> no interior
    full = [(0, 50, 0, 100), (50, 100, 0, 100)]
    q = D.ExchangePlan(0, full)
    assert q.segments == [(0, 50, True)] and q.cuts == []
    assert q.max_push_targets() == 1
    # ... and on 4 ranks one segment would have to feed 3 peers: the fused epilogue (2 targets)
    # cannot, DistSpMV then exchanges through NCCL
    full4 = [(25 * r, 25 * r + 25, 0, 100) for r in range(4)]
    assert D.ExchangePlan(1, full4).max_push_targets() == 3


class CpuShard:
    """Stand-in for CsrDevice on CPU tensors: rows [row0,row1) of a local CSR via the oracle."""

    def __init__(self, O, M, N, IRP, JA_local, AS):
        self.O, self.M, self.N = O, M, N
        self.IRP, self.JA, self.AS = IRP, JA_local, AS

    def spmv(self, x, y, kernel=0, warps_per_block=4, rows=None, push=None):
        r0, r1 = rows if rows is not None else (0, self.M)
        irp = self.IRP[r0:r1 + 1]
        k0, k1 = int(irp[0]), int(irp[-1])
        out = self.O.csr_spmv(r1 - r0, (irp - k0).astype(np.int32), self.JA[k0:k1], self.AS[k0:k1], x.numpy())
        y[r0:r1] = __import__("torch").from_numpy(out)


def _worker(rank, world, port, kind, steps, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import spmv_scpa_b200 as sp
    from spmv_scpa_b200 import dist as D
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if kind == "stencil":
        A = sp.gen_stencil27(6, 5, 4 * world + 1)
    else:
        A = sp.gen_uniform_random(700, 9, 3)  # needs all of x: all-gather exchange
    cuts = D.balanced_row_cuts(A.IRP, world, align=30 if kind == "stencil" else 32)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    k0, k1 = int(A.IRP[r0]), int(A.IRP[r1])
    JA, AS = A.JA[k0:k1], A.AS[k0:k1]
    c0, c1 = D.column_range(JA, r0, r1)
    table = D.gather_table(dist, r0, r1, c0, c1)
    plan = D.ExchangePlan(rank, table)
    shard = CpuShard(O, r1 - r0, c1 - c0, (A.IRP[r0:r1 + 1] - k0).astype(np.int32),
                     (JA.astype(np.int64) - c0).astype(np.int32), AS)
    x0 = np.random.default_rng(1).uniform(-1, 1, A.N)
    it = D.DistSpMV(dist, shard, plan, torch.from_numpy(x0[r0:r1].copy()), "cpu", mode="nccl")
    for _ in range(steps):
        it.step()
    mine = it.result_own().numpy().copy()
    # single-process answer
    x = x0.copy()
    for _ in range(steps):
        x = O.csr_spmv(A.M, A.IRP, A.JA, A.AS, x)
    err = float(np.abs(mine - x[r0:r1]).max() / max(1.0, np.abs(x).max()))
    q.put((rank, err, plan.segments, len(plan.recv)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "stencil"), (3, "stencil"), (2, "allgather")])
def test_iterated_spmv_gloo(world, kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, segs, nrecv in out:
        assert err < 1e-13, (rank, err)
        assert nrecv >= 1
        if kind == "stencil" and world == 3 and rank == 1:
            assert [s[2] for s in segs] == [True, True, False]  # two boundary slabs + interior


def test_row_cuts_equal_the_reference_partition_rule(O):
    """With no alignment, balanced_row_cuts IS the reference's partition_csr_rows
    (src/csr.c:218-276, restated and pinned in the oracle): same cut after every part, the
    unused parts of a short matrix collapse onto M."""
    from spmv_scpa_b200 import dist as D
    rng = np.random.default_rng(0)
    for _ in range(150):
        M = int(rng.integers(1, 400))
        lens = rng.integers(0, 30, M)
        if rng.random() < 0.3:
            lens[rng.integers(0, M)] = 5000       # a hub row
        IRP = np.zeros(M + 1, np.int32)
        IRP[1:] = np.cumsum(lens)
        for parts in (1, 2, 3, 5, 8):
            want = list(O.partition_rows(M, IRP, parts))
            want += [M] * (parts + 1 - len(want))
            assert list(D.balanced_row_cuts(IRP, parts, align=1)) == want
