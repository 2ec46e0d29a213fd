"""2-GPU iterated SpMV (needs >= 2 GPUs; skipped on a 1-GPU box): both exchange modes against the
single-process oracle after several steps, including the fused peer-store epilogue."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, graph, steps, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import spmv_scpa_b200 as sp
    from spmv_scpa_b200 import dist as D
    from bench_dist import x0_slice
    from oracle import oracle as O
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nx, ny, nz = 24, 20, 8 * world
    plane = nx * ny
    z0, z1 = rank * 8, rank * 8 + 8
    r0, r1 = z0 * plane, z1 * plane
    c0, c1 = max(0, z0 - 1) * plane, min(nz, z1 + 1) * plane
    plan = D.ExchangePlan(rank, D.gather_table(dist, r0, r1, c0, c1, device=dev))
    shard = sp.CsrDevice.stencil27(nx, ny, nz, z0, z1, col_offset=c0, n_local=c1 - c0, cuts=plan.cuts)
    x0 = torch.from_numpy(x0_slice(r0, r1)).to(dev)
    it = D.DistSpMV(dist, shard, plan, x0, dev, mode=mode, kernel=4, wpb=4)
    if graph:
        it.step(); it.step()            # eager warm-up, then start again from x0 through the graph
        it.build_graph(2)
        it.X[0].zero_(); it.X[1].zero_(); it.own(0).copy_(x0); it.step_no = 0
        it._initial_exchange()
        it.run(steps)
    else:
        for _ in range(steps):
            it.step()
    torch.cuda.synchronize()
    it.check_errors()
    mine = it.result_own().cpu().numpy()
    it.close()
    A = sp.gen_stencil27(nx, ny, nz)
    x = x0_slice(0, A.N)
    bound = None
    for _ in range(steps):
        bound = O.csr_abs_bound(A.M, A.IRP, A.JA, A.AS, np.abs(x))
        x = O.csr_spmv(A.M, A.IRP, A.JA, A.AS, x)
    ok, worst = O.check_tolerance(mine, x[r0:r1], bound[r0:r1] * steps, 1e-12)
    q.put((rank, ok, worst))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,graph", [("nccl", False), ("push", False), ("push", True), ("nccl", True)])
def test_two_gpu_iterated_spmv(mode, graph):
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
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, graph, 5 if graph else 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, worst in out:
        assert ok, (mode, rank, worst)
