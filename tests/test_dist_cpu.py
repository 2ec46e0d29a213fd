"""Host logic of the multi-GPU iterated SpMV on CPU: the C planner of csrc/dist.cu (partition rule,
shard descriptors, exchange plans) driven through ctypes, checked
  * against the oracle's restatement of the reference's partition rule,
  * across real processes (gloo, world_size 2 and 3) with a CPU stand-in for the GPU shard that
    follows the plan's send / recv lists and row segments,
  * and under an adversarial single-process schedule of the PUSH protocol (a neighbour running one
    step ahead, interior rows executed as late as the stream order allows) on a structurally
    NON-symmetric banded matrix -- the write-after-read hazard the round-1 advisor found.
The CPU stand-ins use the oracle's CSR loop: test infrastructure only."""
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
    # 64-bit row offsets give the same cuts
    assert list(D.balanced_row_cuts(S.IRP.astype(np.int64), 4)) == list(cuts)


def test_row_cuts_equal_the_reference_partition_rule(O):
    """With no alignment, spmv_b200_partition_rows IS the reference's partition_csr_rows
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


def test_exchange_plan_shapes():
    from spmv_scpa_b200 import dist as D
    plane = 100
    table = [(0, 400, 0, 500), (400, 800, 300, 900), (800, 1200, 700, 1200)]
    p1 = D.Plan(1, table)
    assert sorted(p1.recv) == [(0, 300, 400), (2, 800, 900)]
    assert sorted(p1.send) == [(0, 400, 500), (2, 700, 800)]
    assert (p1.boundary_lo, p1.boundary_hi) == (100, 300)
    assert p1.segments == [(0, 100, True), (300, 400, True), (100, 300, False)]
    assert p1.cuts == [100, 300]
    assert p1.halo_bytes() == 8 * 2 * plane
    assert p1.mode == "push"
    p0 = D.Plan(0, table)
    assert p0.send == [(1, 300, 400)] and p0.recv == [(1, 400, 500)]
    assert p0.segments == [(300, 400, True), (0, 300, False)]
    # all-gather case: everybody needs everything -> no interior
    full = [(0, 50, 0, 100), (50, 100, 0, 100)]
    q = D.Plan(0, full)
    assert q.segments == [(0, 50, True)] and q.cuts == []
    assert q.mode == "push" and q.all_gather    # one peer: the push epilogue can feed it
    # ... on 4 ranks one segment feeds 3 peers: still the fused epilogue (up to 8 targets: the
    # all-gather of a general matrix is peer stores from the SpMV kernel), unless
    # SPMV_B200_PUSH_ALL=0 asks for the NCCL exchange; every rank agrees
    full4 = [(25 * r, 25 * r + 25, 0, 100) for r in range(4)]
    for r in range(4):
        p = D.Plan(r, full4)
        assert p.mode == "push" and p.all_gather and len(p.send) == 3
    os.environ["SPMV_B200_PUSH_ALL"] = "0"
    try:
        for r in range(4):
            assert D.Plan(r, full4).mode == "nccl"
        assert D.Plan(1, full4, mode="push").mode == "push"      # asked for explicitly
        assert D.Plan(0, full).mode == "push"                    # halo-sized plans are not affected
    finally:
        del os.environ["SPMV_B200_PUSH_ALL"]
    # ten ranks: nine targets per segment are more than one launch can feed
    full10 = [(10 * r, 10 * r + 10, 0, 100) for r in range(10)]
    assert D.Plan(3, full10).mode == "nccl"
    with pytest.raises(RuntimeError):
        D.Plan(3, full10, mode="push")
    # unequal slices: all ranks need everything but ncclAllGather cannot be used
    ragged = [(0, 30, 0, 100), (30, 100, 0, 100)]
    assert not D.Plan(0, ragged).all_gather
    # row ranges that do not tile a rank's column range are refused
    with pytest.raises(RuntimeError):
        D.Plan(0, [(0, 50, 0, 120), (50, 100, 0, 100)])


def test_stencil_descriptor_equals_scan(sp):
    """The analytic descriptor of a stencil slab equals a scan of the slab's actual columns."""
    from spmv_scpa_b200 import dist as D
    nx, ny, nz = 6, 5, 9
    A = sp.gen_stencil27(nx, ny, nz)
    plane = nx * ny
    for z0, z1 in ((0, 3), (3, 7), (7, 9), (0, 9), (4, 5)):
        r0, r1 = z0 * plane, z1 * plane
        k0, k1 = int(A.IRP[r0]), int(A.IRP[r1])
        got = D.desc_tuple(D.shard_scan(r0, r1, A.IRP[r0:r1 + 1], A.JA[k0:k1]))
        want = D.desc_tuple(D.stencil27_shard_desc(nx, ny, nz, z0, z1))
        assert got == want, (z0, z1, got, want)


def upper_banded(n, width, rng):
    """Row r holds columns r .. r+width-1 (clipped): structurally NON-symmetric."""
    lens = np.minimum(width, n - np.arange(n))
    IRP = np.zeros(n + 1, np.int32)
    IRP[1:] = np.cumsum(lens)
    JA = np.concatenate([np.arange(r, r + lens[r]) for r in range(n)]).astype(np.int32)
    AS = rng.uniform(-1, 1, int(IRP[-1])) / width
    return IRP, JA, AS


def test_boundary_covers_rows_that_read_the_halo():
    """Upper-banded matrix: no peer below needs this rank's LAST rows, but they read the next
    rank's slice.  They must be boundary rows (run before the signal), or a neighbour that is one
    step ahead overwrites the halo under them."""
    from spmv_scpa_b200 import dist as D
    rng = np.random.default_rng(3)
    n, width, world = 640, 40, 4
    IRP, JA, AS = upper_banded(n, width, rng)
    cuts = D.balanced_row_cuts(IRP, world)
    table = [D.shard_scan(int(cuts[r]), int(cuts[r + 1]), IRP[cuts[r]:cuts[r + 1] + 1],
                          JA[IRP[cuts[r]]:IRP[cuts[r + 1]]]) for r in range(world)]
    for r in range(world - 1):
        p = D.Plan(r, table)
        M = int(cuts[r + 1] - cuts[r])
        assert table[r].read_hi == M - (width - 1)          # last width-1 rows read rank r+1
        assert p.boundary_hi <= table[r].read_hi            # ... and are boundary rows
        assert p.send == ([] if r == 0 else [(r - 1, int(cuts[r]), int(cuts[r]) + width - 1)])
        if r > 0:
            assert p.boundary_lo >= width - 1               # the rows rank r-1 needs


class CpuRank:
    """One rank of the PUSH protocol on numpy arrays, following the C plan."""

    def __init__(self, O, plan, IRP, JA, AS, x0):
        self.O, self.P = O, plan
        r0, r1, c0, c1 = plan.r0, plan.r1, plan.c0, plan.c1
        k0, k1 = int(IRP[r0]), int(IRP[r1])
        self.irp = (IRP[r0:r1 + 1] - k0).astype(np.int32)
        self.ja = (JA[k0:k1].astype(np.int64) - c0).astype(np.int32)
        self.as_ = AS[k0:k1]
        self.X = [np.zeros(c1 - c0), np.zeros(c1 - c0)]
        self.X[0][r0 - c0:r1 - c0] = x0[r0:r1]
        self.epoch, self.flags, self.step_no = 0, {}, 0

    def rows(self, a, b, src, dst):
        irp = self.irp[a:b + 1]
        k0, k1 = int(irp[0]), int(irp[-1])
        y = self.O.csr_spmv(b - a, (irp - k0).astype(np.int32), self.ja[k0:k1], self.as_[k0:k1], self.X[src])
        own0 = self.P.r0 - self.P.c0
        self.X[dst][own0 + a:own0 + b] = y
        return y


def run_push_protocol(O, plans, IRP, JA, AS, x0, steps, rng):
    """Adversarial interleaving: wait / boundary+push / signal ops run as early as their wait
    condition allows, interior ops only when nothing else can run (i.e. as late as possible)."""
    ranks = [CpuRank(O, p, IRP, JA, AS, x0) for p in plans]
    nb = [sorted({q for q, _, _ in p.send} | {q for q, _, _ in p.recv}) for p in plans]
    # initial halo + epoch signal (spmv_b200_dist_set_x)
    for r, R in enumerate(ranks):
        for q, g0, g1 in R.P.send:
            Q = ranks[q]
            Q.X[0][g0 - Q.P.c0:g1 - Q.P.c0] = R.X[0][g0 - R.P.c0:g1 - R.P.c0]
        R.epoch += 1
        for q in nb[r]:
            ranks[q].flags[r] = R.epoch
    pending_interior = [None] * len(ranks)      # (src, dst) of a step whose interior has not run

    def front_ready(r):
        R = ranks[r]
        return pending_interior[r] is None and R.step_no < steps and \
            all(R.flags.get(q, 0) >= R.epoch for q in nb[r])

    while any(R.step_no < steps or pending_interior[r] for r, R in enumerate(ranks)):
        ready = [r for r in range(len(ranks)) if front_ready(r)]
        if ready:
            r = int(rng.choice(ready))
            R = ranks[r]
            src, dst = R.step_no % 2, 1 - R.step_no % 2
            for a, b, is_b in R.P.segments:
                if not is_b:
                    continue
                y = R.rows(a, b, src, dst)
                for q, g0, g1 in R.P.send:                    # push epilogue
                    lo, hi = max(g0 - R.P.r0, a), min(g1 - R.P.r0, b)
                    if lo < hi:
                        Q = ranks[q]
                        Q.X[dst][R.P.r0 + lo - Q.P.c0:R.P.r0 + hi - Q.P.c0] = y[lo - a:hi - a]
            R.epoch += 1                                      # signal
            for q in nb[r]:
                ranks[q].flags[r] = R.epoch
            pending_interior[r] = (src, dst)
            R.step_no += 1
            continue
        late = [r for r in range(len(ranks)) if pending_interior[r]]
        assert late, "protocol deadlock"
        r = int(rng.choice(late))
        src, dst = pending_interior[r]
        for a, b, is_b in ranks[r].P.segments:
            if not is_b:
                ranks[r].rows(a, b, src, dst)
        pending_interior[r] = None
    return [R.X[steps % 2][R.P.r0 - R.P.c0:R.P.r1 - R.P.c0] for R in ranks]


@pytest.mark.parametrize("kind", ["upper_banded", "stencil"])
def test_push_protocol_is_safe_under_adversarial_schedules(sp, O, kind):
    from spmv_scpa_b200 import dist as D
    rng = np.random.default_rng(11)
    if kind == "upper_banded":
        n, world = 768, 4
        IRP, JA, AS = upper_banded(n, 48, rng)
    else:
        A = sp.gen_stencil27(5, 4, 12)
        n, world = A.M, 3
        IRP, JA, AS = A.IRP.copy(), A.JA.copy(), A.AS / 30.0
    cuts = D.balanced_row_cuts(IRP, world, align=20 if kind == "stencil" else 32)
    table = [D.shard_scan(int(cuts[r]), int(cuts[r + 1]), IRP[cuts[r]:cuts[r + 1] + 1],
                          JA[IRP[cuts[r]]:IRP[cuts[r + 1]]]) for r in range(world)]
    plans = [D.Plan(r, table, mode="push") for r in range(world)]
    x0 = rng.uniform(-1, 1, n)
    steps = 5
    want = x0.copy()
    for _ in range(steps):
        want = O.csr_spmv(n, IRP, JA, AS, want)
    for seed in range(20):
        got = np.concatenate(run_push_protocol(O, plans, IRP, JA, AS, x0, steps, np.random.default_rng(seed)))
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (kind, seed)


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
    desc = D.shard_scan(r0, r1, A.IRP[r0:r1 + 1], A.JA[k0:k1])
    table = D.gather_table(dist, desc)
    plan = D.Plan(rank, table)
    x0 = np.random.default_rng(1).uniform(-1, 1, A.N)
    R = CpuRank(O, plan, A.IRP, A.JA, A.AS, x0)

    def exchange(buf):
        ops = []
        X = torch.from_numpy(R.X[buf])
        for peer, g0, g1 in plan.send:
            ops.append(dist.P2POp(dist.isend, X[g0 - plan.c0:g1 - plan.c0], peer))
        for peer, g0, g1 in plan.recv:
            ops.append(dist.P2POp(dist.irecv, X[g0 - plan.c0:g1 - plan.c0], peer))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()

    exchange(0)
    for k in range(steps):
        src, dst = k % 2, 1 - k % 2
        for a, b, _ in plan.segments:
            R.rows(a, b, src, dst)
        exchange(dst)
    mine = R.X[steps % 2][r0 - plan.c0:r1 - plan.c0].copy()
    x = x0.copy()
    for _ in range(steps):
        x = O.csr_spmv(A.M, A.IRP, A.JA, A.AS, x)
    err = float(np.abs(mine - x[r0:r1]).max() / max(1.0, np.abs(x).max()))
    q.put((rank, err, plan.segments, len(plan.recv), plan.mode))
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
    for rank, err, segs, nrecv, mode in out:
        assert err < 1e-13, (rank, err)
        assert nrecv >= 1
        if kind == "stencil" and world == 3 and rank == 1:
            assert [s[2] for s in segs] == [True, True, False]  # two boundary slabs + interior
