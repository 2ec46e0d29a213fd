"""Row-partitioned, iterated SpMV over several GPUs: x_{k+1} = A x_k.

NEW SURFACE -- the reference is single-GPU.  What it does offer
is the partition rule: contiguous row ranges with balanced nnz, a range closed
as soon as its running nnz reaches total/parts (partition_csr_rows, reference
src/csr.c:218-276).  `balanced_row_cuts` restates that rule and additionally
rounds cuts to hack boundaries (32 rows).

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the
tests).  Rank r owns rows [r0, r1) of A and the matching slice of x.  Its rows
reference the global column range [c0, c1) -- for a banded / stencil matrix
that is the own slice plus a halo, for a general matrix it is everything
(then the exchange degenerates to an all-gather).  The local x buffer covers
[c0, c1); the shard's column indices are stored relative to c0.

Per iteration:
  1. boundary rows  = own rows some peer needs; computed FIRST
  2. exchange       = those values travel to the peers' halo regions
                      mode "nccl": batched isend/irecv on a side stream
                      mode "push": the boundary-row kernel itself stores them
                                   into the peers' buffers (CUDA IPC mapped, NVLink
                                   peer stores from the SpMV epilogue); a one-thread
                                   signal kernel publishes the step number into the
                                   neighbours' flag words and a one-thread wait kernel
                                   holds back the next step's boundary rows until
                                   every neighbour has signalled -- one stream, no
                                   collective call, no host round trip
  3. interior rows  = everything else, on the compute stream, overlapping 2.
x is double-buffered: step k reads X[k%2] and writes X[(k+1)%2].
"""
import numpy as np

HACK = 32


# ------------------------------------------------------------------ planning --
def balanced_row_cuts(irp, parts, align=HACK):
    """cuts[parts+1] over rows; greedy nnz balance as the reference's partition_csr_rows
    (src/csr.c:218-276), each cut rounded UP to a multiple of `align`.  Trailing parts may
    be empty when the matrix has fewer aligned blocks than parts."""
    irp = np.asarray(irp, dtype=np.int64)
    M = len(irp) - 1
    total = int(irp[-1])
    cuts = [0]
    target = total / parts if parts else 0.0
    running_from = 0
    for _ in range(parts - 1):
        # first row r such that nnz(rows[running_from..r]) >= target
        want = irp[running_from] + target
        r = int(np.searchsorted(irp, want, side="left"))  # irp[r] >= want  -> rows < r reach it
        r = max(r, running_from + 1) if running_from < M else M
        r = min(M, -(-r // align) * align)
        cuts.append(r)
        running_from = r
    cuts.append(M)
    return np.maximum.accumulate(np.asarray(cuts, dtype=np.int64))


def column_range(ja, r0, r1):
    """[c0, c1) touched by a shard; an empty shard needs only its own slice."""
    if len(ja) == 0:
        return r0, r1
    return int(min(int(np.min(ja)), r0)), int(max(int(np.max(ja)) + 1, r1))


class ExchangePlan:
    """Who sends which global index ranges to whom.  Built from every rank's
    (r0, r1, c0, c1); identical on all ranks."""

    def __init__(self, rank, table):
        self.rank = rank
        self.table = [tuple(int(v) for v in row) for row in table]
        r0, r1, c0, c1 = self.table[rank]
        self.r0, self.r1, self.c0, self.c1 = r0, r1, c0, c1
        self.recv = []  # (peer, g0, g1): x[g0:g1] arrives from peer
        self.send = []  # (peer, g0, g1): own y[g0:g1] goes to peer
        for p, (pr0, pr1, pc0, pc1) in enumerate(self.table):
            if p == rank:
                continue
            g0, g1 = max(c0, pr0), min(c1, pr1)
            if g0 < g1:
                self.recv.append((p, g0, g1))
            g0, g1 = max(pc0, r0), min(pc1, r1)
            if g0 < g1:
                self.send.append((p, g0, g1))
        covered = sum(g1 - g0 for _, g0, g1 in self.recv) + (r1 - r0)
        assert covered == c1 - c0, "row ranges must tile the needed column range"
        # local row cut points: [0, lo) and [hi, M) are the rows peers need
        lo = max([g1 for p, g0, g1 in self.send if g0 == r0] + [r0]) - r0
        hi = min([g0 for p, g0, g1 in self.send if g1 == r1] + [r1]) - r0
        M = r1 - r0
        inner = [s for s in self.send if not (s[1] == r0 or s[2] == r1)]
        if inner or lo >= hi:
            # a peer needs rows from the middle (or everything): no interior to overlap
            lo, hi = M, M
        self.boundary_lo, self.boundary_hi = int(lo), int(hi)

    @property
    def cuts(self):
        M = self.r1 - self.r0
        return sorted({c for c in (self.boundary_lo, self.boundary_hi) if 0 < c < M})

    @property
    def segments(self):
        """(row0, row1, is_boundary) local row segments in launch order: boundary first."""
        M = self.r1 - self.r0
        lo, hi = self.boundary_lo, self.boundary_hi
        if lo >= M:  # everything is boundary
            return [(0, M, True)] if M else []
        segs = []
        if lo > 0:
            segs.append((0, lo, True))
        if hi < M:
            segs.append((hi, M, True))
        segs.append((lo, hi, False))
        return segs

    def halo_bytes(self):
        return 8 * sum(g1 - g0 for _, g0, g1 in self.recv)

    def max_push_targets(self):
        """Largest number of peers any single boundary segment has to feed.  The fused push
        epilogue of the SpMV kernels carries two destinations; a plan that needs more (an
        all-gather over more than three ranks) has to exchange through NCCL instead."""
        worst = 0
        for r0, r1, is_boundary in self.segments:
            if not is_boundary:
                continue
            n = sum(1 for _, g0, g1 in self.send if max(g0 - self.r0, r0) < min(g1 - self.r0, r1))
            worst = max(worst, n)
        return worst


def gather_table(dist, r0, r1, c0, c1, device="cpu"):
    """all-gather of the four integers that define every rank's shard."""
    import torch
    mine = torch.tensor([r0, r1, c0, c1], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [t.cpu().tolist() for t in out]


# ----------------------------------------------------------------- execution --
class DistSpMV:
    """Iterated SpMV on one rank's shard.

    shard:  object with .M, .N (= c1 - c0) and .spmv(x, y, kernel, warps_per_block, rows, push)
            -- a CsrDevice on GPU; tests inject a CPU stand-in.
    """

    def __init__(self, dist, shard, plan, x0_own, device, mode="nccl", kernel=4, wpb=4):
        import torch
        self.torch, self.dist, self.shard, self.plan = torch, dist, shard, plan
        self.mode, self.kernel, self.wpb = mode, kernel, wpb
        self.device = device
        n_local = plan.c1 - plan.c0
        assert shard.N == n_local and shard.M == plan.r1 - plan.r0
        self.own0 = plan.r0 - plan.c0
        self.X = [torch.zeros(n_local, dtype=torch.float64, device=device) for _ in range(2)]
        self.X[0][self.own0:self.own0 + shard.M] = x0_own
        self.cuda = str(device).startswith("cuda")
        self.step_no = 0
        if self.cuda:
            # high priority: when the exchange kernel and the (SM-filling, persistent)
            # interior kernel become runnable together, the exchange gets its SM first
            self.comm = torch.cuda.Stream(priority=-1)
        self.graph, self.graph_steps = None, 0
        self._peer_ptrs = None
        self._flag = None
        if mode == "push" and plan.max_push_targets() > 2:
            # e.g. a general matrix on 4+ ranks: every rank needs every slice
            self.mode = mode = "nccl"
        if mode == "push":
            self._setup_push()
        self._initial_exchange()

    # -- helpers ---------------------------------------------------------------
    def own(self, buf):
        return self.X[buf][self.own0:self.own0 + self.shard.M]

    def _p2p_ops(self, buf):
        d, P = self.dist, self.plan
        ops = []
        for peer, g0, g1 in P.send:
            ops.append(d.P2POp(d.isend, self.X[buf][g0 - P.c0:g1 - P.c0], peer))
        for peer, g0, g1 in P.recv:
            ops.append(d.P2POp(d.irecv, self.X[buf][g0 - P.c0:g1 - P.c0], peer))
        return ops

    def _exchange_nccl(self, buf):
        ops = self._p2p_ops(buf)
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def _initial_exchange(self):
        """Halo of x_0."""
        self._exchange_nccl(0)
        if self.cuda:
            self.torch.cuda.synchronize()
        self.dist.barrier()

    def _setup_push(self):
        """Map every neighbour's two x buffers and its flag words into this process (CUDA
        IPC) so the boundary-row kernel can store into them directly and the signal kernel
        can publish this rank's epoch."""
        import ctypes as C
        from . import _lib as L
        torch, d = self.torch, self.dist
        world, rank = d.get_world_size(), d.get_rank()
        self._flags = torch.zeros(world, dtype=torch.int64, device=self.device)
        self._epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
        exported = [self.X[0], self.X[1], self._flags]
        handles = torch.zeros(len(exported) * 64, dtype=torch.uint8)
        for b, t in enumerate(exported):
            buf = (C.c_ubyte * 64)()
            if L.b200.spmv_b200_ipc_export(C.c_void_p(t.data_ptr()), buf):
                raise RuntimeError("ipc export failed: " + L.last_error())
            handles[64 * b:64 * b + 64] = torch.tensor(list(buf), dtype=torch.uint8)
        # torch's caching allocator may hand out an interior pointer of a larger block;
        # the IPC handle maps the BLOCK, so ship the offset inside it as well
        offs = torch.tensor([self._alloc_offset(t) for t in exported], dtype=torch.int64)
        hd, od = handles.to(self.device), offs.to(self.device)
        h_all = [torch.zeros_like(hd) for _ in range(world)]
        o_all = [torch.zeros_like(od) for _ in range(world)]
        d.all_gather(h_all, hd)
        d.all_gather(o_all, od)
        self._peer_ptrs = {}
        self._neighbours = sorted({p for p, _, _ in self.plan.send} | {p for p, _, _ in self.plan.recv})
        opened = {}
        for p in self._neighbours:
            ptrs = []
            for b in range(len(exported)):
                raw = bytes(h_all[p][64 * b:64 * b + 64].cpu().tolist())
                if raw not in opened:  # several tensors may live in one allocator block
                    hbuf = (C.c_ubyte * 64).from_buffer_copy(raw)
                    out = C.c_void_p()
                    if L.b200.spmv_b200_ipc_open(hbuf, C.byref(out)):
                        raise RuntimeError("ipc open failed: " + L.last_error())
                    opened[raw] = out.value
                ptrs.append(opened[raw] + int(o_all[p][b]))
            self._peer_ptrs[p] = ptrs
        n = len(self._neighbours)
        # slot [rank] of each neighbour's flags (I write), slot [p] of my flags (p writes)
        self._peer_slots = (C.c_void_p * max(n, 1))(*[self._peer_ptrs[p][2] + 8 * rank for p in self._neighbours])
        self._my_slots = (C.c_void_p * max(n, 1))(*[self._flags.data_ptr() + 8 * p for p in self._neighbours])
        torch.cuda.synchronize()
        d.barrier()

    def _alloc_offset(self, t):
        """Offset of tensor `t` inside its cudaMalloc block (the unit CUDA IPC exports)."""
        import ctypes as C
        cuda = C.CDLL("libcuda.so.1")
        base, size = C.c_uint64(), C.c_size_t()
        rc = cuda.cuMemGetAddressRange_v2(C.byref(base), C.byref(size), C.c_uint64(t.data_ptr()))
        if rc != 0:
            raise RuntimeError(f"cuMemGetAddressRange failed ({rc})")
        return t.data_ptr() - base.value

    # -- one iteration ---------------------------------------------------------
    def step(self):
        torch, P = self.torch, self.plan
        src, dst = self.step_no % 2, (self.step_no + 1) % 2
        x, y = self.X[src], self.own(dst)
        segs = P.segments
        if not self.cuda:  # CPU path of the tests: no streams
            for r0, r1, _ in segs:
                self.shard.spmv(x, y, kernel=self.kernel, warps_per_block=self.wpb, rows=(r0, r1))
            self._exchange_nccl(dst)
            self.step_no += 1
            return

        boundary = [s for s in segs if s[2]]
        interior = [s for s in segs if not s[2]]
        compute = torch.cuda.current_stream()  # the capture stream while a graph is being built
        if self.mode == "push":
            import ctypes as C
            from . import _lib as L
            st = C.c_void_p(compute.cuda_stream)
            n = len(self._neighbours)
            # every neighbour has finished the boundary rows of the previous step: its
            # pushes into X[src] have landed and it no longer reads the halo of X[dst]
            if L.b200.spmv_b200_wait_peers(C.c_void_p(self._epoch.data_ptr()), n, self._my_slots,
                                           1 << 24, C.c_void_p(self._err.data_ptr()), st):
                raise RuntimeError(L.last_error())
            for r0, r1, _ in boundary:
                push = []
                for peer, g0, g1 in P.send:
                    l0, l1 = g0 - P.r0, g1 - P.r0
                    a, b = max(l0, r0), min(l1, r1)
                    if a < b:
                        pc0 = P.table[peer][2]
                        # row a of my slice is global index P.r0 + a: its place in the peer's buffer
                        push.append((a, b, self._peer_ptrs[peer][dst] + 8 * (P.r0 + a - pc0)))
                assert len(push) <= 2, "push epilogue supports two peers per boundary segment"
                self.shard.spmv(x, y, kernel=self.kernel, warps_per_block=self.wpb, rows=(r0, r1),
                                push=push)
            if L.b200.spmv_b200_signal_peers(C.c_void_p(self._epoch.data_ptr()), n, self._peer_slots, st):
                raise RuntimeError(L.last_error())
            for r0, r1, _ in interior:
                self.shard.spmv(x, y, kernel=self.kernel, warps_per_block=self.wpb, rows=(r0, r1))
            self.step_no += 1
            return
        else:
            for r0, r1, _ in boundary:
                self.shard.spmv(x, y, kernel=self.kernel, warps_per_block=self.wpb, rows=(r0, r1))
            ev = torch.cuda.Event()
            ev.record(compute)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                ops = self._p2p_ops(dst)
                reqs = self.dist.batch_isend_irecv(ops) if ops else []
                for r in reqs:
                    r.wait()
                done = torch.cuda.Event()
                done.record(self.comm)
        for r0, r1, _ in interior:
            self.shard.spmv(x, y, kernel=self.kernel, warps_per_block=self.wpb, rows=(r0, r1))
        compute.wait_event(done)
        self.step_no += 1

    # -- CUDA graph of two consecutive steps (one per x buffer) -----------------
    def build_graph(self, steps=2):
        """Capture `steps` (even) iterations -- kernels, the side-stream exchange and the
        stream joins -- into one CUDA graph so a step costs one graph launch instead of a
        dozen host-side calls.  Call after a few eager warm-up steps (plans built, NCCL
        connections up) and on an even step number."""
        torch = self.torch
        assert self.cuda and steps % 2 == 0 and self.step_no % 2 == 0
        torch.cuda.synchronize()
        self.dist.barrier()
        saved = self.step_no
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(steps):
                self.step()
        self.step_no = saved
        self.graph, self.graph_steps = g, steps
        torch.cuda.synchronize()
        self.dist.barrier()

    def run(self, k):
        """k iterations, through the captured graph where possible."""
        while k > 0:
            if self.graph is not None and k >= self.graph_steps and self.step_no % 2 == 0:
                self.graph.replay()
                self.step_no += self.graph_steps
                k -= self.graph_steps
            else:
                self.step()
                k -= 1

    def check_errors(self):
        """Raise if a wait kernel gave up (a neighbour never signalled)."""
        if self.mode == "push" and self.cuda:
            e = int(self._err.item())
            if e:
                raise RuntimeError(f"rank {self.dist.get_rank()}: wait on neighbour slot {e - 1} timed out")

    def close(self):
        """Drop the captured graph before the process group goes away."""
        self.graph = None

    def result_own(self):
        return self.own(self.step_no % 2)
