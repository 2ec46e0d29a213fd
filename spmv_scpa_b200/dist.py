"""ctypes mirror of the multi-GPU section of include/spmv_b200.h: x_{k+1} = A x_k, rows partitioned
over the GPUs of one box.  Everything that happens per step -- the partition rule, the exchange
plan, the stream orchestration, the CUDA graph, the IPC wiring, the push epilogue and the NCCL
fallback -- lives in csrc/dist.cu; this module only (a) exposes the planning calls to the tests and
(b) moves the 512-byte connection blobs between the rank processes with torch.distributed, which
is plumbing, not the product.

NEW SURFACE -- the reference is single-GPU.  Its partition rule (contiguous row
ranges with balanced nnz, a range closed when its running nnz reaches total/parts:
partition_csr_rows, reference src/csr.c:218-276) is what `balanced_row_cuts` computes, with cuts
rounded up to hack boundaries.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import structs as S

HACK = 32
AUTO, PUSH, NCCL = S.DIST_AUTO, S.DIST_PUSH, S.DIST_NCCL
MODES = {"auto": AUTO, "push": PUSH, "nccl": NCCL}


def _check(rc, what):
    if rc:
        raise RuntimeError(f"{what}: {L.last_error()} (rc={rc})")


# ------------------------------------------------------------------ planning --
def balanced_row_cuts(irp, parts, align=HACK):
    """cuts[parts+1] over rows (spmv_b200_partition_rows): the reference's greedy nnz balance,
    each cut rounded UP to a multiple of `align`.  Trailing parts may be empty."""
    irp = np.ascontiguousarray(irp)
    if irp.dtype not in (np.int32, np.int64):
        irp = irp.astype(np.int64)
    cuts = np.zeros(parts + 1, np.int64)
    _check(L.b200.spmv_b200_partition_rows(len(irp) - 1, irp.ctypes.data_as(C.c_void_p), irp.dtype.itemsize,
                                           parts, align, cuts.ctypes.data_as(L.c_i64p)), "partition_rows")
    return cuts


def shard_scan(r0, r1, irp_rows, ja_rows):
    """Descriptor of rows [r0, r1): irp_rows = their r1-r0+1 offsets, ja_rows = their (global) columns."""
    irp = np.ascontiguousarray(irp_rows)
    if irp.dtype not in (np.int32, np.int64):
        irp = irp.astype(np.int64)
    ja = np.ascontiguousarray(ja_rows, np.int32)
    d = S.shard_desc()
    _check(L.b200.spmv_b200_shard_scan(r0, r1, irp.ctypes.data_as(C.c_void_p), irp.dtype.itemsize,
                                       ja.ctypes.data_as(L.c_ip), C.byref(d)), "shard_scan")
    return d


def stencil27_shard_desc(nx, ny, nz, z0, z1):
    d = S.shard_desc()
    _check(L.b200.spmv_b200_stencil27_shard_desc(nx, ny, nz, z0, z1, C.byref(d)), "stencil27_shard_desc")
    return d


def desc_tuple(d):
    return (d.r0, d.r1, d.c0, d.c1, d.read_lo, d.read_hi)


def make_desc(r0, r1, c0, c1, read_lo=None, read_hi=None):
    """Descriptor from plain numbers.  Without read_lo / read_hi the shard is taken to read its
    halo only from the rows its neighbours need (structurally symmetric matrices)."""
    d = S.shard_desc(r0, r1, c0, c1, 0 if read_lo is None else read_lo, (r1 - r0) if read_hi is None else read_hi)
    return d


class Plan:
    """spmv_b200_dist_plan of one rank (identical inputs on every rank give consistent plans)."""

    def __init__(self, rank, table, mode=AUTO):
        self.table = (S.shard_desc * len(table))(*[t if isinstance(t, S.shard_desc) else make_desc(*t) for t in table])
        self.c = S.dist_plan()
        _check(L.b200.spmv_b200_dist_make_plan(rank, len(table), self.table, MODES.get(mode, mode), C.byref(self.c)),
               "dist_make_plan")
        me = self.table[rank]
        self.rank, self.world = rank, len(table)
        self.r0, self.r1, self.c0, self.c1 = me.r0, me.r1, me.c0, me.c1
        self.send = [(x.peer, x.g0, x.g1) for x in self.c.send[:self.c.n_send]]
        self.recv = [(x.peer, x.g0, x.g1) for x in self.c.recv[:self.c.n_recv]]
        self.boundary_lo, self.boundary_hi = self.c.boundary_lo, self.c.boundary_hi
        self.cuts = list(self.c.cuts[:self.c.n_cuts])
        self.mode = {PUSH: "push", NCCL: "nccl"}[self.c.mode]
        self.all_gather = bool(self.c.all_gather)

    @property
    def segments(self):
        """(row0, row1, is_boundary) local row segments in launch order: boundary first."""
        M = self.r1 - self.r0
        lo, hi = self.boundary_lo, self.boundary_hi
        if M <= 0:
            return []
        if lo >= M:
            return [(0, M, True)]
        segs = []
        if lo > 0:
            segs.append((0, lo, True))
        if hi < M:
            segs.append((hi, M, True))
        segs.append((lo, hi, False))
        return segs

    def halo_bytes(self):
        return int(self.c.halo_bytes)


def gather_table(dist, desc, device="cpu"):
    """all-gather of the six integers that define every rank's shard."""
    import torch
    mine = torch.tensor(desc_tuple(desc), dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [make_desc(*[int(v) for v in t.cpu().tolist()]) for t in out]


# ----------------------------------------------------------------- execution --
class DistSpMV:
    """One rank of the iterated SpMV (spmv_b200_dist).  `shard` is a CsrDevice created with
    col_offset = c0, n_local = c1 - c0 and cuts = plan.cuts."""

    def __init__(self, dist, shard, plan, kernel=4, wpb=4, device=None):
        import torch
        self.torch, self.dist, self.shard, self.plan = torch, dist, shard, plan
        self._h = L.b200.spmv_b200_dist_create(C.byref(plan.c), plan.table, shard._h, kernel, wpb)
        if not self._h:
            raise RuntimeError("dist_create: " + L.last_error())
        self._h = C.c_void_p(self._h)
        self.M = plan.r1 - plan.r0
        world = plan.world
        if world > 1:
            blob = (C.c_ubyte * S.DIST_BLOB_BYTES)()
            _check(L.b200.spmv_b200_dist_export(self._h, blob), "dist_export")
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            out = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(out, mine)
            allb = b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)
            buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
            _check(L.b200.spmv_b200_dist_connect(self._h, buf), "dist_connect")
        self.mode = {PUSH: "push", NCCL: "nccl"}[L.b200.spmv_b200_dist_mode(self._h)]

    # -- collective: the whole job is idle (sync + barrier) when the halos are rewritten --------
    def set_x(self, x_own):
        """x_0: own slice, a float64 numpy array or CUDA tensor of M entries."""
        torch = self.torch
        self.sync()
        self.dist.barrier()
        if isinstance(x_own, np.ndarray):
            x_own = np.ascontiguousarray(x_own, np.float64)
            assert x_own.shape == (self.M,)
            ptr = x_own.ctypes.data_as(C.c_void_p)
        else:
            assert x_own.dtype == torch.float64 and x_own.is_contiguous() and x_own.numel() == self.M
            torch.cuda.current_stream().synchronize()
            ptr = C.c_void_p(x_own.data_ptr())
        _check(L.b200.spmv_b200_dist_set_x(self._h, ptr), "dist_set_x")
        self.sync()            # the copy of x_own has been consumed (the caller may free it)
        self.dist.barrier()

    def iterate(self, k):
        _check(L.b200.spmv_b200_dist_iterate(self._h, k), "dist_iterate")

    def time(self, k, reps):
        """`reps` event-timed regions of k steps each; ms per region (this rank)."""
        ms = (C.c_double * reps)()
        _check(L.b200.spmv_b200_dist_time(self._h, k, reps, ms), "dist_time")
        return list(ms)

    def sync(self):
        _check(L.b200.spmv_b200_dist_sync(self._h), "dist_sync")

    @property
    def stream(self):
        """The rank's CUDA stream as a torch stream (to order torch collectives against the steps)."""
        return self.torch.cuda.ExternalStream(L.b200.spmv_b200_dist_stream(self._h))

    @property
    def steps(self):
        return L.b200.spmv_b200_dist_steps(self._h)

    @property
    def has_graph(self):
        return bool(L.b200.spmv_b200_dist_has_graph(self._h))

    def x_ptr(self):
        """Device pointer of the own slice of the current x."""
        return L.b200.spmv_b200_dist_x(self._h)

    def xlocal_ptr(self):
        return L.b200.spmv_b200_dist_xlocal(self._h)

    def result_own(self):
        out = np.zeros(self.M, np.float64)
        _check(L.b200.spmv_b200_dist_get_x(self._h, out.ctypes.data_as(L.c_dp)), "dist_get_x")
        return out

    def close(self):
        if self._h:
            L.b200.spmv_b200_dist_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistGroup:
    """One process driving n GPUs (spmv_b200_dist_group_*): host x in, k steps, host x out."""

    def __init__(self, handle, N):
        if not handle:
            raise RuntimeError("dist_group: " + L.last_error())
        self._h = C.c_void_p(handle)
        self.N = N
        self.n = L.b200.spmv_b200_dist_group_size(self._h)

    @classmethod
    def from_csr(cls, A, n_gpus, kernel=4, wpb=4, mode=AUTO):
        return cls(L.b200.spmv_b200_dist_group_create(A.ptr, n_gpus, kernel, wpb, MODES.get(mode, mode)), A.N)

    @classmethod
    def stencil27(cls, nx, ny, nz, n_gpus, kernel=4, wpb=4, mode=AUTO):
        return cls(L.b200.spmv_b200_dist_group_stencil27(nx, ny, nz, n_gpus, kernel, wpb, MODES.get(mode, mode)),
                   nx * ny * nz)

    def mode(self, rank=0):
        return {PUSH: "push", NCCL: "nccl"}[L.b200.spmv_b200_dist_mode(L.b200.spmv_b200_dist_group_rank(self._h, rank))]

    def has_graph(self, rank=0):
        return bool(L.b200.spmv_b200_dist_has_graph(L.b200.spmv_b200_dist_group_rank(self._h, rank)))

    def set_x(self, x):
        x = np.ascontiguousarray(x, np.float64)
        assert x.shape == (self.N,)
        _check(L.b200.spmv_b200_dist_group_set_x(self._h, x.ctypes.data_as(L.c_dp)), "dist_group_set_x")

    def iterate(self, k):
        ms = C.c_double()
        _check(L.b200.spmv_b200_dist_group_iterate(self._h, k, C.byref(ms)), "dist_group_iterate")
        return ms.value

    def get_x(self):
        out = np.zeros(self.N, np.float64)
        _check(L.b200.spmv_b200_dist_group_get_x(self._h, out.ctypes.data_as(L.c_dp)), "dist_group_get_x")
        return out

    def close(self):
        if self._h:
            L.b200.spmv_b200_dist_group_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
