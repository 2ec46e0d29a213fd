"""TEST INFRASTRUCTURE -- Python face of the oracle.  Not product code.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module; nothing under spmv_scpa_b200/ does.

Two checkers live here:

  port  liboracle.so, the plain-C restatement in oracle.c (always available
        once built; strict IEEE, serial unless stated);
  ref   oracle/_ref/libspmv_ref_host*.so, the reference's OWN host sources
        (src/{mmio,utils,vector,logger,csr,hll}.c) compiled unmodified by
        oracle/Makefile with the reference's flags.  Built only where
        /root/reference exists; the prebuilt file travels to the GPU box.

The ctypes structs below restate the reference ABI (include/csr.h:7-13,
include/hll.h:13-37, include/utils.h:32-47, include/vector.h:6-9) on their
own so the oracle does not depend on the product package.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HACK = 32

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def _ci(a):
    return a.ctypes.data_as(_ip)


def _cd(a):
    return a.ctypes.data_as(_dp)


def _ci64(a):
    return a.ctypes.data_as(_i64p)


def aligned(n, dtype):
    """64-byte aligned array (the reference's OpenMP path asserts it, src/csr.c:309-311)."""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + 64, dtype=np.uint8)
    shift = (-raw.ctypes.data) % 64
    return raw[shift:shift + n * item].view(dtype)


def aligned_copy(a, dtype):
    out = aligned(len(a), dtype)
    out[:] = a
    return out


# ------------------------------------------------------------------ the port --
_port = None


def port():
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
        lib = C.CDLL(path)
        lib.orc_load_mtx.restype = C.c_int
        lib.orc_load_mtx.argtypes = [C.c_char_p, _ip, _ip, _ip, C.POINTER(_ip), C.POINTER(_ip),
                                     C.POINTER(_dp)]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_hll_shape.restype = C.c_int64
        lib.orc_hll_shape.argtypes = [C.c_int, _ip, _ip, _ip, _ip, _i64p]
        lib.orc_hll_pack.argtypes = [C.c_int, _ip, _ip, _dp, C.c_int, _i64p, _ip, _dp]
        lib.orc_hll_patch_pads.argtypes = [C.c_int, _ip, _ip, _i64p, C.c_int, _ip]
        lib.orc_csr_spmv.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp]
        lib.orc_csr_spmv64.argtypes = [C.c_int64, _i64p, _ip, _dp, _dp, _dp]
        lib.orc_csr_abs_bound.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp]
        lib.orc_hll_spmv.argtypes = [C.c_int, _ip, _ip, _i64p, C.c_int, _ip, _dp, _dp, _dp]
        lib.orc_partition_rows.restype = C.c_int
        lib.orc_partition_rows.argtypes = [C.c_int, _ip, C.c_int, _ip]
        lib.orc_csr_spmv_timed.restype = C.c_double
        lib.orc_csr_spmv_timed.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int]
        lib.orc_max_threads.restype = C.c_int
        lib.orc_stencil27_nnz.restype = C.c_int64
        lib.orc_stencil27_nnz.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]
        lib.orc_gen_stencil27_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _ip, _ip, _dp]
        lib.orc_gen_poisson2d.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp]
        lib.orc_gen_uniform_rows.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, _ip, _ip, _dp]
        lib.orc_gen_rmat.restype = C.c_int
        lib.orc_gen_rmat.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                     _ip, _ip, _dp]
        _port = lib
    return _port


def load_mtx(path):
    """(M, N, IRP, JA, AS) following reference src/csr.c:31-171; OSError(errno) on failure."""
    M, N, NZ = C.c_int(), C.c_int(), C.c_int()
    irp, ja, as_ = _ip(), _ip(), _dp()
    rc = port().orc_load_mtx(os.fsencode(path), C.byref(M), C.byref(N), C.byref(NZ), C.byref(irp),
                             C.byref(ja), C.byref(as_))
    if rc:
        raise OSError(-rc, os.strerror(-rc))
    IRP = np.ctypeslib.as_array(irp, shape=(M.value + 1,)).copy()
    JA = np.ctypeslib.as_array(ja, shape=(max(NZ.value, 1),))[:NZ.value].copy()
    AS = np.ctypeslib.as_array(as_, shape=(max(NZ.value, 1),))[:NZ.value].copy()
    for p in (irp, ja, as_):
        port().orc_free(C.cast(p, C.c_void_p))
    return M.value, N.value, IRP, JA, AS


def csr_to_hll(M, IRP, JA, AS, col_major):
    """rows, width, nz, off, ja, as  (reference src/hll.c:19-95)."""
    nb = (M + HACK - 1) // HACK
    IRP = np.ascontiguousarray(IRP, np.int32)
    JA = np.ascontiguousarray(JA, np.int32)
    AS = np.ascontiguousarray(AS, np.float64)
    rows = np.zeros(nb, np.int32)
    width = np.zeros(nb, np.int32)
    nz = np.zeros(nb, np.int32)
    off = np.zeros(nb + 1, np.int64)
    slots = port().orc_hll_shape(M, _ci(IRP), _ci(rows), _ci(width), _ci(nz), _ci64(off))
    ja = np.zeros(slots, np.int32)
    as_ = np.zeros(slots, np.float64)
    port().orc_hll_pack(M, _ci(IRP), _ci(JA), _cd(AS), int(bool(col_major)), _ci64(off), _ci(ja),
                        _cd(as_))
    return rows, width, nz, off, ja, as_


def hll_patch_pads(rows, width, off, col_major, ja):
    """Device padding convention (reference src/cuda_hll.cu:173-195), returns a patched copy."""
    out = np.ascontiguousarray(ja, np.int32).copy()
    port().orc_hll_patch_pads(len(rows), _ci(rows), _ci(width), _ci64(off), int(bool(col_major)),
                              _ci(out))
    return out


def hll_device_layout(M, IRP, JA, AS):
    """What libspmv_b200 must hold in HBM for this matrix (include/cuda_hll.h): per hack,
    column-major with stride 32, pads patched, rows beyond M all pads.
    Returns hoff[nb+1], ja[slots32], as[slots32]."""
    rows, width, nz, off, ja, as_ = csr_to_hll(M, IRP, JA, AS, True)
    ja = hll_patch_pads(rows, width, off, True, ja)
    nb = len(rows)
    hoff = np.zeros(nb + 1, np.int64)
    hoff[1:] = np.cumsum(32 * width.astype(np.int64))
    dja = np.zeros(hoff[-1], np.int32)
    das = np.zeros(hoff[-1], np.float64)
    for b in range(nb):
        r, w = int(rows[b]), int(width[b])
        if w == 0:
            continue
        src_j = ja[off[b]:off[b + 1]].reshape(w, r)
        src_a = as_[off[b]:off[b + 1]].reshape(w, r)
        dst_j = dja[hoff[b]:hoff[b + 1]].reshape(w, 32)
        dst_a = das[hoff[b]:hoff[b + 1]].reshape(w, 32)
        dst_j[:, :r] = src_j
        dst_a[:, :r] = src_a
    return hoff, dja, das


def sellp_layout(M, N, IRP, JA, AS, K, sigma, max_row):
    """What libspmv_b200 must hold in HBM for the SELL-P form of this matrix
    (spmv_scpa_b200/csrc/sell_kernels.cuh; no reference counterpart -- the rule restated here is the
    one documented in that header, written independently in numpy):
      panel p = columns [N*p//K, N*(p+1)//K); rows longer than max_row belong to no slice;
      inside every window of `sigma` rows, rows are ordered by their entry count in the panel,
      descending, ties by row index; slice = 32 consecutive rows of that order, width = count of
      its first row; column-major with stride 32; pads: value 0.0, index = previous valid column
      of the row inside the panel, or the panel's first column.
    Returns soff[K, S+1], perm[K, S*32], ja[slots], as[slots]."""
    IRP = np.asarray(IRP, np.int64)
    JA = np.asarray(JA, np.int64)
    AS = np.asarray(AS, np.float64)
    S = (M + 31) // 32
    pc = [N * p // K for p in range(K + 1)]
    lens = np.diff(IRP)
    perm = np.full((K, S * 32), -1, np.int32)
    width = np.zeros((K, S), np.int64)
    rows_entries = {}
    for p in range(K):
        cnt = np.zeros(M, np.int64)
        for r in range(M):
            if lens[r] > max_row:
                cnt[r] = -1
                continue
            cols = JA[IRP[r]:IRP[r + 1]]
            keep = np.nonzero((cols >= pc[p]) & (cols < pc[p + 1]))[0] if K > 1 else np.arange(len(cols))
            rows_entries[(p, r)] = IRP[r] + keep
            cnt[r] = len(keep)
        for w0 in range(0, M, sigma):
            w1 = min(M, w0 + sigma)
            rows = [r for r in range(w0, w1) if cnt[r] >= 0]
            rows.sort(key=lambda r: (-cnt[r], r))
            perm[p, w0:w0 + len(rows)] = rows
            for s in range(w0 // 32, (w1 + 31) // 32):
                first = perm[p, s * 32]
                width[p, s] = cnt[first] if first >= 0 else 0
    soff = np.zeros((K, S + 1), np.int64)
    run = 0
    for p in range(K):
        for s in range(S):
            soff[p, s] = run
            run += 32 * width[p, s]
        soff[p, S] = run
    ja = np.zeros(run, np.int32)
    as_ = np.zeros(run, np.float64)
    for p in range(K):
        for s in range(S):
            w = int(width[p, s])
            blk_j = ja[soff[p, s]:soff[p, s + 1]].reshape(w, 32)
            blk_a = as_[soff[p, s]:soff[p, s + 1]].reshape(w, 32)
            for i in range(32):
                r = perm[p, s * 32 + i]
                ks = rows_entries[(p, r)] if r >= 0 else np.zeros(0, np.int64)
                n = len(ks)
                blk_j[:n, i] = JA[ks]
                blk_a[:n, i] = AS[ks]
                blk_j[n:, i] = JA[ks[-1]] if n else pc[p]
    return soff, perm, ja, as_


def sellv_layout(M, IRP, JA, AS, sigma, chunk):
    """SELL-P with virtual rows (one panel; sell_kernels.cuh): every row longer than `chunk` entries is
    cut into pieces of at most `chunk` entries, in order; the pieces ("virtual rows", numbered row by
    row) are ordered inside windows of `sigma` virtual rows by length, descending, ties by number;
    a lane's destination is the row (unsplit), or -2-i for piece i of the split rows (pieces numbered
    in row order), or -1.  Returns soff[1, S+1], dest[1, S*32], ja, as, split_row, split_first."""
    IRP = np.asarray(IRP, np.int64)
    JA = np.asarray(JA, np.int64)
    AS = np.asarray(AS, np.float64)
    pieces = []            # (first entry, length, destination)
    split_row, split_first, p = [], [], 0
    for r in range(M):
        L = int(IRP[r + 1] - IRP[r])
        if L <= chunk:
            pieces.append((int(IRP[r]), L, r))
            continue
        split_row.append(r)
        split_first.append(p)
        for j0 in range(0, L, chunk):
            pieces.append((int(IRP[r]) + j0, min(chunk, L - j0), -2 - p))
            p += 1
    split_first.append(p)
    V = len(pieces)
    S = (V + 31) // 32
    order = np.full(S * 32, -1, np.int64)
    for w0 in range(0, V, sigma):
        w1 = min(V, w0 + sigma)
        vs = sorted(range(w0, w1), key=lambda v: (-pieces[v][1], v))
        order[w0:w0 + len(vs)] = vs
    width = np.array([pieces[order[s * 32]][1] if order[s * 32] >= 0 else 0 for s in range(S)], np.int64)
    soff = np.zeros((1, S + 1), np.int64)
    soff[0, 1:] = np.cumsum(32 * width)
    dest = np.full((1, S * 32), -1, np.int32)
    ja = np.zeros(soff[0, -1], np.int32)
    as_ = np.zeros(soff[0, -1], np.float64)
    for s in range(S):
        w = int(width[s])
        bj = ja[soff[0, s]:soff[0, s + 1]].reshape(w, 32)
        ba = as_[soff[0, s]:soff[0, s + 1]].reshape(w, 32)
        for i in range(32):
            v = order[s * 32 + i]
            if v < 0:
                continue
            k0, n, d = pieces[v]
            dest[0, s * 32 + i] = d
            bj[:n, i] = JA[k0:k0 + n]
            ba[:n, i] = AS[k0:k0 + n]
            bj[n:, i] = JA[k0 + n - 1] if n else 0
    return soff, dest, ja, as_, np.array(split_row, np.int32), np.array(split_first, np.int32)


def csr_spmv(M, IRP, JA, AS, x):
    """y = A x, strict left-to-right FP64 (reference src/csr.c:201-216)."""
    IRP = np.ascontiguousarray(IRP)
    JA = np.ascontiguousarray(JA, np.int32)
    AS = np.ascontiguousarray(AS, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(M, np.float64)
    if IRP.dtype == np.int64:
        port().orc_csr_spmv64(M, _ci64(IRP), _ci(JA), _cd(AS), _cd(x), _cd(y))
    else:
        IRP = IRP.astype(np.int32, copy=False)
        port().orc_csr_spmv(M, _ci(IRP), _ci(JA), _cd(AS), _cd(x), _cd(y))
    return y


def csr_abs_bound(M, IRP, JA, AS, x):
    """bound_i = sum_j |a_ij x_j| (scale of the north-star tolerance)."""
    IRP = np.ascontiguousarray(IRP)
    if IRP.dtype == np.int64:
        # numpy path for 64-bit offsets (tests only use it on small shards)
        prod = np.abs(np.asarray(AS) * np.asarray(x)[np.asarray(JA)])
        cs = np.concatenate([[0.0], np.cumsum(prod)])
        return cs[IRP[1:]] - cs[IRP[:-1]]
    IRP = IRP.astype(np.int32, copy=False)
    JA = np.ascontiguousarray(JA, np.int32)
    AS = np.ascontiguousarray(AS, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    b = np.zeros(M, np.float64)
    port().orc_csr_abs_bound(M, _ci(IRP), _ci(JA), _cd(AS), _cd(x), _cd(b))
    return b


def hll_spmv(M, rows, width, off, col_major, ja, as_, x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(len(rows) * HACK, np.float64)
    port().orc_hll_spmv(len(rows), _ci(rows), _ci(width), _ci64(off), int(bool(col_major)),
                        _ci(np.ascontiguousarray(ja, np.int32)),
                        _cd(np.ascontiguousarray(as_, np.float64)), _cd(x), _cd(y))
    return y[:M]


# ---- synthetic inputs, oracle side (same definitions as include/spmv_gen.h) ----
def gen_stencil27_rows(nx, ny, nz, row0, row1):
    """(M, N, IRP, JA, AS) of rows [row0,row1) of the 27-point stencil, global columns."""
    nnz = port().orc_stencil27_nnz(nx, ny, nz, row0, row1)
    assert nnz < 2**31
    M = row1 - row0
    IRP, JA, AS = aligned(M + 1, np.int32), aligned(nnz, np.int32), aligned(nnz, np.float64)
    port().orc_gen_stencil27_rows(nx, ny, nz, row0, row1, _ci(IRP), _ci(JA), _cd(AS))
    return M, nx * ny * nz, IRP, JA, AS


def gen_stencil27(nx, ny, nz):
    return gen_stencil27_rows(nx, ny, nz, 0, nx * ny * nz)


def gen_poisson2d(nx, ny):
    n = nx * ny
    nnz = 5 * n - 2 * nx - 2 * ny
    IRP, JA, AS = aligned(n + 1, np.int32), aligned(nnz, np.int32), aligned(nnz, np.float64)
    port().orc_gen_poisson2d(nx, ny, _ci(IRP), _ci(JA), _cd(AS))
    return n, n, IRP, JA, AS


def gen_uniform_rows(n, k, seed, row0, row1):
    M = row1 - row0
    IRP, JA, AS = aligned(M + 1, np.int32), aligned(M * k, np.int32), aligned(M * k, np.float64)
    port().orc_gen_uniform_rows(n, k, seed, row0, row1, _ci(IRP), _ci(JA), _cd(AS))
    return M, n, IRP, JA, AS


def gen_rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=42):
    n = 1 << scale
    m = n * edge_factor
    IRP, JA, AS = aligned(n + 1, np.int32), aligned(m, np.int32), aligned(m, np.float64)
    rc = port().orc_gen_rmat(scale, edge_factor, a, b, c, seed, _ci(IRP), _ci(JA), _cd(AS))
    if rc:
        raise MemoryError("orc_gen_rmat")
    return n, n, IRP, JA, AS


def partition_rows(M, IRP, parts):
    IRP = np.ascontiguousarray(IRP, np.int32)
    cut = np.zeros(parts + 1, np.int32)
    used = port().orc_partition_rows(M, _ci(IRP), parts, _ci(cut))
    return cut[:used + 1].copy()


def csr_spmv_timed(M, IRP, JA, AS, x, threads=1):
    """(milliseconds, y) of the port's CSR loop with `threads` OpenMP threads."""
    y = np.zeros(M, np.float64)
    ms = port().orc_csr_spmv_timed(M, _ci(IRP), _ci(JA), _cd(AS), _cd(x), _cd(y), threads)
    return ms, y


def max_threads():
    return port().orc_max_threads()


def check_tolerance(y, y_ref, bound, rel=1e-12):
    """North-star gate: |y_i - y_ref_i| <= rel * sum_j |a_ij x_j| for every row.
    Returns (ok, worst_ratio) with worst_ratio = max_i |dy_i| / (rel*bound_i)."""
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    err = np.abs(y - y_ref)
    lim = rel * np.asarray(bound, np.float64)
    bad = err > lim
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(lim > 0, err / lim, np.where(err > 0, np.inf, 0.0))
    worst = float(ratio.max()) if len(ratio) else 0.0
    return (not bool(bad.any())), worst


# ------------------------------------------------------- the reference itself --
class _vec(C.Structure):
    _fields_ = [("len", C.c_size_t), ("data", _dp)]


class _bench(C.Structure):
    _fields_ = [("duration_ms", C.c_double), ("gflops", C.c_double), ("data", _vec)]


class _bench_omp(C.Structure):
    _fields_ = [("bench", _bench), ("name", C.c_char * 64), ("num_threads", C.c_int)]


class _csr(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("M", C.c_int), ("N", C.c_int), ("NZ", C.c_int),
                ("IRP", _ip), ("JA", _ip), ("AS", _dp)]


class _blk(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("NZ", C.c_int), ("max_NZ", C.c_int), ("JA", _ip),
                ("AS", _dp)]


class _hll(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("M", C.c_int), ("N", C.c_int), ("NZ", C.c_int),
                ("hack_size", C.c_int), ("num_blocks", C.c_int), ("blocks", C.POINTER(_blk))]


_ref = None
_ref_kind = None


def _native_ok(path):
    """Run a tiny SpMV through `path` in a child process: a SIGILL (different CPU than
    the build host) must not take the caller down."""
    code = ("import ctypes,sys; l=ctypes.CDLL(sys.argv[1]); "
            "l.aligned_malloc.restype=ctypes.c_void_p; p=l.aligned_malloc(64); print('ok')")
    try:
        r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, timeout=60)
        return r.returncode == 0 and b"ok" in r.stdout
    except Exception:
        return False


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libspmv_ref_host.so"))


def ref():
    """The reference host library (ctypes), or FileNotFoundError."""
    global _ref, _ref_kind
    if _ref is None:
        base = os.path.join(HERE, "_ref", "libspmv_ref_host.so")
        native = os.path.join(HERE, "_ref", "libspmv_ref_host_native.so")
        if not os.path.exists(base):
            raise FileNotFoundError(f"{base} missing: run `make -C oracle` where /root/reference exists")
        path, _ref_kind = base, "x86-64-v3"
        if os.environ.get("SPMV_ORACLE_NATIVE", "1") == "1" and os.path.exists(native) and _native_ok(native):
            path, _ref_kind = native, "native"
        lib = C.CDLL(path)
        lib.io_load_csr.restype = C.c_void_p
        lib.io_load_csr.argtypes = [C.c_char_p]
        lib.csr_free.argtypes = [C.c_void_p]
        lib.csr_to_hll.restype = C.c_void_p
        lib.csr_to_hll.argtypes = [C.POINTER(_csr), C.c_bool]
        lib.hll_free.argtypes = [C.c_void_p]
        lib.bench_csr_serial.argtypes = [C.POINTER(_csr), _dp, C.POINTER(_bench)]
        lib.bench_hll_serial.argtypes = [C.POINTER(_hll), _dp, C.POINTER(_bench)]
        lib.bench_csr_omp_guided.argtypes = [C.POINTER(_csr), _dp, C.POINTER(_bench_omp)]
        lib.bench_csr_omp_nnz_balancing.argtypes = [C.POINTER(_csr), _dp, C.POINTER(_bench_omp)]
        lib.bench_hll_omp.argtypes = [C.POINTER(_hll), _dp, C.POINTER(_bench_omp)]
        lib.vec_put.argtypes = [C.POINTER(_vec)]
        lib.vec_create.restype = _vec
        lib.vec_create.argtypes = [C.c_size_t]
        lib.vec_fill_random.argtypes = [C.POINTER(_vec)]
        _ref = lib
    return _ref


def ref_kind():
    ref()
    return _ref_kind


def _is_err(addr):
    return addr is None or addr == 0 or addr > (1 << 64) - 4096


def ref_load_mtx(path):
    """The reference's io_load_csr (src/csr.c:31-171) -> (M, N, IRP, JA, AS) copies."""
    addr = ref().io_load_csr(os.fsencode(path))
    if _is_err(addr):
        err = (1 << 64) - addr if addr else 12
        raise OSError(err, os.strerror(err))
    A = _csr.from_address(addr)
    M, N, NZ = A.M, A.N, A.NZ
    IRP = np.ctypeslib.as_array(A.IRP, shape=(M + 1,)).copy()
    JA = np.ctypeslib.as_array(A.JA, shape=(max(NZ, 1),))[:NZ].copy()
    AS = np.ctypeslib.as_array(A.AS, shape=(max(NZ, 1),))[:NZ].copy()
    name = A.name.decode()
    ref().csr_free(addr)
    return M, N, IRP, JA, AS, name


class RefCsr:
    """A sparse_csr for the reference library over aligned copies of numpy arrays."""

    def __init__(self, M, N, IRP, JA, AS, name="m"):
        self.IRP = aligned_copy(IRP, np.int32)
        self.JA = aligned_copy(JA, np.int32)
        self.AS = aligned_copy(AS, np.float64)
        self.st = _csr()
        self.st.name = name.encode()[:63]
        self.st.M, self.st.N, self.st.NZ = M, N, len(JA)
        self.st.IRP, self.st.JA, self.st.AS = _ci(self.IRP), _ci(self.JA), _cd(self.AS)

    @property
    def ptr(self):
        return C.byref(self.st)


def ref_csr_to_hll(A: RefCsr, col_major):
    """The reference's csr_to_hll (src/hll.c:19-95), flattened like csr_to_hll() above."""
    addr = ref().csr_to_hll(A.ptr, bool(col_major))
    if _is_err(addr):
        raise MemoryError("reference csr_to_hll failed")
    H = _hll.from_address(addr)
    nb = H.num_blocks
    rows = np.zeros(nb, np.int32)
    width = np.zeros(nb, np.int32)
    nz = np.zeros(nb, np.int32)
    off = np.zeros(nb + 1, np.int64)
    ja, as_ = [], []
    for b in range(nb):
        blk = H.blocks[b]
        n = blk.M * blk.max_NZ
        rows[b], width[b], nz[b] = blk.M, blk.max_NZ, blk.NZ
        off[b + 1] = off[b] + n
        if n:
            ja.append(np.ctypeslib.as_array(blk.JA, shape=(n,)).copy())
            as_.append(np.ctypeslib.as_array(blk.AS, shape=(n,)).copy())
    meta = dict(M=H.M, N=H.N, NZ=H.NZ, hack_size=H.hack_size, num_blocks=nb)
    ref().hll_free(addr)
    JA = np.concatenate(ja) if ja else np.zeros(0, np.int32)
    AS = np.concatenate(as_) if as_ else np.zeros(0, np.float64)
    return rows, width, nz, off, JA, AS, meta


def _take(bench):
    n = bench.data.len
    y = np.ctypeslib.as_array(bench.data.data, shape=(max(n, 1),))[:n].copy()
    ref().vec_put(C.byref(bench.data))
    return y


def ref_csr_serial(A: RefCsr, x):
    """The parity oracle proper: the reference's serial CSR SpMV (src/csr.c:201-216,
    via bench_csr_serial :342-344).  Returns (ms, y)."""
    xa = aligned_copy(x, np.float64)
    b = _bench()
    rc = ref().bench_csr_serial(A.ptr, _cd(xa), C.byref(b))
    if rc:
        raise OSError(-rc, "bench_csr_serial")
    return b.duration_ms, _take(b)


def ref_csr_omp(A: RefCsr, x, threads, schedule="guided"):
    """Reference OpenMP CSR (src/csr.c:278-339).  The reference asserts
    threads <= omp_get_max_threads(); the caller must have OMP_NUM_THREADS set accordingly."""
    xa = aligned_copy(x, np.float64)
    b = _bench_omp()
    b.num_threads = threads
    fn = ref().bench_csr_omp_guided if schedule == "guided" else ref().bench_csr_omp_nnz_balancing
    rc = fn(A.ptr, _cd(xa), C.byref(b))
    if rc:
        raise OSError(-rc, "bench_csr_omp")
    return b.bench.duration_ms, _take(b.bench), b.num_threads


def ref_hll_bench(A: RefCsr, x, threads):
    """The reference's HLL CPU paths on its own row-major packing of A: serial
    (src/hll.c:127-150 via bench_hll_serial) and OpenMP guided over hacks (:178-211 via
    bench_hll_omp).  Returns (serial_ms, omp_ms, y_serial)."""
    addr = ref().csr_to_hll(A.ptr, False)
    if _is_err(addr):
        raise MemoryError("reference csr_to_hll failed")
    try:
        H = C.cast(C.c_void_p(addr), C.POINTER(_hll))
        xa = aligned_copy(x, np.float64)
        b = _bench()
        rc = ref().bench_hll_serial(H, _cd(xa), C.byref(b))
        if rc:
            raise OSError(-rc, "bench_hll_serial")
        serial_ms, y = b.duration_ms, _take(b)
        bo = _bench_omp()
        bo.num_threads = threads
        rc = ref().bench_hll_omp(H, _cd(xa), C.byref(bo))
        if rc:
            raise OSError(-rc, "bench_hll_omp")
        omp_ms = bo.bench.duration_ms
        _take(bo.bench)
    finally:
        ref().hll_free(addr)
    return serial_ms, omp_ms, y


def ref_rand_x(n):
    """x as the reference makes it: vec_fill_random -> rand()/RAND_MAX (src/vector.c:36-41)."""
    v = ref().vec_create(n)
    ref().vec_fill_random(C.byref(v))
    out = np.ctypeslib.as_array(v.data, shape=(max(n, 1),))[:n].copy()
    ref().vec_put(C.byref(v))
    return out


# ------------------------------------ the reference's own CUDA kernels on this GPU --
class _bench_cuda(C.Structure):
    _fields_ = [("bench", _bench), ("warps_per_block", C.c_int)]


def ref_cuda_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libspmv_ref_cuda.so"))


def ref_cuda_bench(M, N, IRP, JA, AS, x, reps=5, wpbs=(4, 8)):
    """Times the reference's UNMODIFIED kernels (src/cuda_csr.cu, src/cuda_hll.cu rebuilt for
    sm_100a by oracle/Makefile) through the reference's own bench_*_cuda_* trampolines
    (src/csr.c:382-415, src/hll.c:226-256): every call uploads, runs ONE launch between CUDA events,
    downloads (src/cuda_csr.cu:210-233).  Must run in a process that has NOT loaded libspmv_b200
    (both export csr_spmv_cuda_*): bench.py calls `python -m oracle.oracle refcuda ...`.
    Returns {name: {"ms_min", "ms_median", "gflops_best", "ok"}}."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libspmv_ref_cuda.so"))
    lib.csr_to_hll.restype = C.c_void_p
    lib.csr_to_hll.argtypes = [C.POINTER(_csr), C.c_bool]
    lib.hll_free.argtypes = [C.c_void_p]
    lib.vec_put.argtypes = [C.POINTER(_vec)]
    A = RefCsr(M, N, IRP, JA, AS)
    xa = aligned_copy(x, np.float64)
    y_ref = csr_spmv(M, IRP, JA, AS, x)
    bound = csr_abs_bound(M, IRP, JA, AS, x)
    out = {}
    nnz = len(JA)

    def run(fname, mat_ptr, label):
        fn = getattr(lib, fname)
        fn.argtypes = [C.c_void_p, _dp, C.POINTER(_bench_cuda)]
        for wpb in wpbs:
            ms, ok = [], True
            for _ in range(reps):
                b = _bench_cuda()
                b.warps_per_block = wpb
                rc = fn(mat_ptr, _cd(xa), C.byref(b))
                if rc:
                    ok = False
                    break
                n = b.bench.data.len
                y = np.ctypeslib.as_array(b.bench.data.data, shape=(max(n, 1),))[:n].copy()
                lib.vec_put(C.byref(b.bench.data))
                ms.append(b.bench.duration_ms)
                ok = ok and check_tolerance(y, y_ref, bound, 1e-12)[0]
            if ms and min(ms) > 0:
                out[f"{label}_wpb{wpb}"] = {"ms_min": min(ms), "ms_median": sorted(ms)[len(ms) // 2],
                                            "gflops_best": 2.0 * nnz / (min(ms) * 1e6), "ok": bool(ok)}

    import ctypes
    a_ptr = ctypes.cast(A.ptr, C.c_void_p)
    run("bench_csr_cuda_halfwarp_row", a_ptr, "csr_k2_halfwarp_row")
    run("bench_csr_cuda_warp_row", a_ptr, "csr_k1_warp_row")
    if nnz <= 100_000_000:     # the reference's HLL upload does 2 cudaMalloc + 3 cudaMemcpy per hack and call
        Hc = lib.csr_to_hll(A.ptr, True)
        if not _is_err(Hc):
            run("bench_hll_cuda_threads_col_major", C.c_void_p(Hc), "hll_k1_threads_col_major")
            run("bench_hll_cuda_warp_block", C.c_void_p(Hc), "hll_k2_warp_block")
            lib.hll_free(Hc)
    return out


if __name__ == "__main__":
    import json
    if len(sys.argv) >= 3 and sys.argv[1] == "refcuda":
        spec = sys.argv[2]
        if spec == "c2":
            M, N, IRP, JA, AS = gen_stencil27(128, 128, 128)
        elif spec == "c1":
            M, N, IRP, JA, AS = gen_poisson2d(1000, 1000)
        elif spec == "tiny":
            M, N, IRP, JA, AS = gen_stencil27(24, 24, 24)
        else:
            raise SystemExit("refcuda: c1 | c2 | tiny")
        x = np.random.default_rng(0).uniform(0, 1, N)
        print(json.dumps(ref_cuda_bench(M, N, IRP, JA, AS, x)))
