"""ctypes mirrors of the C structs that cross the library boundary.

Layouts follow include/csr.h, include/hll.h, include/utils.h, include/vector.h
(ABI-identical to the reference's include/csr.h:7-13, include/hll.h:13-37,
include/utils.h:32-47, include/vector.h:6-9; sizeof checked in
tests/test_abi.py: sparse_csr 104, ellpack_block 32, sparse_hll 96).
"""
import ctypes as C

MAX_NAME = 64
HACK_SIZE = 32


class sparse_csr(C.Structure):
    _fields_ = [
        ("name", C.c_char * MAX_NAME),
        ("M", C.c_int),
        ("N", C.c_int),
        ("NZ", C.c_int),
        ("IRP", C.POINTER(C.c_int)),
        ("JA", C.POINTER(C.c_int)),
        ("AS", C.POINTER(C.c_double)),
    ]


class ellpack_block(C.Structure):
    _fields_ = [
        ("M", C.c_int),
        ("N", C.c_int),
        ("NZ", C.c_int),
        ("max_NZ", C.c_int),
        ("JA", C.POINTER(C.c_int)),
        ("AS", C.POINTER(C.c_double)),
    ]


class sparse_hll(C.Structure):
    _fields_ = [
        ("name", C.c_char * MAX_NAME),
        ("M", C.c_int),
        ("N", C.c_int),
        ("NZ", C.c_int),
        ("hack_size", C.c_int),
        ("num_blocks", C.c_int),
        ("blocks", C.POINTER(ellpack_block)),
    ]


class vec(C.Structure):
    _fields_ = [("len", C.c_size_t), ("data", C.POINTER(C.c_double))]


class bench(C.Structure):
    _fields_ = [("duration_ms", C.c_double), ("gflops", C.c_double), ("data", vec)]


class bench_omp(C.Structure):
    _fields_ = [("bench", bench), ("name", C.c_char * MAX_NAME), ("num_threads", C.c_int)]


class bench_cuda(C.Structure):
    _fields_ = [("bench", bench), ("warps_per_block", C.c_int)]


class cuda_timer(C.Structure):
    _fields_ = [("start", C.c_void_p), ("stop", C.c_void_p)]


class devinfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 128),
        ("cc_major", C.c_int),
        ("cc_minor", C.c_int),
        ("sm_count", C.c_int),
        ("l2_bytes_mb", C.c_int),
        ("hbm_bytes", C.c_int64),
        ("max_smem_per_block", C.c_int),
    ]


MAX_RANKS = 16
DIST_BLOB_BYTES = 512
DIST_AUTO, DIST_PUSH, DIST_NCCL = 0, 1, 2


class shard_desc(C.Structure):
    """spmv_b200_shard_desc (include/spmv_b200.h)."""
    _fields_ = [(n, C.c_int64) for n in ("r0", "r1", "c0", "c1", "read_lo", "read_hi")]


class xfer(C.Structure):
    _fields_ = [("peer", C.c_int), ("g0", C.c_int64), ("g1", C.c_int64)]


class dist_plan(C.Structure):
    """spmv_b200_dist_plan (include/spmv_b200.h)."""
    _fields_ = [
        ("rank", C.c_int), ("world", C.c_int), ("mode", C.c_int), ("all_gather", C.c_int),
        ("covered", C.c_int), ("n_send", C.c_int), ("n_recv", C.c_int), ("n_cuts", C.c_int),
        ("boundary_lo", C.c_int64), ("boundary_hi", C.c_int64),
        ("cuts", C.c_int64 * 2), ("halo_bytes", C.c_int64),
        ("send", xfer * MAX_RANKS), ("recv", xfer * MAX_RANKS),
    ]


def is_err_ptr(addr):
    """IS_ERR() of include/err.h on an integer address."""
    return addr is not None and addr > (1 << 64) - 4096


def ptr_err(addr):
    return int(addr) - (1 << 64)
