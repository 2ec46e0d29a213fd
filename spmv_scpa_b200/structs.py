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


def is_err_ptr(addr):
    """IS_ERR() of include/err.h on an integer address."""
    return addr is not None and addr > (1 << 64) - 4096


def ptr_err(addr):
    return int(addr) - (1 << 64)
