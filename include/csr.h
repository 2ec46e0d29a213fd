/* csr.h -- host CSR container, Matrix Market loader and CSR benchmarks.
 *
 * Drop-in for the reference's include/csr.h: struct layout (:7-13,
 * sizeof == 104 on x86-64), init_csr (:15-24) and every prototype (:29-49)
 * keep their meaning.  Indices are 0-based int32, values FP64, IRP has M+1
 * entries, within-row order is the order of appearance in the .mtx file.
 */
#ifndef SPMV_B200_CSR_H
#define SPMV_B200_CSR_H

#include <stdio.h>

#include "utils.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sparse_matrix_csr {
      char name[MAX_NAME]; /* basename of the .mtx without extension */
      int M, N, NZ;        /* rows, cols, stored entries */
      int *IRP;            /* [M+1] row pointers */
      int *JA;             /* [NZ] column indices */
      double *AS;          /* [NZ] values */
} sparse_csr;

/* Wrap caller-owned arrays in a sparse_csr (no copy). */
static inline void init_csr(sparse_csr *A, const char *name, int M, int N,
                            int NZ, int *IRP, int *JA, double *AS) {
      snprintf(A->name, sizeof A->name, "%s", name);
      A->M = M, A->N = N, A->NZ = NZ;
      A->IRP = IRP, A->JA = JA, A->AS = AS;
}

/* Matrix Market coordinate {real,pattern} x {general,symmetric,...} -> CSR.
 * Failure is an ERR_PTR(-errno): -EINVAL (unsupported banner / size line),
 * -ERANGE (index outside the declared shape), -EIO (short file), -ENOMEM,
 * or -errno from fopen (reference: src/csr.c:31-171). */
sparse_csr *io_load_csr(const char *path);

/* Frees the three arrays and the struct; NULL is ignored. */
void csr_free(sparse_csr *A);

/* CPU paths (serial / OpenMP).  They exist so the CLI keeps writing
 * serial.csv and omp.csv; they are never used by the GPU path. */
int bench_csr_serial(const sparse_csr *A, const double *x, bench *out);
int bench_csr_omp_guided(const sparse_csr *A, const double *x, bench_omp *out);
int bench_csr_omp_nnz_balancing(const sparse_csr *A, const double *x,
                                bench_omp *out);

/* GPU paths: set out->warps_per_block, call; out->bench is filled with the
 * kernel time (ms), GFLOP/s and a freshly allocated y.  Each forwards to the
 * matching csr_spmv_cuda_* entry of libspmv_b200 (cuda_csr.h). */
int bench_csr_cuda_thread_row(const sparse_csr *A, const double *x,
                              bench_cuda *out);
int bench_csr_cuda_warp_row(const sparse_csr *A, const double *x,
                            bench_cuda *out);
int bench_csr_cuda_halfwarp_row(const sparse_csr *A, const double *x,
                                bench_cuda *out);
int bench_csr_cuda_block_row(const sparse_csr *A, const double *x,
                             bench_cuda *out);
int bench_csr_cuda_halfwarp_row_text(const sparse_csr *A, const double *x,
                                     bench_cuda *out);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_CSR_H */
