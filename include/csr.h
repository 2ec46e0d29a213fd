/* csr.h -- host CSR matrix, its Matrix Market loader, and the CSR benchmark entry points.
 *
 * Binary-compatible with the reference (include/csr.h:7-13: 104 bytes on x86-64; prototypes
 * :29-49).  Conventions: 0-based int32 indices, FP64 values, IRP has M+1 entries, entries of a
 * row keep the order in which the .mtx file listed them (duplicates and explicit zeros kept).
 */
#ifndef SPMV_B200_CSR_H
#define SPMV_B200_CSR_H

#include <stdio.h>

#include "utils.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sparse_matrix_csr {
    char name[MAX_NAME]; /* file name without directory and ".mtx"        */
    int M, N, NZ;        /* rows, columns, stored entries                  */
    int *IRP;            /* row offsets, M + 1 of them                     */
    int *JA;             /* column of every entry                          */
    double *AS;          /* value of every entry                           */
} sparse_csr;

/* Describe arrays the caller already owns; nothing is copied or allocated. */
static inline void init_csr(sparse_csr *dst, const char *name, int rows, int cols, int nnz,
                            int *irp, int *ja, double *as) {
    snprintf(dst->name, sizeof dst->name, "%s", name);
    dst->M = rows;
    dst->N = cols;
    dst->NZ = nnz;
    dst->IRP = irp;
    dst->JA = ja;
    dst->AS = as;
}

/* "matrix coordinate {real|pattern} {general|symmetric|...}" file -> CSR.
 * On failure the result satisfies IS_ERR() and PTR_ERR() is
 *   -EINVAL  banner / size line not acceptable (integer, complex, array, ...)
 *   -ERANGE  an index outside the declared shape
 *   -EIO     fewer entries than declared, or a token that is not a number
 *   -ENOMEM, or -errno of fopen()
 * exactly as the reference loader decides (src/csr.c:31-171). */
sparse_csr *io_load_csr(const char *mtx_path);

/* releases the three arrays and the struct (NULL is fine) */
void csr_free(sparse_csr *matrix);

/* ---- benchmarks: each allocates y, runs one variant, fills duration / GFLOP/s ------------
 * CPU variants (kept so the driver still writes serial.csv / omp.csv; never a fallback for
 * the GPU path): */
int bench_csr_serial(const sparse_csr *matrix, const double *x, bench *result);
int bench_csr_omp_guided(const sparse_csr *matrix, const double *x, bench_omp *result);
int bench_csr_omp_nnz_balancing(const sparse_csr *matrix, const double *x, bench_omp *result);

/* GPU variants: result->warps_per_block is an input.  They forward to the csr_spmv_cuda_*
 * symbol of the same suffix in libspmv_b200 (cuda_csr.h lists the kernel behind each). */
#define SPMV_CSR_CUDA_VARIANTS(X) \
    X(thread_row) X(warp_row) X(halfwarp_row) X(block_row) X(halfwarp_row_text)
#define SPMV_DECLARE(suffix) \
    int bench_csr_cuda_##suffix(const sparse_csr *matrix, const double *x, bench_cuda *result);
SPMV_CSR_CUDA_VARIANTS(SPMV_DECLARE)
#undef SPMV_DECLARE

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_CSR_H */
