/* hll.h -- HLL ("hacked ELLPACK") host matrix: the rows are cut into hacks of 32, every hack
 * is an ELLPACK block padded to its own widest row.
 *
 * Binary-compatible with the reference (include/hll.h:13-18 ellpack_block = 32 bytes,
 * :31-37 sparse_hll = 96 bytes; prototypes :54-70).  Inside a hack of `M` rows and width
 * `max_NZ`, entry j of row i sits at
 *      i * max_NZ + j     row-major
 *      j * M + i          column-major   (stride = rows of THIS hack: the last hack of a
 *                                         matrix whose row count is not a multiple of 32 is
 *                                         narrower)
 * and every unused slot holds JA = -1, AS = 0.0.
 */
#ifndef SPMV_B200_HLL_H
#define SPMV_B200_HLL_H

#include <stdbool.h>
#include <stdlib.h>

#include "csr.h"
#include "utils.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HACK_SIZE 32

typedef struct {
    int M, N, NZ; /* rows of this hack, columns of the matrix, real entries in the hack */
    int max_NZ;   /* slots per row                                                        */
    int *JA;      /* M * max_NZ column indices                                            */
    double *AS;   /* M * max_NZ values                                                    */
} ellpack_block;

typedef struct {
    char name[MAX_NAME];
    int M, N, NZ;
    int hack_size;  /* HACK_SIZE                       */
    int num_blocks; /* (M + HACK_SIZE - 1) / HACK_SIZE */
    ellpack_block *blocks;
} sparse_hll;

static inline void init_ellpack_block(ellpack_block *dst, int rows, int cols, int entries,
                                      int width) {
    dst->M = rows;
    dst->N = cols;
    dst->NZ = entries;
    dst->max_NZ = width;
    dst->JA = NULL;
    dst->AS = NULL;
}

static inline void init_hll(sparse_hll *dst, const char *name, int rows, int cols, int nnz,
                            int hacks) {
    snprintf(dst->name, sizeof dst->name, "%s", name);
    dst->M = rows;
    dst->N = cols;
    dst->NZ = nnz;
    dst->hack_size = HACK_SIZE;
    dst->num_blocks = hacks;
    dst->blocks = NULL;
}

/* Pack a CSR matrix; every field of every hack equals what the reference packer produces
 * (src/hll.c:19-95).  IS_ERR() result with -ENOMEM when memory runs out. */
sparse_hll *csr_to_hll(const sparse_csr *matrix, bool column_major);
void hll_free(sparse_hll *matrix);

/* CPU variants (row-major hacks; padding skipped by testing JA == -1) */
int bench_hll_serial(const sparse_hll *matrix, const double *x, bench *result);
int bench_hll_omp(const sparse_hll *matrix, const double *x, bench_omp *result);

/* GPU variants.  The entry point implies the host layout it is given (reference
 * src/main.c:324-325): row-major for threads_row_major and halfwarp_row, column-major for
 * threads_col_major and warp_block. */
#define SPMV_HLL_CUDA_VARIANTS(X) \
    X(threads_row_major) X(threads_col_major) X(warp_block) X(halfwarp_row)
#define SPMV_DECLARE(suffix) \
    int bench_hll_cuda_##suffix(const sparse_hll *matrix, const double *x, bench_cuda *result);
SPMV_HLL_CUDA_VARIANTS(SPMV_DECLARE)
#undef SPMV_DECLARE

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_HLL_H */
