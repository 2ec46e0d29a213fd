/* hll.h -- hacked-ELLPACK (HLL) host container and packer, hack = 32 rows.
 *
 * Drop-in for the reference's include/hll.h: ellpack_block (:13-18,
 * sizeof == 32), sparse_hll (:31-37, sizeof == 96), init helpers (:20-28,
 * :39-48) and prototypes (:54-70).  A hack stores `M` consecutive rows padded
 * to the widest one (`max_NZ`); padding is JA = -1 / AS = 0.0; the layout is
 * row-major (slot = i*max_NZ + j) or column-major (slot = j*M + i, stride =
 * rows of THIS hack, so the last hack may use a stride < 32).
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
      int M, N, NZ; /* rows in this hack, matrix cols, real entries */
      int max_NZ;   /* padded width */
      int *JA;      /* [M*max_NZ] */
      double *AS;   /* [M*max_NZ] */
} ellpack_block;

static inline void init_ellpack_block(ellpack_block *blk, int M, int N, int NZ,
                                      int max_NZ) {
      blk->M = M, blk->N = N, blk->NZ = NZ;
      blk->max_NZ = max_NZ;
      blk->JA = NULL;
      blk->AS = NULL;
}

typedef struct {
      char name[MAX_NAME];
      int M, N, NZ;
      int hack_size;  /* always HACK_SIZE */
      int num_blocks; /* ceil(M / HACK_SIZE) */
      ellpack_block *blocks;
} sparse_hll;

static inline void init_hll(sparse_hll *H, const char *name, int M, int N,
                            int NZ, int num_blocks) {
      snprintf(H->name, sizeof H->name, "%s", name);
      H->M = M, H->N = N, H->NZ = NZ;
      H->hack_size = HACK_SIZE;
      H->num_blocks = num_blocks;
      H->blocks = NULL;
}

/* CSR -> HLL, bit-exact with the reference packer (src/hll.c:19-95).
 * Returns ERR_PTR(-ENOMEM) on allocation failure. */
sparse_hll *csr_to_hll(const sparse_csr *A, bool is_col_major);

void hll_free(sparse_hll *H);

int bench_hll_serial(const sparse_hll *H, const double *x, bench *out);
int bench_hll_omp(const sparse_hll *H, const double *x, bench_omp *out);

/* GPU paths; the layout of H is implied by the entry point: row-major for
 * _threads_row_major and _halfwarp_row, column-major for the other two
 * (reference: src/main.c:324-325). */
int bench_hll_cuda_threads_row_major(const sparse_hll *H, const double *x,
                                     bench_cuda *out);
int bench_hll_cuda_threads_col_major(const sparse_hll *H, const double *x,
                                     bench_cuda *out);
int bench_hll_cuda_warp_block(const sparse_hll *H, const double *x,
                              bench_cuda *out);
int bench_hll_cuda_halfwarp_row(const sparse_hll *H, const double *x,
                                bench_cuda *out);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_HLL_H */
