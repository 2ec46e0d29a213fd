/* cuda_csr.h -- GPU CSR SpMV entry points exported by libspmv_b200.
 *
 * These seven symbols are exactly what the reference's host code binds
 * (reference include/cuda_csr.h:10-25; callers src/csr.c:382-415).  Contract
 * kept from the reference launchers (src/cuda_csr.cu:210-371):
 *   - A, x (length A->N) and y (length A->M) are HOST pointers owned by the
 *     caller; nothing is retained beyond an internal device-side cache;
 *   - y is fully written when the call returns (synchronous);
 *   - the return value is the elapsed time of the SpMV kernel(s) only, in
 *     milliseconds (transfers excluded);
 *   - set_csr_warps_per_block() selects the CTA size (32*wppb threads) used by
 *     the next call from the same thread.
 * Difference: every CUDA call is checked; on failure a line is printed to
 * stderr and a value <= 0 is returned (=> compute_gflops() == 0).
 *
 * Which sm_100a kernel sits behind each entry (the CSV `kernel` id keeps its
 * position, reference src/main.c:259-263):
 *   0 csr_spmv_cuda_thread_row        one thread per row
 *   1 csr_spmv_cuda_warp_row          one warp per row, shuffle reduction
 *   2 csr_spmv_cuda_halfwarp_row      ADAPTIVE by row-length profile and gather locality:
 *                                     TMA-staged tiles (regular rows; per-warp rings for
 *                                     rows of at most 8 entries), sorted slices with
 *                                     virtual rows (ragged rows), column panels (x > L2,
 *                                     scattered columns); sub-warp / warp / block-per-row
 *                                     bins on row ranges of cut shards
 *   3 csr_spmv_cuda_block_row         one CTA per row
 *   4 csr_spmv_cuda_halfwarp_row_text STREAM: value/index streams staged in
 *                                     shared memory by cp.async.bulk (TMA); the adaptive
 *                                     choice where staging cannot win (ragged rows,
 *                                     scattered columns with x > L2)
 */
#ifndef SPMV_B200_CUDA_CSR_H
#define SPMV_B200_CUDA_CSR_H

#include "csr.h"

#ifdef __cplusplus
extern "C" {
#endif

void set_csr_warps_per_block(int wppb);

/* one prototype per variant of SPMV_CSR_CUDA_VARIANTS (see csr.h):
 *   double csr_spmv_cuda_<variant>(const sparse_csr *matrix, const double *x_host,
 *                                    double *y_host, void *unused);                  */
#define SPMV_DECLARE(suffix)                                                          \
    double csr_spmv_cuda_##suffix(const sparse_csr *matrix, const double *x_host,        \
                                    double *y_host, void *unused);
SPMV_CSR_CUDA_VARIANTS(SPMV_DECLARE)
#undef SPMV_DECLARE

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_CUDA_CSR_H */
