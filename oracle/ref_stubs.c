/* ref_stubs.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * The reference's host sources (src/csr.c, src/hll.c) reference the eleven
 * GPU boundary symbols of include/cuda_csr.h:10-25 and include/cuda_hll.h:10-22.
 * To load the reference's CPU half as a shared object (oracle/_ref/) without
 * its CUDA half, those symbols are satisfied here by stubs that compute
 * nothing and report a negative duration.
 */
struct sparse_matrix_csr;
struct sparse_hll_opaque;

void set_csr_warps_per_block(int w) { (void)w; }
void set_hll_warps_per_block(int w) { (void)w; }

#define STUB(name)                                                             \
      double name(const void *A, const double *x, double *y, void *u) {        \
            (void)A, (void)x, (void)y, (void)u;                                \
            return -1.0;                                                       \
      }

STUB(csr_spmv_cuda_thread_row)
STUB(csr_spmv_cuda_warp_row)
STUB(csr_spmv_cuda_halfwarp_row)
STUB(csr_spmv_cuda_block_row)
STUB(csr_spmv_cuda_halfwarp_row_text)
STUB(hll_spmv_cuda_threads_row_major)
STUB(hll_spmv_cuda_threads_col_major)
STUB(hll_spmv_cuda_warp_block)
STUB(hll_spmv_cuda_halfwarp_row)
