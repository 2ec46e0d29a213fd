/* spmv_b200.h -- resident-matrix ("handle") API of libspmv_b200.
 *
 * NEW SURFACE, no reference counterpart.  The reference launchers
 * (src/cuda_csr.cu:180-233, src/cuda_hll.cu:154-260) upload the matrix, run
 * one launch and free everything on every call.  The entry points in
 * cuda_csr.h / cuda_hll.h keep that calling convention; underneath they use
 * the API below, which keeps a matrix resident in HBM, runs y = A*x on
 * DEVICE pointers on a caller-supplied stream, and times launches properly.
 * It is also what bench.py, the parity tests and the multi-GPU iterated SpMV
 * bind.  Plain C ABI: pointers, sizes, ints -- no CUDA or torch types
 * (a stream is a `void *` holding a cudaStream_t; NULL = default stream).
 *
 * Conventions
 *   - functions returning int: 0 = success, negative errno-style code on
 *     failure; functions returning a pointer: NULL on failure.  In both cases
 *     spmv_b200_last_error() holds a message for the calling thread.
 *   - "d_" parameters are device pointers, everything else is host memory.
 *   - there is NO CPU fallback: with no usable GPU every create/run fails.
 *   - threading: a handle is used by one thread at a time (launch plans are built lazily on
 *     first use of a kernel); different handles may be used concurrently; the reference-style
 *     entry points of cuda_csr.h / cuda_hll.h serialise on an internal lock.
 *   - devices: the intended model is one process per GPU (spmv_b200_set_device once, early).
 *     Handles remember their device; the scratch buffers of the reference-style entry points
 *     and of spmv_b200_flush_l2 belong to the device that was current when first used.
 */
#ifndef SPMV_B200_H
#define SPMV_B200_H

#include <stddef.h>
#include <stdint.h>

#include "csr.h"
#include "hll.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- kernel selectors (same numbering as the CSV `kernel` column) -------- */
enum spmv_b200_csr_kernel {
      SPMV_B200_CSR_THREAD_ROW = 0, /* one thread per row */
      SPMV_B200_CSR_WARP_ROW = 1,   /* one warp per row */
      SPMV_B200_CSR_ADAPTIVE = 2,   /* rows binned by length (headline) */
      SPMV_B200_CSR_BLOCK_ROW = 3,  /* one CTA per row */
      SPMV_B200_CSR_STREAM = 4,     /* cp.async.bulk staged row tiles */
      SPMV_B200_CSR_NUM_KERNELS = 5
};

enum spmv_b200_hll_kernel {
      SPMV_B200_HLL_THREAD_ROW_RM = 0, /* thread per row (row-major input) */
      SPMV_B200_HLL_THREAD_ROW = 1,    /* thread per row, scalar loads */
      SPMV_B200_HLL_WARP_HACK = 2,     /* warp per hack, vector loads (headline) */
      SPMV_B200_HLL_STREAM = 3,        /* cp.async.bulk staged hacks */
      SPMV_B200_HLL_NUM_KERNELS = 4
};

/* ---- library / device ---------------------------------------------------- */
const char *spmv_b200_last_error(void);
const char *spmv_b200_version(void);
int spmv_b200_device_count(void);
int spmv_b200_set_device(int ordinal);

typedef struct {
      char name[128];
      int cc_major, cc_minor;
      int sm_count;
      int l2_bytes_mb;
      int64_t hbm_bytes;
      int max_smem_per_block;
} spmv_b200_devinfo;
int spmv_b200_device_info(spmv_b200_devinfo *out);

/* ---- raw device memory (for C hosts and ctypes callers without torch) ---- */
void *spmv_b200_dmalloc(size_t bytes);
int spmv_b200_dfree(void *d_ptr);
int spmv_b200_h2d(void *d_dst, const void *src, size_t bytes, void *stream);
int spmv_b200_d2h(void *dst, const void *d_src, size_t bytes, void *stream);
int spmv_b200_dmemset(void *d_dst, int byte, size_t bytes, void *stream);
int spmv_b200_stream_sync(void *stream);
/* Pinned host staging memory (e2e path). */
void *spmv_b200_host_alloc(size_t bytes);
int spmv_b200_host_free(void *ptr);
/* Overwrite a scratch buffer larger than L2 so the next launch starts cold. */
int spmv_b200_flush_l2(void *stream);

/* ---- resident CSR --------------------------------------------------------- */
typedef struct spmv_b200_csr spmv_b200_csr;

/* Upload a host CSR (copies; the host arrays may be freed afterwards) and
 * build the launch plan (row bins, stream tiles). */
spmv_b200_csr *spmv_b200_csr_create(const sparse_csr *A);

/* General form used for shards: `irp` holds M+1 row offsets of width
 * `irp_bytes` (4 or 8); column indices are stored as (JA[k] - col_offset) and
 * must land in [0, n_local): the shard multiplies a local slice of x that
 * starts at global column `col_offset`.  `cuts`/`n_cuts` (may be NULL/0) are
 * row indices at which launch tiles must break so that
 * spmv_b200_csr_spmv_rows() can be called on [cuts[i], cuts[i+1]). */
spmv_b200_csr *spmv_b200_csr_create_ex(int64_t M, int64_t n_local, int64_t NZ,
                                       const void *irp, int irp_bytes,
                                       const int *JA, const double *AS,
                                       int64_t col_offset, const int64_t *cuts,
                                       int n_cuts);

/* 27-point stencil on an nx*ny*nz grid (diag 26, others -1, lexicographic
 * numbering, x fastest, sorted columns), rows of planes [z0, z1) generated
 * directly in HBM.  Columns are stored relative to `col_offset`. */
spmv_b200_csr *spmv_b200_csr_gen_stencil27(int nx, int ny, int nz, int z0,
                                           int z1, int64_t col_offset,
                                           int64_t n_local,
                                           const int64_t *cuts, int n_cuts);

int64_t spmv_b200_csr_rows(const spmv_b200_csr *h);
int64_t spmv_b200_csr_cols(const spmv_b200_csr *h);
int64_t spmv_b200_csr_nnz(const spmv_b200_csr *h);
/* Copy the device arrays back (any of the outputs may be NULL).  `irp64`
 * receives M+1 64-bit offsets. */
int spmv_b200_csr_download(const spmv_b200_csr *h, int64_t *irp64, int *JA,
                           double *AS);
/* Rows per adaptive bin: out[0..5] = {empty, sub-warp, warp, block, split,
 * tiles}.  Diagnostic. */
int spmv_b200_csr_plan_info(const spmv_b200_csr *h, int64_t *out, int n_out);

/* y[0..M) = A * x on `stream` (asynchronous).  d_x has n_local entries. */
int spmv_b200_csr_spmv(spmv_b200_csr *h, int kernel, int warps_per_block,
                       const double *d_x, double *d_y, void *stream);
/* Same for rows [row0, row1) only; y is still indexed by local row. */
int spmv_b200_csr_spmv_rows(spmv_b200_csr *h, int kernel, int warps_per_block,
                            int64_t row0, int64_t row1, const double *d_x,
                            double *d_y, void *stream);
/* Fused epilogue for the multi-GPU halo exchange: besides y[row], rows in
 * [push_row0[i], push_row1[i]) are also stored to d_push_dst[i][row -
 * push_row0[i]] (a peer GPU's halo buffer mapped through CUDA IPC).  Up to 2
 * push ranges. */
int spmv_b200_csr_spmv_rows_push(spmv_b200_csr *h, int kernel,
                                 int warps_per_block, int64_t row0,
                                 int64_t row1, const double *d_x, double *d_y,
                                 int n_push, const int64_t *push_row0,
                                 const int64_t *push_row1,
                                 double *const *d_push_dst, void *stream);
/* Kernel launches one spmv call issues for this matrix/kernel. */
int spmv_b200_csr_launches(const spmv_b200_csr *h, int kernel);
/* Run `warmup` untimed + `reps` timed launches, each bracketed by CUDA events
 * on `stream`; ms_out[reps] receives per-launch milliseconds.  flush_l2 != 0
 * evicts L2 before every timed launch (outside the timed interval). */
int spmv_b200_csr_time(spmv_b200_csr *h, int kernel, int warps_per_block,
                       const double *d_x, double *d_y, int warmup, int reps,
                       int flush_l2, double *ms_out, void *stream);
void spmv_b200_csr_destroy(spmv_b200_csr *h);

/* ---- resident HLL --------------------------------------------------------- */
typedef struct spmv_b200_hll spmv_b200_hll;

/* Flatten + upload a host HLL of either layout. */
spmv_b200_hll *spmv_b200_hll_create(const sparse_hll *H, int is_col_major);
/* Build the device HLL straight from a resident CSR, on the GPU. */
spmv_b200_hll *spmv_b200_hll_from_csr(const spmv_b200_csr *A);

int64_t spmv_b200_hll_rows(const spmv_b200_hll *h);
int64_t spmv_b200_hll_cols(const spmv_b200_hll *h);
int64_t spmv_b200_hll_nnz(const spmv_b200_hll *h);
int64_t spmv_b200_hll_num_hacks(const spmv_b200_hll *h);
int64_t spmv_b200_hll_slots(const spmv_b200_hll *h); /* padded entries */
/* Device layout back to the host for bit-compare: hoff[num_hacks+1] slot
 * offsets, JA/AS[slots] column-major per hack with stride 32. */
int spmv_b200_hll_download(const spmv_b200_hll *h, int64_t *hoff, int *JA,
                           double *AS);

int spmv_b200_hll_spmv(spmv_b200_hll *h, int kernel, int warps_per_block,
                       const double *d_x, double *d_y, void *stream);
int spmv_b200_hll_launches(const spmv_b200_hll *h, int kernel);
int spmv_b200_hll_time(spmv_b200_hll *h, int kernel, int warps_per_block,
                       const double *d_x, double *d_y, int warmup, int reps,
                       int flush_l2, double *ms_out, void *stream);
void spmv_b200_hll_destroy(spmv_b200_hll *h);

/* ---- cache used by the reference-style entry points ---------------------- */
/* Drop every device copy kept by csr_spmv_cuda_* / hll_spmv_cuda_*. */
void spmv_b200_release_all(void);
/* Timing policy of the reference-style entry points (defaults 1 / 3; also
 * settable with SPMV_B200_WARMUP / SPMV_B200_REPS).  The returned duration is
 * the median of the timed repetitions.  reps = 0 asks for no separately timed
 * launches: the pipelined entry (banded matrix, page-locked x and y: x is
 * uploaded in column order while row chunks compute and y chunks travel back)
 * then returns the span of its single pass; other paths time one launch. */
void spmv_b200_set_timing(int warmup, int reps);
/* Counters since load: kernel launches issued, bytes copied H2D / D2H. */
void spmv_b200_counters(int64_t *launches, int64_t *h2d_bytes,
                        int64_t *d2h_bytes);

/* Experiment knobs used by bin/kbench sweeps ("stream_hints", "csr_stream_cfg",
 * "hll_vec", "hll_stream_cfg", "regular_lpr", "adaptive_direct", "force_wide", "pipeline").  0 or -EINVAL.  Knobs that
 * change planning ("regular_lpr") must be set before a handle is created. */
int spmv_b200_set_knob(const char *key, int value);

/* ---- inter-process peer memory (one process per GPU) --------------------- */
#define SPMV_B200_IPC_HANDLE_BYTES 64
int spmv_b200_ipc_export(void *d_ptr, unsigned char *handle64);
int spmv_b200_ipc_open(const unsigned char *handle64, void **d_ptr_out);
int spmv_b200_ipc_close(void *d_ptr);
int spmv_b200_enable_peer(int peer_device);

/* Step ordering between ranks without a host round trip (multi-GPU halo push):
 * spmv_b200_signal_peers  epoch := epoch + 1 (device word), then store it into
 *                         each of the n peer slots (peer HBM mapped through IPC);
 * spmv_b200_wait_peers    spin (at most max_spins polls per slot, then *d_error
 *                         is set to 1 + slot index and the kernel returns) until
 *                         each of my n slots holds a value >= my epoch.
 * Both are single-thread kernels on `stream`, capturable in a CUDA graph. */
int spmv_b200_signal_peers(void *d_epoch, int n, void *const *d_peer_slots,
                           void *stream);
int spmv_b200_wait_peers(const void *d_epoch, int n, void *const *d_my_slots,
                         uint64_t max_spins, int *d_error, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_H */
