/* utils.h -- constants, benchmark records and helpers shared by the host layer.
 *
 * The three records are ABI (reference include/utils.h:32-47): they are filled by the
 * bench_* functions and passed BY VALUE to the logger.
 */
#ifndef SPMV_B200_UTILS_H
#define SPMV_B200_UTILS_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include "spmv_errptr.h"
#include "vector.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ALIGNMENT = 64, /* bytes; every array the loader / packer hands out     */
    MAX_NAME = 64,  /* matrix and bench name buffers                        */
    MAX_PATH = 256  /* csv path buffer                                      */
};
#define ARRAY_SIZE(a) (sizeof(a) / sizeof((a)[0]))

/* result of one timed y = A x */
typedef struct benchmark_result {
    double duration_ms; /* kernel / loop time                                */
    double gflops;      /* 2 nnz / time                                      */
    vec data;           /* the y that was produced; the caller releases it   */
} bench;

typedef struct benchmark_omp {
    bench bench;
    char name[MAX_NAME]; /* "omp_nnz" or "omp_guided" (csv column `bench`)   */
    int num_threads;
} bench_omp;

typedef struct benchmark_cuda {
    bench bench;
    int warps_per_block; /* in: CTA size knob; echoed into cuda.csv          */
} bench_cuda;

/* start an OpenMP team once so the first timed region does not pay for thread creation */
void omp_warmup(int num_threads);
#define OMP_WARMUP(n) omp_warmup(n)

void log_prog_usage(const char *prog);
void print_result_vector(const vec y);

/* the `-d` gate: 0 when the lengths agree and ||expected - got||_2 <= 0.1, otherwise -1
 * (reference src/utils.c:39-60) */
int validation_vec_result(const vec expected, const vec got);

/* posix_memalign(ALIGNMENT); NULL on failure */
void *aligned_malloc(size_t bytes);

/* process CPU time in milliseconds -- the clock of the serial paths */
static inline double now(void) { return 1e3 * (double)clock() / (double)CLOCKS_PER_SEC; }

/* GFLOP/s of 2*nnz flop done in `ms` milliseconds; a failed run (ms <= 0) scores 0 */
static inline double compute_gflops(double ms, int nnz) {
    if (!(ms > 0.0))
        return 0.0;
    return 2.0 * (double)nnz / (ms * 1e6);
}

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_UTILS_H */
