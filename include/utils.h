/* utils.h -- shared constants, benchmark result records, small helpers.
 *
 * Mirrors the observable surface of the reference's include/utils.h:
 * constants (:12-18), bench / bench_omp / bench_cuda records (:32-47),
 * log_prog_usage / validation_vec_result / aligned_malloc (:61-77),
 * now() and compute_gflops() (:68-75).  Record layouts are ABI.
 */
#ifndef SPMV_B200_UTILS_H
#define SPMV_B200_UTILS_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include "vector.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ALIGNMENT 64 /* host array alignment, bytes */
#define MAX_NAME 64  /* matrix / bench name buffer */
#define MAX_PATH 256 /* csv path buffer */

#define ARRAY_SIZE(a) (sizeof(a) / sizeof((a)[0]))

/* One timed SpMV: milliseconds, 2*nnz/t GFLOP/s and the produced y. */
typedef struct benchmark_result {
      double duration_ms;
      double gflops;
      vec data;
} bench;

typedef struct benchmark_omp {
      bench bench;
      char name[MAX_NAME]; /* "omp_nnz" | "omp_guided" */
      int num_threads;
} bench_omp;

typedef struct benchmark_cuda {
      bench bench;
      int warps_per_block;
} bench_cuda;

#define LOG_WARN(fmt, ...)                                                     \
      do {                                                                     \
            fprintf(stdout, "[WARN ] %s:%d: " fmt "\n", __FILE__, __LINE__,    \
                    ##__VA_ARGS__);                                            \
      } while (0)

#define LOG_INFO(fmt, ...)                                                     \
      do {                                                                     \
            fprintf(stdout, "[INFO ] %s:%d: " fmt "\n", __FILE__, __LINE__,    \
                    ##__VA_ARGS__);                                            \
      } while (0)

/* Spin up an OpenMP team once so the first timed region is not charged for
 * thread creation (reference: OMP_WARMUP, include/utils.h:20-30). */
void omp_warmup(int num_threads);
#define OMP_WARMUP(nt) omp_warmup(nt)

void log_prog_usage(const char *prog);
void print_result_vector(const vec res);

/* 0 when ||expected - res||_2 <= 0.1 and lengths match, else -1
 * (reference: src/utils.c:39-60; the `-d` gate of the CLI). */
int validation_vec_result(const vec expected, const vec res);

/* posix_memalign(ALIGNMENT) or NULL. */
void *aligned_malloc(size_t size);

/* CPU time in milliseconds (clock()), as the serial paths of the reference. */
static inline double now(void) {
      return (double)clock() * 1e3 / (double)CLOCKS_PER_SEC;
}

/* 2*nnz flop in `duration` ms -> GFLOP/s; non-positive time -> 0. */
static inline double compute_gflops(double duration, int nnz) {
      return duration > 0.0 ? (2.0 * (double)nnz) / (duration * 1e6) : 0.0;
}

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_UTILS_H */
