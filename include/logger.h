/* logger.h -- serial.csv, omp.csv and cuda.csv, opened in append mode under one directory.
 *
 * The files are what the reference's scripts read (scripts/results.py, scripts/plots.py), so
 * headers and row formats are byte-identical to src/logger.c:31-41 / :89-153; the result
 * records are taken by value like there.
 */
#ifndef SPMV_B200_LOGGER_H
#define SPMV_B200_LOGGER_H

#include "csr.h"
#include "hll.h"
#include "utils.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The directory must exist.  A header line is written only into files that did not exist.
 * 0 when all three files are open, -1 otherwise. */
int logger_init(const char *directory);
void logger_close(void);

/* One row per benchmark result.
 *   X(function, matrix type, result record, extra trailing parameters)
 * serial.csv : log_{csr,hll}_serial_benchmark(matrix, bench)
 * omp.csv    : log_{csr,hll}_omp_benchmark(matrix, bench_omp)
 * cuda.csv   : log_{csr,hll}_cuda_benchmark(matrix, bench_cuda, kernel_id)            */
#define SPMV_LOG_FUNCTIONS(X)                                              \
    X(log_csr_serial_benchmark, sparse_csr, bench, )                       \
    X(log_hll_serial_benchmark, sparse_hll, bench, )                       \
    X(log_csr_omp_benchmark, sparse_csr, bench_omp, )                      \
    X(log_hll_omp_benchmark, sparse_hll, bench_omp, )                      \
    X(log_csr_cuda_benchmark, sparse_csr, bench_cuda, SPMV_LOG_KERNEL_ID)  \
    X(log_hll_cuda_benchmark, sparse_hll, bench_cuda, SPMV_LOG_KERNEL_ID)
#define SPMV_LOG_KERNEL_ID , int kernel_id
#define SPMV_DECLARE(fn, matrix_t, record_t, extra) void fn(const matrix_t *matrix, record_t result extra);
SPMV_LOG_FUNCTIONS(SPMV_DECLARE)
#undef SPMV_DECLARE

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_LOGGER_H */
