/* logger.h -- the three append-mode CSV sinks (serial.csv, omp.csv, cuda.csv).
 *
 * Schema and row formats are byte-identical to the reference's
 * (src/logger.c:31-41 headers, :89-153 rows) because scripts/results.py and
 * scripts/plots.py consume them.
 */
#ifndef SPMV_B200_LOGGER_H
#define SPMV_B200_LOGGER_H

#include "csr.h"
#include "hll.h"
#include "utils.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Opens <base_path>/{serial,omp,cuda}.csv (directory must exist); a header
 * line is written only when the file did not exist.  0 / -1. */
int logger_init(const char *base_path);
void logger_close(void);

void log_csr_serial_benchmark(const sparse_csr *A, bench res);
void log_hll_serial_benchmark(const sparse_hll *H, bench res);
void log_csr_omp_benchmark(const sparse_csr *A, bench_omp res);
void log_hll_omp_benchmark(const sparse_hll *H, bench_omp res);
void log_csr_cuda_benchmark(const sparse_csr *A, bench_cuda res, int kernel_id);
void log_hll_cuda_benchmark(const sparse_hll *H, bench_cuda res, int kernel_id);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_LOGGER_H */
