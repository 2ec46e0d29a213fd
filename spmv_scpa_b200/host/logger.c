/* logger.c -- append-mode CSV writers for serial / OpenMP / CUDA results.
 *
 * Headers and row formats are byte-identical to reference src/logger.c:31-41
 * and :89-153 (consumed by scripts/results.py / scripts/plots.py): CSR rows
 * leave num_blocks empty, floats are "%f".
 */
#include <stdio.h>
#include <sys/stat.h>

#include "err.h"
#include "logger.h"

enum sink { SINK_SERIAL, SINK_OMP, SINK_CUDA, SINK_COUNT };

static const struct {
      const char *file;
      const char *header;
} k_sinks[SINK_COUNT] = {
    [SINK_SERIAL] = {"serial.csv",
                     "matrix,format,rows,cols,nnz,num_blocks,duration_ms,gflops"},
    [SINK_OMP] = {"omp.csv", "matrix,format,bench,rows,cols,nnz,num_blocks,"
                             "num_threads,duration_ms,gflops"},
    [SINK_CUDA] = {"cuda.csv", "matrix,format,kernel,warps_per_block,rows,cols,"
                               "nnz,num_blocks,duration_ms,gflops"},
};

static FILE *g_out[SINK_COUNT];

static FILE *open_sink(const char *dir, enum sink s) {
      char path[MAX_PATH];
      snprintf(path, sizeof path, "%s/%s", dir, k_sinks[s].file);

      struct stat st;
      const int is_new = stat(path, &st) != 0;

      FILE *f = fopen(path, "a");
      if (f && is_new) {
            fprintf(f, "%s\n", k_sinks[s].header);
            fflush(f);
      }
      return f;
}

int logger_init(const char *base_path) {
      int ok = 1;
      for (int s = 0; s < SINK_COUNT; ++s) {
            g_out[s] = open_sink(base_path, (enum sink)s);
            ok &= g_out[s] != NULL;
      }
      return ok ? 0 : -1;
}

void logger_close(void) {
      for (int s = 0; s < SINK_COUNT; ++s) {
            if (g_out[s])
                  fclose(g_out[s]);
            g_out[s] = NULL;
      }
}

static FILE *sink_or_complain(enum sink s, const char *what) {
      if (!g_out[s])
            LOG_ERR("%s log not initialized", what);
      return g_out[s];
}

void log_csr_serial_benchmark(const sparse_csr *A, bench res) {
      FILE *f = sink_or_complain(SINK_SERIAL, "Serial");
      if (!f)
            return;
      fprintf(f, "%s,CSR,%d,%d,%d,,%f,%f\n", A->name, A->M, A->N, A->NZ,
              res.duration_ms, res.gflops);
      fflush(f);
}

void log_hll_serial_benchmark(const sparse_hll *H, bench res) {
      FILE *f = sink_or_complain(SINK_SERIAL, "Serial");
      if (!f)
            return;
      fprintf(f, "%s,HLL,%d,%d,%d,%d,%f,%f\n", H->name, H->M, H->N, H->NZ,
              H->num_blocks, res.duration_ms, res.gflops);
      fflush(f);
}

void log_csr_omp_benchmark(const sparse_csr *A, bench_omp res) {
      FILE *f = sink_or_complain(SINK_OMP, "OMP");
      if (!f)
            return;
      fprintf(f, "%s,CSR,%s,%d,%d,%d,,%d,%f,%f\n", A->name, res.name, A->M,
              A->N, A->NZ, res.num_threads, res.bench.duration_ms,
              res.bench.gflops);
      fflush(f);
}

void log_hll_omp_benchmark(const sparse_hll *H, bench_omp res) {
      FILE *f = sink_or_complain(SINK_OMP, "OMP");
      if (!f)
            return;
      fprintf(f, "%s,HLL,%s,%d,%d,%d,%d,%d,%f,%f\n", H->name, res.name, H->M,
              H->N, H->NZ, H->num_blocks, res.num_threads,
              res.bench.duration_ms, res.bench.gflops);
      fflush(f);
}

void log_csr_cuda_benchmark(const sparse_csr *A, bench_cuda res,
                            int kernel_id) {
      FILE *f = sink_or_complain(SINK_CUDA, "CUDA");
      if (!f)
            return;
      fprintf(f, "%s,CSR,%d,%d,%d,%d,%d,,%f,%f\n", A->name, kernel_id,
              res.warps_per_block, A->M, A->N, A->NZ, res.bench.duration_ms,
              res.bench.gflops);
      fflush(f);
}

void log_hll_cuda_benchmark(const sparse_hll *H, bench_cuda res,
                            int kernel_id) {
      FILE *f = sink_or_complain(SINK_CUDA, "CUDA");
      if (!f)
            return;
      fprintf(f, "%s,HLL,%d,%d,%d,%d,%d,%d,%f,%f\n", H->name, kernel_id,
              res.warps_per_block, H->M, H->N, H->NZ, H->num_blocks,
              res.bench.duration_ms, res.bench.gflops);
      fflush(f);
}
