/* main.c -- the `spmv -m <matrix.mtx> -o <out-dir> [-d] [-h]` driver.
 *
 * Same command line, same benchmark schedule and same CSV output as the
 * reference driver (src/main.c:28-109 option handling, :361-379 schedule):
 *   serial CSR, serial HLL                           -> serial.csv
 *   OpenMP CSR nnz-balanced, CSR guided, HLL guided,
 *     each at 2,4,8,16,32,40 threads                 -> omp.csv
 *   GPU CSR kernels 0..4 x warps/block {2,4,8}
 *   GPU HLL kernels 0..3 x warps/block {2,4,8}       -> cuda.csv
 * With -d every variant is compared with the serial CSR result
 * (validation_vec_result).  The GPU rows come from libspmv_b200.
 *
 * Deliberate differences from the reference:
 *   - a failed load is detected with IS_ERR (the reference tests `!A` on an
 *     ERR_PTR and crashes, src/main.c:78-85);
 *   - OpenMP team sizes larger than the machine are run oversubscribed
 *     instead of aborting on an assert (src/csr.c:320, src/hll.c:184);
 *   - SPMV_B200_SKIP_CPU=1 in the environment skips the serial/OpenMP rows
 *     (useful for the multi-GB synthetic inputs; -d then has no effect);
 *   - SPMV_B200_GPUS=N (environment, so scripts/results.py keeps working
 *     unmodified) adds the row-partitioned iterated SpMV x_{k+1} = A x_k on N
 *     GPUs after the reference schedule (square matrices; SPMV_B200_STEPS
 *     steps, default 10) and appends one row to <out-dir>/b200_dist.csv --
 *     a separate file, the three reference CSVs keep their schema.  With -d
 *     the first step is validated against the serial CSR result.
 */
#include <errno.h>
#include <getopt.h>
#include <libgen.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csr.h"
#include "err.h"
#include "hll.h"
#include "logger.h"
#include "spmv_b200.h"
#include "utils.h"

static const int k_omp_threads[] = {2, 4, 8, 16, 32, 40};
static const int k_warps_per_block[] = {2, 4, 8};

struct app {
      bool debug;
      bool logger_open;
      sparse_csr *A;
      sparse_hll *H_rm; /* row-major hacks */
      sparse_hll *H_cm; /* column-major hacks */
      vec x;
      vec expected; /* serial CSR result, kept only with -d */
};

static struct app g;

static void teardown(void) {
      vec_put(&g.expected);
      if (!IS_ERR_OR_NULL(g.H_cm))
            hll_free(g.H_cm);
      if (!IS_ERR_OR_NULL(g.H_rm))
            hll_free(g.H_rm);
      vec_put(&g.x);
      if (!IS_ERR_OR_NULL(g.A))
            csr_free(g.A);
      if (g.logger_open)
            logger_close();
      memset(&g, 0, sizeof g);
}

static void die(void) {
      teardown();
      exit(EXIT_FAILURE);
}

/* With -d: compare with the serial CSR result; releases y either way. */
static void check_and_release(vec *y, const char *what) {
      if (g.debug && g.expected.data &&
          validation_vec_result(g.expected, *y) != 0) {
            vec_put(y);
            LOG_ERR("%s validation failed", what);
            die();
      }
      vec_put(y);
}

static void run_serial(void) {
      bench r;
      int rc = bench_csr_serial(g.A, g.x.data, &r);
      if (rc) {
            LOG_ERR("[CSR serial] failed with error %d", rc);
            die();
      }
      log_csr_serial_benchmark(g.A, r);
      if (g.debug)
            g.expected = r.data; /* ownership moves */
      else
            vec_put(&r.data);

      rc = bench_hll_serial(g.H_rm, g.x.data, &r);
      if (rc) {
            LOG_ERR("[HLL serial] failed with error %d", rc);
            die();
      }
      log_hll_serial_benchmark(g.H_rm, r);
      check_and_release(&r.data, "[HLL serial]");
}

typedef int (*csr_omp_fn)(const sparse_csr *, const double *, bench_omp *);

static void run_csr_omp(csr_omp_fn fn) {
      for (size_t i = 0; i < ARRAY_SIZE(k_omp_threads); ++i) {
            bench_omp b = {.num_threads = k_omp_threads[i]};
            OMP_WARMUP(b.num_threads);
            int rc = fn(g.A, g.x.data, &b);
            if (rc) {
                  LOG_ERR("[CSR OMP] failed with error %d", rc);
                  die();
            }
            log_csr_omp_benchmark(g.A, b);
            check_and_release(&b.bench.data, "[CSR OMP]");
      }
}

static void run_hll_omp(void) {
      for (size_t i = 0; i < ARRAY_SIZE(k_omp_threads); ++i) {
            bench_omp b = {.num_threads = k_omp_threads[i]};
            OMP_WARMUP(b.num_threads);
            int rc = bench_hll_omp(g.H_rm, g.x.data, &b);
            if (rc) {
                  LOG_ERR("[HLL OMP] failed with error %d", rc);
                  die();
            }
            log_hll_omp_benchmark(g.H_rm, b);
            check_and_release(&b.bench.data, "[HLL OMP]");
      }
}

typedef int (*csr_cuda_fn)(const sparse_csr *, const double *, bench_cuda *);
typedef int (*hll_cuda_fn)(const sparse_hll *, const double *, bench_cuda *);

static void run_csr_cuda(void) {
      static const csr_cuda_fn kernels[] = {
          bench_csr_cuda_thread_row,        bench_csr_cuda_warp_row,
          bench_csr_cuda_halfwarp_row,      bench_csr_cuda_block_row,
          bench_csr_cuda_halfwarp_row_text,
      };
      char what[64];
      for (int kid = 0; kid < (int)ARRAY_SIZE(kernels); ++kid) {
            for (size_t w = 0; w < ARRAY_SIZE(k_warps_per_block); ++w) {
                  bench_cuda b = {.warps_per_block = k_warps_per_block[w]};
                  if (kernels[kid](g.A, g.x.data, &b) != 0 ||
                      b.bench.duration_ms <= 0.0) {
                        LOG_ERR("Failed CSR CUDA [kernel %d, warps_per_block "
                                "%d]",
                                kid, b.warps_per_block);
                        vec_put(&b.bench.data);
                        die();
                  }
                  snprintf(what, sizeof what, "CSR CUDA [kernel %d]", kid);
                  check_and_release(&b.bench.data, what);
                  log_csr_cuda_benchmark(g.A, b, kid);
            }
      }
}

static void run_hll_cuda(void) {
      static const hll_cuda_fn kernels[] = {
          bench_hll_cuda_threads_row_major,
          bench_hll_cuda_threads_col_major,
          bench_hll_cuda_warp_block,
          bench_hll_cuda_halfwarp_row,
      };
      char what[64];
      for (int kid = 0; kid < (int)ARRAY_SIZE(kernels); ++kid) {
            /* kernels 0 and 3 take the row-major copy, 1 and 2 the
             * column-major one (reference src/main.c:324-325) */
            const sparse_hll *H = (kid == 0 || kid == 3) ? g.H_rm : g.H_cm;
            for (size_t w = 0; w < ARRAY_SIZE(k_warps_per_block); ++w) {
                  bench_cuda b = {.warps_per_block = k_warps_per_block[w]};
                  if (kernels[kid](H, g.x.data, &b) != 0 ||
                      b.bench.duration_ms <= 0.0) {
                        LOG_ERR("Failed HLL CUDA [kernel %d, warps_per_block "
                                "%d]",
                                kid, b.warps_per_block);
                        vec_put(&b.bench.data);
                        die();
                  }
                  snprintf(what, sizeof what,
                           "HLL CUDA [kernel %d, warps_per_block %d]", kid,
                           b.warps_per_block);
                  check_and_release(&b.bench.data, what);
                  log_hll_cuda_benchmark(H, b, kid);
            }
      }
}

/* Multi-GPU iterated SpMV through the C ABI (include/spmv_b200.h, csrc/dist.cu). */
static void run_multi_gpu(const char *out_dir, int gpus) {
      const char *env = getenv("SPMV_B200_STEPS");
      const int steps = env && atoi(env) > 0 ? atoi(env) : 10;
      if (g.A->M != g.A->N) {
            LOG_ERR("SPMV_B200_GPUS: x_{k+1} = A x_k needs a square matrix, skipped");
            return;
      }
      spmv_b200_dist_group *grp = spmv_b200_dist_group_create(
          g.A, gpus, SPMV_B200_CSR_ADAPTIVE, 4, SPMV_B200_DIST_AUTO);
      if (!grp) {
            LOG_ERR("multi-GPU setup failed: %s", spmv_b200_last_error());
            die();
      }
      vec y = vec_create((size_t)g.A->M);
      double ms = 0.0;
      int rc = y.data ? 0 : -ENOMEM;
      /* one step first: y = A x, comparable with the serial CSR result */
      rc = rc ? rc : spmv_b200_dist_group_set_x(grp, g.x.data);
      rc = rc ? rc : spmv_b200_dist_group_iterate(grp, 1, NULL);
      rc = rc ? rc : spmv_b200_dist_group_get_x(grp, y.data);
      if (!rc && g.debug && g.expected.data &&
          validation_vec_result(g.expected, y) != 0) {
            LOG_ERR("[multi-GPU CSR, %d GPUs] validation failed", gpus);
            rc = -EIO;
      }
      /* then the timed iteration (warm: plans and the CUDA graph exist) */
      rc = rc ? rc : spmv_b200_dist_group_set_x(grp, g.x.data);
      rc = rc ? rc : spmv_b200_dist_group_iterate(grp, 4, NULL);
      rc = rc ? rc : spmv_b200_dist_group_iterate(grp, steps, &ms);
      if (rc) {
            LOG_ERR("multi-GPU run failed: %s", spmv_b200_last_error());
            vec_put(&y);
            spmv_b200_dist_group_destroy(grp);
            die();
      }
      char path[MAX_PATH];
      snprintf(path, sizeof path, "%s/b200_dist.csv", out_dir);
      FILE *probe = fopen(path, "r");
      FILE *f = fopen(path, "a");
      if (f) {
            if (!probe)
                  fprintf(f, "matrix,gpus,exchange,steps,rows,cols,nnz,ms_per_step,gflops\n");
            const int push = spmv_b200_dist_mode(spmv_b200_dist_group_rank(grp, 0)) ==
                             SPMV_B200_DIST_PUSH;
            fprintf(f, "%s,%d,%s,%d,%d,%d,%d,%f,%f\n", g.A->name, gpus,
                    push ? "push" : "nccl", steps, g.A->M, g.A->N, g.A->NZ, ms / steps,
                    compute_gflops(ms / steps, g.A->NZ));
            fclose(f);
      }
      if (probe)
            fclose(probe);
      vec_put(&y);
      spmv_b200_dist_group_destroy(grp);
}

int main(int argc, char **argv) {
      static const struct option long_opts[] = {
          {"matrix", required_argument, NULL, 'm'},
          {"out", required_argument, NULL, 'o'},
          {"bench", required_argument, NULL, 'b'}, /* accepted, unsupported */
          {"debug", no_argument, NULL, 'd'},
          {"help", no_argument, NULL, 'h'},
          {NULL, 0, NULL, 0}};

      const char *matrix_path = NULL, *out_dir = NULL;
      int opt;
      while ((opt = getopt_long(argc, argv, "m:o:b:dh", long_opts, NULL)) !=
             -1) {
            switch (opt) {
            case 'm':
                  matrix_path = optarg;
                  break;
            case 'o':
                  out_dir = optarg;
                  break;
            case 'd':
                  g.debug = true;
                  break;
            case 'h':
                  log_prog_usage(basename(argv[0]));
                  return EXIT_SUCCESS;
            default: /* includes -b, as in the reference */
                  log_prog_usage(basename(argv[0]));
                  return EXIT_FAILURE;
            }
      }
      if (!matrix_path || !out_dir) {
            log_prog_usage(basename(argv[0]));
            return EXIT_FAILURE;
      }

      if (logger_init(out_dir) != 0) {
            LOG_ERR("Failed to open log file: %s", out_dir);
            logger_close();
            return EXIT_FAILURE;
      }
      g.logger_open = true;

      g.A = io_load_csr(matrix_path);
      if (IS_ERR_OR_NULL(g.A)) {
            LOG_ERR("Failed to load matrix: %s (err %d)", matrix_path,
                    PTR_ERR(g.A));
            die();
      }

      g.H_rm = csr_to_hll(g.A, false);
      g.H_cm = csr_to_hll(g.A, true);
      if (IS_ERR_OR_NULL(g.H_rm) || IS_ERR_OR_NULL(g.H_cm)) {
            LOG_ERR("Failed to convert CSR to HLL");
            die();
      }

      g.x = vec_create((size_t)g.A->N);
      if (!g.x.data)
            die();
      vec_fill_random(&g.x);

      const char *skip = getenv("SPMV_B200_SKIP_CPU");
      if (!(skip && skip[0] == '1')) {
            run_serial();
            run_csr_omp(bench_csr_omp_nnz_balancing);
            run_csr_omp(bench_csr_omp_guided);
            run_hll_omp();
      }
      run_csr_cuda();
      run_hll_cuda();
      const char *gpus = getenv("SPMV_B200_GPUS");
      if (gpus && atoi(gpus) >= 1)
            run_multi_gpu(out_dir, atoi(gpus));

      teardown();
      return EXIT_SUCCESS;
}
