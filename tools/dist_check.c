/* dist_check.c -- multi-GPU iterated SpMV through the C ABI only (no Python in the loop).
 *
 *   dist_check [--gpus N] [--steps K] [--mode auto|push|nccl] [--matrix SPEC] [--bench] [--knob key=value]
 *   SPEC: stencil:NX:NY:NZ (generated per shard in HBM) | c5 (= stencil:512:512:512) |
 *         upper:N:W (upper-banded, structurally NON-symmetric) | uniform:N:K | poisson:NX:NY |
 *         rmat:SCALE:EDGEFACTOR (power-law rows and columns)
 *
 * One process drives N GPUs (spmv_b200_dist_group_*).  x_K = A^K x_0 from the N-GPU run is
 * compared with the same K steps on ONE GPU (resident handle, device buffers): identical kernels
 * per row, so the two must agree to the last bit for halo plans, and within 1e-12 * |row scale|
 * otherwise.  --bench times K steps (CUDA events, longest GPU) against the single-GPU time.
 * Exit code 0 = PASS.  This is also the usage example of the multi-GPU section of spmv_b200.h.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csr.h"
#include "spmv_b200.h"
#include "spmv_gen.h"

static double x0_at(int64_t g) { /* (0,1), a pure function of the global index */
      uint64_t z = (uint64_t)g + 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}

static sparse_csr *upper_banded(int n, int w) {
      int64_t nnz = 0;
      for (int r = 0; r < n; ++r)
            nnz += (n - r < w) ? n - r : w;
      int *irp = aligned_malloc(((size_t)n + 1) * sizeof(int));
      int *ja = aligned_malloc(((size_t)nnz + 16) * sizeof(int));
      double *as = aligned_malloc(((size_t)nnz + 8) * sizeof(double));
      sparse_csr *A = malloc(sizeof *A);
      if (!irp || !ja || !as || !A)
            return NULL;
      int k = 0;
      for (int r = 0; r < n; ++r) {
            irp[r] = k;
            for (int j = 0; j < w && r + j < n; ++j, ++k) {
                  ja[k] = r + j;
                  as[k] = (0.5 + x0_at(7ll * k)) / (2.0 * w); /* row sums below 1: x_k stays bounded */
            }
      }
      irp[n] = k;
      init_csr(A, "upper_banded", n, n, k, irp, ja, as);
      return A;
}

int main(int argc, char **argv) {
      int gpus = spmv_b200_device_count(), steps = 6, bench = 0, mode = SPMV_B200_DIST_AUTO;
      const char *spec = "stencil:48:40:64";
      for (int i = 1; i < argc; ++i) {
            if (!strcmp(argv[i], "--gpus") && i + 1 < argc)
                  gpus = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--steps") && i + 1 < argc)
                  steps = atoi(argv[++i]);
            else if (!strcmp(argv[i], "--matrix") && i + 1 < argc)
                  spec = argv[++i];
            else if (!strcmp(argv[i], "--bench"))
                  bench = 1;
            else if (!strcmp(argv[i], "--knob") && i + 1 < argc) { /* key=value, see spmv_b200_set_knob */
                  char key[64];
                  int val = 0;
                  if (sscanf(argv[++i], "%63[^=]=%d", key, &val) != 2 || spmv_b200_set_knob(key, val)) {
                        fprintf(stderr, "dist_check: bad --knob %s\n", argv[i]);
                        return 2;
                  }
            }
            else if (!strcmp(argv[i], "--mode") && i + 1 < argc) {
                  const char *m = argv[++i];
                  mode = !strcmp(m, "push") ? SPMV_B200_DIST_PUSH
                                            : (!strcmp(m, "nccl") ? SPMV_B200_DIST_NCCL : SPMV_B200_DIST_AUTO);
            }
      }
      if (gpus < 1) {
            fprintf(stderr, "dist_check: no GPU (%s)\n", spmv_b200_last_error());
            return 2;
      }
      if (!strcmp(spec, "c5"))
            spec = "stencil:512:512:512";

      int nx = 0, ny = 0, nz = 0, a = 0, b = 0;
      sparse_csr *A = NULL;
      spmv_b200_dist_group *g = NULL;
      spmv_b200_csr *one = NULL;
      int64_t n = 0, nnz = 0;
      if (sscanf(spec, "stencil:%d:%d:%d", &nx, &ny, &nz) == 3) {
            n = (int64_t)nx * ny * nz;
            g = spmv_b200_dist_group_stencil27(nx, ny, nz, gpus, SPMV_B200_CSR_STREAM, 4, mode);
            spmv_b200_set_device(0);
            one = g ? spmv_b200_csr_gen_stencil27(nx, ny, nz, 0, nz, 0, n, NULL, 0) : NULL;
      } else {
            if (sscanf(spec, "upper:%d:%d", &a, &b) == 2)
                  A = upper_banded(a, b);
            else if (sscanf(spec, "uniform:%d:%d", &a, &b) == 2)
                  A = gen_uniform_random(a, b, 42);
            else if (sscanf(spec, "poisson:%d:%d", &a, &b) == 2)
                  A = gen_poisson2d(a, b);
            else if (sscanf(spec, "rmat:%d:%d", &a, &b) == 2) {
                  A = gen_rmat(a, b, 0.57, 0.19, 0.19, 42);
                  int longest = 1; /* keep |x_k| bounded: row sums of |a| below 1 */
                  for (int r = 0; A && r < A->M; ++r)
                        if (A->IRP[r + 1] - A->IRP[r] > longest)
                              longest = A->IRP[r + 1] - A->IRP[r];
                  for (int k = 0; A && k < A->NZ; ++k)
                        A->AS[k] /= longest;
            }
            if (!A) {
                  fprintf(stderr, "dist_check: cannot build matrix '%s'\n", spec);
                  return 2;
            }
            if (!strncmp(spec, "uniform", 7) || !strncmp(spec, "poisson", 7))
                  for (int k = 0; k < A->NZ; ++k) /* keep |x_k| bounded over the iteration */
                        A->AS[k] /= 8.0 * (!strncmp(spec, "uniform", 7) ? b : 1);
            n = A->N;
            g = spmv_b200_dist_group_create(A, gpus, SPMV_B200_CSR_ADAPTIVE, 4, mode);
            spmv_b200_set_device(0);
            one = g ? spmv_b200_csr_create(A) : NULL;
      }
      if (!g || !one) {
            fprintf(stderr, "dist_check: setup failed: %s\n", spmv_b200_last_error());
            return 1;
      }
      nnz = spmv_b200_csr_nnz(one);
      const int kernel = A ? SPMV_B200_CSR_ADAPTIVE : SPMV_B200_CSR_STREAM;
      spmv_b200_dist *r0 = spmv_b200_dist_group_rank(g, 0);
      printf("# %s: n=%lld nnz=%lld on %d GPU(s), exchange=%s\n", spec, (long long)n, (long long)nnz,
             gpus, spmv_b200_dist_mode(r0) == SPMV_B200_DIST_PUSH ? "push" : "nccl");

      double *x0 = malloc((size_t)n * 8), *xg = malloc((size_t)n * 8), *x1 = malloc((size_t)n * 8);
      for (int64_t i = 0; i < n; ++i)
            x0[i] = x0_at(i);

      /* N GPUs */
      double ms_n = 0.0;
      if (spmv_b200_dist_group_set_x(g, x0) || spmv_b200_dist_group_iterate(g, steps, &ms_n) ||
          spmv_b200_dist_group_get_x(g, xg)) {
            fprintf(stderr, "dist_check: group run failed: %s\n", spmv_b200_last_error());
            return 1;
      }

      /* one GPU, same kernels */
      spmv_b200_set_device(0);
      double *d_a = spmv_b200_dmalloc((size_t)n * 8 + 256), *d_b = spmv_b200_dmalloc((size_t)n * 8 + 256);
      if (!d_a || !d_b)
            return 1;
      spmv_b200_h2d(d_a, x0, (size_t)n * 8, NULL);
      for (int k = 0; k < steps; ++k) {
            if (spmv_b200_csr_spmv(one, kernel, 4, d_a, d_b, NULL)) {
                  fprintf(stderr, "dist_check: single-GPU step failed: %s\n", spmv_b200_last_error());
                  return 1;
            }
            double *t = d_a;
            d_a = d_b, d_b = t;
      }
      spmv_b200_d2h(x1, d_a, (size_t)n * 8, NULL);
      spmv_b200_stream_sync(NULL);

      double worst = 0.0, scale = 0.0;
      int64_t diff_bits = 0;
      for (int64_t i = 0; i < n; ++i)
            scale = fmax(scale, fabs(x1[i]));
      for (int64_t i = 0; i < n; ++i) {
            const double e = fabs(xg[i] - x1[i]);
            if (e > worst)
                  worst = e;
            diff_bits += memcmp(&xg[i], &x1[i], 8) != 0;
            if (!isfinite(xg[i]))
                  worst = INFINITY;
      }
      const int ok = worst <= 1e-12 * steps * fmax(scale, 1e-300);
      printf("steps=%d  max|x_N - x_1| = %.3e (scale %.3e)  entries differing in any bit: %lld  graph=%d  %s\n",
             steps, worst, scale, (long long)diff_bits, spmv_b200_dist_has_graph(r0), ok ? "PASS" : "FAIL");

      if (bench && ok) {
            double ms1[8], med1;
            spmv_b200_csr_time(one, kernel, 4, d_a, d_b, 3, 8, 0, ms1, NULL);
            for (int i = 0; i < 8; ++i)
                  for (int j = i + 1; j < 8; ++j)
                        if (ms1[j] < ms1[i]) {
                              double t = ms1[i];
                              ms1[i] = ms1[j], ms1[j] = t;
                        }
            med1 = ms1[4];
            double best = 1e30;
            for (int rep = 0; rep < 5; ++rep) {
                  spmv_b200_dist_group_set_x(g, x0);
                  spmv_b200_dist_group_iterate(g, 4, NULL); /* warm: plans, graph */
                  double ms = 0;
                  if (spmv_b200_dist_group_iterate(g, steps, &ms))
                        return 1;
                  if (ms / steps < best)
                        best = ms / steps;
                  printf("  rep %d: %.4f ms/step\n", rep, ms / steps);
            }
            printf("1 GPU: %.4f ms/step (%.1f GFLOP/s)   %d GPUs: %.4f ms/step (%.1f GFLOP/s)   "
                   "parallel efficiency %.1f%%\n",
                   med1, 2.0 * nnz / (med1 * 1e6), gpus, best, 2.0 * nnz / (best * 1e6),
                   100.0 * med1 / (gpus * best));
      }
      spmv_b200_dist_group_destroy(g);
      spmv_b200_csr_destroy(one);
      spmv_b200_dfree(d_a), spmv_b200_dfree(d_b);
      free(x0), free(xg), free(x1);
      if (A)
            csr_free(A);
      return ok ? 0 : 1;
}
