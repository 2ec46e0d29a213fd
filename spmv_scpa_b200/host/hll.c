/* hll.c -- CSR -> HLL packer, CPU HLL SpMV paths, GPU trampolines.
 *
 * Fresh code with the observable behaviour of reference src/hll.c:
 *   csr_to_hll          :19-95   every hack's M / N / NZ / max_NZ / JA / AS
 *                                 is bit-identical (pads: JA = -1, AS = 0.0;
 *                                 column-major stride = rows of that hack)
 *   hll_free            :97-106
 *   bench_hll_serial / bench_hll_omp             :127-224
 *   bench_hll_cuda_*    :226-256 (forward to libspmv_b200, cuda_hll.h)
 *
 * Storage differs from the reference in one deliberate way: instead of two
 * posix_memalign calls per hack, all JA arrays live in one slab and all AS
 * arrays in another, each hack starting on a 64-byte boundary inside its
 * slab (so every blk->JA / blk->AS keeps the reference's alignment
 * guarantee).  A 16M-row matrix is then two allocations instead of one
 * million, and the GPU upload can stream the slabs.  Packing runs in parallel
 * over hacks.  Which sparse_hll objects own slabs is kept in a process-local
 * registry (the struct itself is ABI and has no spare field), so hll_free()
 * also releases an HLL that was built the reference's way -- one allocation
 * per hack, src/hll.c:60-61 -- correctly; an HLL built HERE must be released
 * by THIS hll_free.
 */
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "csr.h"
#include "cuda_hll.h"
#include "err.h"
#include "hll.h"
#include "vector.h"

#define SLAB_ALIGN_JA (ALIGNMENT / sizeof(int))    /* 16 ints  */
#define SLAB_ALIGN_AS (ALIGNMENT / sizeof(double)) /*  8 doubles */

static inline size_t round_up(size_t v, size_t to) {
      return (v + to - 1) / to * to;
}

/* ---- registry of slab-backed objects: H -> (slab_ja, slab_as) ---------- */
struct slab_owner {
      const sparse_hll *H;
      int *ja;
      double *as;
};
static struct slab_owner *g_owners;
static size_t g_n_owners, g_cap_owners;

static int owners_add(const sparse_hll *H, int *ja, double *as) {
      int rc = 0;
#pragma omp critical(spmv_hll_owners)
      {
            if (g_n_owners == g_cap_owners) {
                  const size_t cap = g_cap_owners ? 2 * g_cap_owners : 16;
                  struct slab_owner *p = realloc(g_owners, cap * sizeof *p);
                  if (p)
                        g_owners = p, g_cap_owners = cap;
                  else
                        rc = -ENOMEM;
            }
            if (!rc)
                  g_owners[g_n_owners++] = (struct slab_owner){H, ja, as};
      }
      return rc;
}

/* 1 and the slab bases if H was built by csr_to_hll() below (entry removed) */
static int owners_take(const sparse_hll *H, int **ja, double **as) {
      int found = 0;
#pragma omp critical(spmv_hll_owners)
      {
            for (size_t i = 0; i < g_n_owners; ++i)
                  if (g_owners[i].H == H) {
                        *ja = g_owners[i].ja, *as = g_owners[i].as;
                        g_owners[i] = g_owners[--g_n_owners];
                        found = 1;
                        break;
                  }
      }
      return found;
}

sparse_hll *csr_to_hll(const sparse_csr *A, bool is_col_major) {
      const int M = A->M;
      const int nb = (M + HACK_SIZE - 1) / HACK_SIZE;

      sparse_hll *H = malloc(sizeof *H);
      if (!H)
            return ERR_PTR(-ENOMEM);
      init_hll(H, A->name, A->M, A->N, A->NZ, nb);

      H->blocks = aligned_malloc((size_t)nb * sizeof *H->blocks);
      size_t *off_ja = malloc(((size_t)nb + 1) * sizeof *off_ja);
      size_t *off_as = malloc(((size_t)nb + 1) * sizeof *off_as);
      if (!H->blocks || !off_ja || !off_as)
            goto nomem;

      /* Shape of every hack, and where it starts inside the slabs. */
      off_ja[0] = off_as[0] = 0;
      for (int b = 0; b < nb; ++b) {
            const int r0 = b * HACK_SIZE;
            const int r1 = r0 + HACK_SIZE < M ? r0 + HACK_SIZE : M;
            int widest = 0;
            for (int r = r0; r < r1; ++r) {
                  const int len = A->IRP[r + 1] - A->IRP[r];
                  if (len > widest)
                        widest = len;
            }
            init_ellpack_block(&H->blocks[b], r1 - r0, A->N,
                               A->IRP[r1] - A->IRP[r0], widest);
            const size_t slots = (size_t)(r1 - r0) * (size_t)widest;
            off_ja[b + 1] = off_ja[b] + round_up(slots, SLAB_ALIGN_JA);
            off_as[b + 1] = off_as[b] + round_up(slots, SLAB_ALIGN_AS);
      }

      /* +1 element so that an all-empty matrix still owns valid pointers */
      int *slab_ja = aligned_malloc((off_ja[nb] + SLAB_ALIGN_JA) * sizeof(int));
      double *slab_as =
          aligned_malloc((off_as[nb] + SLAB_ALIGN_AS) * sizeof(double));
      if (!slab_ja || !slab_as || owners_add(H, slab_ja, slab_as)) {
            free(slab_ja);
            free(slab_as);
            goto nomem;
      }

#pragma omp parallel for schedule(static, 256)
      for (int b = 0; b < nb; ++b) {
            ellpack_block *blk = &H->blocks[b];
            int *ja = blk->JA = slab_ja + off_ja[b];
            double *as = blk->AS = slab_as + off_as[b];
            const int rows = blk->M, width = blk->max_NZ;
            const int r0 = b * HACK_SIZE;

            /* inter-hack alignment gap: keep it defined */
            for (size_t s = (size_t)rows * width; s < off_ja[b + 1] - off_ja[b];
                 ++s)
                  ja[s] = -1;
            for (size_t s = (size_t)rows * width; s < off_as[b + 1] - off_as[b];
                 ++s)
                  as[s] = 0.0;

            for (int i = 0; i < rows; ++i) {
                  const int k0 = A->IRP[r0 + i];
                  const int len = A->IRP[r0 + i + 1] - k0;
                  if (is_col_major) {
                        for (int j = 0; j < len; ++j) {
                              ja[(size_t)j * rows + i] = A->JA[k0 + j];
                              as[(size_t)j * rows + i] = A->AS[k0 + j];
                        }
                        for (int j = len; j < width; ++j) {
                              ja[(size_t)j * rows + i] = -1;
                              as[(size_t)j * rows + i] = 0.0;
                        }
                  } else {
                        int *rja = ja + (size_t)i * width;
                        double *ras = as + (size_t)i * width;
                        memcpy(rja, A->JA + k0, (size_t)len * sizeof(int));
                        memcpy(ras, A->AS + k0, (size_t)len * sizeof(double));
                        for (int j = len; j < width; ++j) {
                              rja[j] = -1;
                              ras[j] = 0.0;
                        }
                  }
            }
      }

      free(off_ja);
      free(off_as);
      return H;

nomem:
      free(off_ja);
      free(off_as);
      free(H->blocks);
      free(H);
      return ERR_PTR(-ENOMEM);
}

void hll_free(sparse_hll *H) {
      if (!H)
            return;
      int *slab_ja = NULL;
      double *slab_as = NULL;
      if (owners_take(H, &slab_ja, &slab_as)) {
            free(slab_ja);
            free(slab_as);
      } else if (H->blocks) {
            /* built elsewhere with one allocation per hack (reference src/hll.c:60-61,
             * released per hack in src/hll.c:97-106) */
            for (int b = 0; b < H->num_blocks; ++b) {
                  free(H->blocks[b].JA);
                  free(H->blocks[b].AS);
            }
      }
      free(H->blocks);
      free(H);
}

/* ------------------------------------------------------------ bench glue */

typedef double (*hll_spmv_fn)(const sparse_hll *, const double *, double *,
                              void *);

static int run_variant(const sparse_hll *H, const double *x, bench *out,
                       void *arg, hll_spmv_fn fn) {
      vec y = vec_create((size_t)H->M);
      if (!y.data)
            return -ENOMEM;
      const double ms = fn(H, x, y.data, arg);
      out->duration_ms = ms;
      out->gflops = compute_gflops(ms, H->NZ);
      out->data = y;
      return 0;
}

/* ------------------------------------------------------------- CPU paths */
/* Row-major hacks, padding skipped by testing JA == -1 (reference
 * src/hll.c:127-150, :178-211). */

static inline void hack_times_x(const ellpack_block *blk, const double *x,
                                double *y_hack) {
      const int width = blk->max_NZ;
      for (int i = 0; i < blk->M; ++i) {
            const int *ja = blk->JA + (size_t)i * width;
            const double *as = blk->AS + (size_t)i * width;
            double acc = 0.0;
            for (int j = 0; j < width; ++j)
                  if (ja[j] != -1)
                        acc += as[j] * x[ja[j]];
            y_hack[i] = acc;
      }
}

static double cpu_hll_serial(const sparse_hll *H, const double *x, double *y,
                             void *unused) {
      (void)unused;
      const double t0 = now();
      for (int b = 0; b < H->num_blocks; ++b)
            hack_times_x(&H->blocks[b], x, y + (size_t)b * HACK_SIZE);
      return now() - t0;
}

static double cpu_hll_omp(const sparse_hll *H, const double *x, double *y,
                          void *arg) {
      const int nt = *(const int *)arg;
#ifdef _OPENMP
      const double t0 = omp_get_wtime();
#else
      const double t0 = now() * 1e-3;
#endif
#pragma omp parallel for schedule(guided) num_threads(nt)
      for (int b = 0; b < H->num_blocks; ++b)
            hack_times_x(&H->blocks[b], x, y + (size_t)b * HACK_SIZE);
#ifdef _OPENMP
      return (omp_get_wtime() - t0) * 1e3;
#else
      return (now() * 1e-3 - t0) * 1e3;
#endif
}

int bench_hll_serial(const sparse_hll *H, const double *x, bench *out) {
      return run_variant(H, x, out, NULL, cpu_hll_serial);
}

int bench_hll_omp(const sparse_hll *H, const double *x, bench_omp *out) {
      snprintf(out->name, sizeof out->name, "omp_guided");
      return run_variant(H, x, &out->bench, &out->num_threads, cpu_hll_omp);
}

/* ------------------------------------------------------------- GPU paths */

#define DEFINE_HLL_CUDA_BENCH(suffix)                                          \
      int bench_hll_cuda_##suffix(const sparse_hll *H, const double *x,        \
                                  bench_cuda *out) {                           \
            set_hll_warps_per_block(out->warps_per_block);                     \
            return run_variant(H, x, &out->bench, NULL,                        \
                               hll_spmv_cuda_##suffix);                        \
      }

DEFINE_HLL_CUDA_BENCH(threads_row_major)
DEFINE_HLL_CUDA_BENCH(threads_col_major)
DEFINE_HLL_CUDA_BENCH(warp_block)
DEFINE_HLL_CUDA_BENCH(halfwarp_row)
