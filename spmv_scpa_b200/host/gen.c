/* gen.c -- synthetic matrix generators declared in include/spmv_gen.h.
 *
 * New code (the reference has no generators).  Shapes follow BASELINE.json /
 * C1 gen_poisson2d(1000,1000), C2 gen_stencil27(128,128,128),
 * C3 gen_uniform_random(16000000,32,42), C4 gen_rmat(24,16,.57,.19,.19,42),
 * C5 gen_stencil27_rows(512,512,512,...) per shard.
 */
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mmio.h"
#include "spmv_gen.h"
#include "utils.h"

uint64_t gen_mix64(uint64_t z) {
      z += 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
}

/* counter-based draw: (seed, a, b) -> 64 random bits */
static inline uint64_t draw(uint64_t seed, uint64_t a, uint64_t b) {
      return gen_mix64(gen_mix64(seed ^ (a * 0xD1342543DE82EF95ull)) + b);
}

/* 53 random bits -> (-1, 1) */
static inline double to_pm1(uint64_t bits) {
      return ((double)(bits >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

static sparse_csr *alloc_csr(const char *name, int64_t M, int64_t N,
                             int64_t NZ) {
      if (M > INT_MAX - 1 || N > INT_MAX || NZ > INT_MAX)
            return NULL;
      sparse_csr *A = malloc(sizeof *A);
      int *irp = aligned_malloc(((size_t)M + 1) * sizeof(int));
      int *ja = aligned_malloc(((size_t)NZ + 16) * sizeof(int));
      double *as = aligned_malloc(((size_t)NZ + 8) * sizeof(double));
      if (!A || !irp || !ja || !as) {
            free(A), free(irp), free(ja), free(as);
            return NULL;
      }
      init_csr(A, name, (int)M, (int)N, (int)NZ, irp, ja, as);
      return A;
}

/* ----------------------------------------------------------- 2D 5-point */

sparse_csr *gen_poisson2d(int nx, int ny) {
      if (nx < 1 || ny < 1)
            return NULL;
      const int64_t n = (int64_t)nx * ny;
      const int64_t nnz = 5 * n - 2 * (int64_t)nx - 2 * (int64_t)ny;
      char name[MAX_NAME];
      snprintf(name, sizeof name, "poisson2d_%dx%d", nx, ny);
      sparse_csr *A = alloc_csr(name, n, n, nnz);
      if (!A)
            return NULL;
      int k = 0;
      for (int iy = 0; iy < ny; ++iy) {
            for (int ix = 0; ix < nx; ++ix) {
                  const int r = iy * nx + ix;
                  A->IRP[r] = k;
                  if (iy > 0)
                        A->JA[k] = r - nx, A->AS[k++] = -1.0;
                  if (ix > 0)
                        A->JA[k] = r - 1, A->AS[k++] = -1.0;
                  A->JA[k] = r, A->AS[k++] = 4.0;
                  if (ix < nx - 1)
                        A->JA[k] = r + 1, A->AS[k++] = -1.0;
                  if (iy < ny - 1)
                        A->JA[k] = r + nx, A->AS[k++] = -1.0;
            }
      }
      A->IRP[n] = k;
      return A;
}

/* ---------------------------------------------------------- 3D 27-point */

/* neighbours of coordinate c on an axis of length n that lie inside */
static inline int axis_span(int c, int n) {
      return 1 + (c > 0) + (c < n - 1);
}

sparse_csr *gen_stencil27_rows(int nx, int ny, int nz, int64_t row0,
                               int64_t row1) {
      if (nx < 1 || ny < 1 || nz < 1)
            return NULL;
      const int64_t plane = (int64_t)nx * ny, n = plane * nz;
      if (row0 < 0 || row1 > n || row0 > row1)
            return NULL;
      const int64_t rows = row1 - row0;

      /* per-row lengths first (parallel), then a prefix sum */
      if (rows > INT_MAX - 1 || n > INT_MAX)
            return NULL;
      int *len = malloc(((size_t)rows + 1) * sizeof *len);
      if (!len)
            return NULL;
#pragma omp parallel for schedule(static)
      for (int64_t r = 0; r < rows; ++r) {
            const int64_t g = row0 + r;
            const int ix = (int)(g % nx), iy = (int)(g / nx % ny),
                      iz = (int)(g / plane);
            len[r] = axis_span(ix, nx) * axis_span(iy, ny) * axis_span(iz, nz);
      }
      int64_t nnz = 0;
      for (int64_t r = 0; r < rows; ++r)
            nnz += len[r];

      char name[MAX_NAME];
      if (rows == n)
            snprintf(name, sizeof name, "stencil27_%dx%dx%d", nx, ny, nz);
      else
            snprintf(name, sizeof name, "stencil27_%dx%dx%d_r%lld", nx, ny, nz,
                     (long long)row0);
      sparse_csr *A = alloc_csr(name, rows, n, nnz);
      if (!A) {
            free(len);
            return NULL;
      }
      A->IRP[0] = 0;
      for (int64_t r = 0; r < rows; ++r)
            A->IRP[r + 1] = A->IRP[r] + len[r];
      free(len);

#pragma omp parallel for schedule(static)
      for (int64_t r = 0; r < rows; ++r) {
            const int64_t g = row0 + r;
            const int ix = (int)(g % nx), iy = (int)(g / nx % ny),
                      iz = (int)(g / plane);
            int k = A->IRP[r];
            for (int dz = -1; dz <= 1; ++dz) {
                  if (iz + dz < 0 || iz + dz >= nz)
                        continue;
                  for (int dy = -1; dy <= 1; ++dy) {
                        if (iy + dy < 0 || iy + dy >= ny)
                              continue;
                        for (int dx = -1; dx <= 1; ++dx) {
                              if (ix + dx < 0 || ix + dx >= nx)
                                    continue;
                              A->JA[k] = (int)(g + dz * plane + dy * nx + dx);
                              A->AS[k] = (dz | dy | dx) ? -1.0 : 26.0;
                              ++k;
                        }
                  }
            }
      }
      return A;
}

sparse_csr *gen_stencil27(int nx, int ny, int nz) {
      return gen_stencil27_rows(nx, ny, nz, 0, (int64_t)nx * ny * nz);
}

/* ------------------------------------------------------- uniform random */

static int cmp_int(const void *a, const void *b) {
      const int x = *(const int *)a, y = *(const int *)b;
      return (x > y) - (x < y);
}

sparse_csr *gen_uniform_random(int n, int k, uint64_t seed) {
      if (n < 1 || k < 0 || k > n || k > 4096)
            return NULL;
      char name[MAX_NAME];
      snprintf(name, sizeof name, "uniform_n%d_k%d_s%llu", n, k,
               (unsigned long long)seed);
      sparse_csr *A = alloc_csr(name, n, n, (int64_t)n * k);
      if (!A)
            return NULL;
#pragma omp parallel for schedule(static)
      for (int r = 0; r <= n; ++r)
            A->IRP[r] = (int)((int64_t)r * k);

#pragma omp parallel for schedule(static, 4096)
      for (int r = 0; r < n; ++r) {
            int *cols = A->JA + (size_t)r * k;
            double *vals = A->AS + (size_t)r * k;
            /* draw until k distinct columns are accepted */
            int have = 0;
            for (uint64_t t = 0; have < k; ++t) {
                  const int c = (int)(draw(seed, (uint64_t)r, t) % (uint64_t)n);
                  int dup = 0;
                  for (int q = 0; q < have; ++q)
                        dup |= cols[q] == c;
                  if (!dup)
                        cols[have++] = c;
            }
            qsort(cols, (size_t)k, sizeof *cols, cmp_int);
            for (int j = 0; j < k; ++j)
                  vals[j] = to_pm1(
                      draw(seed ^ 0xA5A5A5A5A5A5A5A5ull, (uint64_t)r, (uint64_t)j));
      }
      return A;
}

/* ---------------------------------------------------------------- R-MAT */

sparse_csr *gen_rmat(int scale, int edge_factor, double a, double b, double c,
                     uint64_t seed) {
      if (scale < 1 || scale > 30 || edge_factor < 1)
            return NULL;
      const int64_t n = (int64_t)1 << scale;
      const int64_t m = n * edge_factor;
      if (m > INT_MAX)
            return NULL;
      char name[MAX_NAME];
      snprintf(name, sizeof name, "rmat_s%d_e%d_s%llu", scale, edge_factor,
               (unsigned long long)seed);
      sparse_csr *A = alloc_csr(name, n, n, m);
      int *src = malloc((size_t)m * sizeof *src);
      int *dst = malloc((size_t)m * sizeof *dst);
      int *cursor = calloc((size_t)n + 1, sizeof *cursor);
      if (!A || !src || !dst || !cursor) {
            csr_free(A), free(src), free(dst), free(cursor);
            return NULL;
      }
      /* thresholds on a 32-bit uniform draw, one draw per level */
      const double ab = a + b, abc = a + b + c;
      const uint64_t ta = (uint64_t)(a * 4294967296.0),
                     tab = (uint64_t)(ab * 4294967296.0),
                     tabc = (uint64_t)(abc * 4294967296.0);

#pragma omp parallel for schedule(static, 65536)
      for (int64_t e = 0; e < m; ++e) {
            int i = 0, j = 0;
            uint64_t bits = 0;
            for (int lvl = 0; lvl < scale; ++lvl) {
                  if ((lvl & 1) == 0) /* two 32-bit draws per 64-bit word */
                        bits = draw(seed, (uint64_t)e, (uint64_t)(lvl >> 1));
                  const uint64_t u = bits & 0xFFFFFFFFull;
                  bits >>= 32;
                  const int q = (u >= ta) + (u >= tab) + (u >= tabc);
                  i = (i << 1) | (q >> 1);
                  j = (j << 1) | (q & 1);
            }
            src[e] = i, dst[e] = j;
      }

      /* stable counting sort by row: a row keeps edge-generation order */
      for (int64_t e = 0; e < m; ++e)
            ++cursor[src[e] + 1];
      A->IRP[0] = 0;
      for (int64_t r = 0; r < n; ++r) {
            A->IRP[r + 1] = A->IRP[r] + cursor[r + 1];
            cursor[r] = A->IRP[r];
      }
      for (int64_t e = 0; e < m; ++e) {
            const int k = cursor[src[e]]++;
            A->JA[k] = dst[e];
            A->AS[k] = to_pm1(
                draw(seed ^ 0x5A5A5A5A5A5A5A5Aull, (uint64_t)e, 0xC0FFEEull));
      }
      free(src), free(dst), free(cursor);
      return A;
}

/* --------------------------------------------------------------- ragged */

sparse_csr *gen_ragged(int n, int max_len, uint64_t seed) {
      if (n < 1 || max_len < 0)
            return NULL;
      if (max_len > n)
            max_len = n;
      int64_t nnz = 0;
      for (int r = 0; r < n; ++r)
            nnz += (int)(((int64_t)r * 7919) % (max_len + 1));
      char name[MAX_NAME];
      snprintf(name, sizeof name, "ragged_n%d_w%d", n, max_len);
      sparse_csr *A = alloc_csr(name, n, n, nnz);
      if (!A)
            return NULL;
      int k = 0;
      for (int r = 0; r < n; ++r) {
            const int len = (int)(((int64_t)r * 7919) % (max_len + 1));
            int c0 = r - len / 2;
            if (c0 < 0)
                  c0 = 0;
            if (c0 + len > n)
                  c0 = n - len;
            A->IRP[r] = k;
            for (int j = 0; j < len; ++j, ++k) {
                  A->JA[k] = c0 + j;
                  A->AS[k] = to_pm1(draw(seed, (uint64_t)r, (uint64_t)j));
            }
      }
      A->IRP[n] = k;
      return A;
}

/* ------------------------------------------------------------ .mtx sink */

int gen_write_mtx(const sparse_csr *A, const char *path) {
      FILE *f = fopen(path, "w");
      if (!f)
            return -errno;
      MM_typecode tc = {'M', 'C', 'R', 'G'};
      int rc = mm_write_banner(f, tc);
      fprintf(f, "%% generated by spmv-b200 (%s)\n", A->name);
      rc |= mm_write_mtx_crd_size(f, A->M, A->N, A->NZ);
      for (int r = 0; r < A->M && !rc; ++r)
            for (int k = A->IRP[r]; k < A->IRP[r + 1]; ++k)
                  if (fprintf(f, "%d %d %.17g\n", r + 1, A->JA[k] + 1,
                              A->AS[k]) < 0) {
                        rc = -EIO;
                        break;
                  }
      if (fclose(f) != 0 && !rc)
            rc = -EIO;
      return rc ? -EIO : 0;
}
