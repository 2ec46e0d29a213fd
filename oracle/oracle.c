/* oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference algorithms.
 *
 * NOT product code.  Nothing under spmv_scpa_b200/ links, loads or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may, and only as the checker / the CPU side of a
 * comparison.
 *
 * Each function restates, in plain strict-IEEE C (no -ffast-math), what the
 * reference computes on the SpMV hot path, citing the lines it follows.
 * PARITY PINNING: tests/test_oracle_pinned.py checks every function here
 * (a) against the reference's own code compiled into oracle/_ref/ (see
 * Makefile) on the seeded inputs of tests/golden/make_golden.py, and
 * (b) against the golden vectors committed in tests/golden/ (npz files), which were
 * produced by that reference build.  The reference ships no tests or golden
 * vectors of its own, so reference-run fixtures are the pin.
 *
 * Flat array conventions (no structs cross this ABI):
 *   CSR : irp[M+1], ja[NZ], as[NZ]
 *   HLL : per hack b  rows[b], width[b], nz[b], off[b]  (off[nb] = slots),
 *         ja[slots], as[slots]; inside a hack the layout is row-major
 *         (i*width + j) or column-major (j*rows + i).
 */
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_HACK 32 /* reference include/hll.h:10 */

/* ---------------------------------------------------------------- loader */

/* Banner check of reference src/mmio.c:93-166 reduced to what
 * src/csr.c:48-59 accepts: "%%MatrixMarket matrix coordinate {real|pattern}
 * <any symmetry>".  Returns 0 and sets *pattern / *symmetric, or -1. */
static int orc_banner(FILE *f, int *pattern, int *symmetric) {
      char line[1025], w[5][64];
      if (!fgets(line, sizeof line, f))
            return -1;
      if (sscanf(line, "%63s %63s %63s %63s %63s", w[0], w[1], w[2], w[3],
                 w[4]) != 5)
            return -1;
      if (strncmp(w[0], "%%MatrixMarket", 14) != 0)
            return -1;
      for (int t = 1; t < 5; ++t)
            for (char *p = w[t]; *p; ++p)
                  if (*p >= 'A' && *p <= 'Z')
                        *p += 'a' - 'A';
      if (strcmp(w[1], "matrix") || strcmp(w[2], "coordinate"))
            return -1;
      if (!strcmp(w[3], "real"))
            *pattern = 0;
      else if (!strcmp(w[3], "pattern"))
            *pattern = 1;
      else
            return -1; /* complex / integer: rejected by src/csr.c:50 */
      if (!strcmp(w[4], "symmetric"))
            *symmetric = 1;
      else if (!strcmp(w[4], "general") || !strcmp(w[4], "hermitian") ||
               !strcmp(w[4], "skew-symmetric"))
            *symmetric = 0; /* src/csr.c:58: only 'S' mirrors entries */
      else
            return -1;
      return 0;
}

/* One coordinate line via fscanf, as src/csr.c:70-80 / :124-134. */
static int orc_entry(FILE *f, int pattern, int *i, int *j, double *v) {
      if (pattern) {
            *v = 1.0;
            return fscanf(f, "%d %d", i, j) == 2 ? 0 : -1;
      }
      return fscanf(f, "%d %d %lf", i, j, v) == 3 ? 0 : -1;
}

/* Matrix Market -> CSR following reference src/csr.c:31-171: two fscanf
 * passes; pass 1 counts per row (+ the mirrored entry of a symmetric
 * off-diagonal) and validates indices, IRP is the prefix sum, pass 2 appends
 * each entry to its row in file order, the mirror right after the entry.
 * Outputs are malloc'd; returns 0 or a negative errno like the reference. */
int orc_load_mtx(const char *path, int *M_out, int *N_out, int *NZ_out,
                 int **irp_out, int **ja_out, double **as_out) {
      FILE *f = fopen(path, "r");
      if (!f)
            return -errno;
      int pattern = 0, symmetric = 0, M = 0, N = 0, nz = 0, rc = 0;
      int *count = NULL, *irp = NULL, *ja = NULL;
      double *as = NULL;
      char line[1025];

      if (orc_banner(f, &pattern, &symmetric)) {
            rc = -EINVAL;
            goto done;
      }
      /* size line: skip '%' lines (src/mmio.c:183-187) */
      do {
            if (!fgets(line, sizeof line, f)) {
                  rc = -EINVAL;
                  goto done;
            }
      } while (line[0] == '%');
      if (sscanf(line, "%d %d %d", &M, &N, &nz) != 3 &&
          fscanf(f, "%d %d %d", &M, &N, &nz) != 3) {
            rc = -EINVAL;
            goto done;
      }

      long body = ftell(f);
      count = calloc(M > 0 ? (size_t)M : 1, sizeof *count);
      long total = 0;
      for (int e = 0; e < nz; ++e) {
            int i, j;
            double v;
            if (orc_entry(f, pattern, &i, &j, &v)) {
                  rc = -EIO;
                  goto done;
            }
            if (i < 1 || i > M || j < 1 || j > N) {
                  rc = -ERANGE;
                  goto done;
            }
            count[i - 1]++, total++;
            if (symmetric && i != j)
                  count[j - 1]++, total++;
      }

      irp = malloc(((size_t)M + 1) * sizeof *irp);
      ja = malloc((total ? (size_t)total : 1) * sizeof *ja);
      as = malloc((total ? (size_t)total : 1) * sizeof *as);
      irp[0] = 0;
      for (int r = 0; r < M; ++r)
            irp[r + 1] = irp[r] + count[r];
      memset(count, 0, (M > 0 ? (size_t)M : 1) * sizeof *count);

      fseek(f, body, SEEK_SET);
      for (int e = 0; e < nz; ++e) {
            int i, j;
            double v;
            if (orc_entry(f, pattern, &i, &j, &v)) {
                  rc = -EIO;
                  goto done;
            }
            --i, --j;
            long at = (long)irp[i] + count[i]++;
            ja[at] = j, as[at] = v;
            if (symmetric && i != j) {
                  at = (long)irp[j] + count[j]++;
                  ja[at] = i, as[at] = v;
            }
      }
      *M_out = M, *N_out = N, *NZ_out = (int)total;
      *irp_out = irp, *ja_out = ja, *as_out = as;

done:
      free(count);
      if (rc)
            free(irp), free(ja), free(as);
      fclose(f);
      return rc;
}

void orc_free(void *p) { free(p); }

/* ---------------------------------------------------------------- packer */

/* Shapes of all hacks (reference src/hll.c:39-56): rows, widest row, entries,
 * and the running slot offset.  Arrays have nb (= ceil(M/32)) entries, off
 * has nb+1.  Returns total slots. */
int64_t orc_hll_shape(int M, const int *irp, int *rows, int *width, int *nz,
                      int64_t *off) {
      const int nb = (M + ORC_HACK - 1) / ORC_HACK;
      off[0] = 0;
      for (int b = 0; b < nb; ++b) {
            const int lo = b * ORC_HACK;
            const int hi = lo + ORC_HACK < M ? lo + ORC_HACK : M;
            int widest = 0, entries = 0;
            for (int r = lo; r < hi; ++r) {
                  const int len = irp[r + 1] - irp[r];
                  entries += len;
                  widest = len > widest ? len : widest;
            }
            rows[b] = hi - lo, width[b] = widest, nz[b] = entries;
            off[b + 1] = off[b] + (int64_t)(hi - lo) * widest;
      }
      return off[nb];
}

/* Fill ja/as (slots entries, see orc_hll_shape) exactly as reference
 * src/hll.c:73-91: everything -1 / 0.0 first, then row i, entry j of the
 * hack goes to i*width+j (row-major) or j*rows+i (column-major). */
void orc_hll_pack(int M, const int *irp, const int *cja, const double *cas,
                  int col_major, const int64_t *off, int *ja, double *as) {
      const int nb = (M + ORC_HACK - 1) / ORC_HACK;
      for (int b = 0; b < nb; ++b) {
            const int lo = b * ORC_HACK;
            const int hi = lo + ORC_HACK < M ? lo + ORC_HACK : M;
            const int rows = hi - lo;
            const int64_t slots = off[b + 1] - off[b];
            const int width = rows ? (int)(slots / rows) : 0;
            int *bj = ja + off[b];
            double *ba = as + off[b];
            for (int64_t s = 0; s < slots; ++s)
                  bj[s] = -1, ba[s] = 0.0;
            for (int i = 0; i < rows; ++i) {
                  const int k0 = irp[lo + i], len = irp[lo + i + 1] - k0;
                  for (int j = 0; j < len; ++j) {
                        const int64_t s = col_major ? (int64_t)j * rows + i
                                                    : (int64_t)i * width + j;
                        bj[s] = cja[k0 + j];
                        ba[s] = cas[k0 + j];
                  }
            }
      }
}

/* The device-side padding convention, reference src/cuda_hll.cu:173-195: in
 * the copy that goes to the GPU every JA == -1 becomes the column of the
 * previous slot of the same row, or 0 when the row has no previous slot. */
void orc_hll_patch_pads(int nb, const int *rows, const int *width,
                        const int64_t *off, int col_major, int *ja) {
      for (int b = 0; b < nb; ++b) {
            int *bj = ja + off[b];
            for (int i = 0; i < rows[b]; ++i)
                  for (int j = 0; j < width[b]; ++j) {
                        const int64_t s = col_major ? (int64_t)j * rows[b] + i
                                                    : (int64_t)i * width[b] + j;
                        if (bj[s] != -1)
                              continue;
                        if (j == 0)
                              bj[s] = 0;
                        else
                              bj[s] = col_major
                                          ? bj[(int64_t)(j - 1) * rows[b] + i]
                                          : bj[s - 1];
                  }
      }
}

/* ------------------------------------------------------------------ SpMV */

/* y = A x, the textbook row loop of reference src/csr.c:201-216, summed
 * strictly left to right in FP64.  (The reference binary is built with
 * -ffast-math, so ITS sums are re-associated; the parity tolerance
 * 1e-12 * sum|a x| covers both orders.) */
void orc_csr_spmv(int M, const int *irp, const int *ja, const double *as,
                  const double *x, double *y) {
      for (int r = 0; r < M; ++r) {
            double acc = 0.0;
            for (int k = irp[r]; k < irp[r + 1]; ++k)
                  acc += as[k] * x[ja[k]];
            y[r] = acc;
      }
}

/* 64-bit offsets variant for shards beyond 2^31 entries. */
void orc_csr_spmv64(int64_t M, const int64_t *irp, const int *ja,
                    const double *as, const double *x, double *y) {
#pragma omp parallel for schedule(static, 4096)
      for (int64_t r = 0; r < M; ++r) {
            double acc = 0.0;
            for (int64_t k = irp[r]; k < irp[r + 1]; ++k)
                  acc += as[k] * x[ja[k]];
            y[r] = acc;
      }
}

/* bound[r] = sum_j |a_rj * x_j| : the per-row scale of the tolerance
 * |y - y_ref| <= 1e-12 * bound (BASELINE.json north_star). */
void orc_csr_abs_bound(int M, const int *irp, const int *ja, const double *as,
                       const double *x, double *bound) {
#pragma omp parallel for schedule(static, 4096)
      for (int r = 0; r < M; ++r) {
            double acc = 0.0;
            for (int k = irp[r]; k < irp[r + 1]; ++k)
                  acc += fabs(as[k] * x[ja[k]]);
            bound[r] = acc;
      }
}

/* Row-major HLL times x, skipping pads by JA == -1 (reference
 * src/hll.c:127-150); col_major != 0 follows the unused column-major twin
 * (:152-176). */
void orc_hll_spmv(int nb, const int *rows, const int *width, const int64_t *off,
                  int col_major, const int *ja, const double *as,
                  const double *x, double *y) {
      for (int b = 0; b < nb; ++b) {
            const int *bj = ja + off[b];
            const double *ba = as + off[b];
            for (int i = 0; i < rows[b]; ++i) {
                  double acc = 0.0;
                  for (int j = 0; j < width[b]; ++j) {
                        const int64_t s = col_major ? (int64_t)j * rows[b] + i
                                                    : (int64_t)i * width[b] + j;
                        if (bj[s] != -1)
                              acc += ba[s] * x[bj[s]];
                  }
                  y[(int64_t)b * ORC_HACK + i] = acc;
            }
      }
}

/* Greedy nnz-balanced contiguous row split, reference partition_csr_rows
 * (src/csr.c:218-276): close a part once its running nnz >= total/parts;
 * at most `parts` parts, possibly fewer.  cut[] needs parts+1 entries.
 * Returns the number of parts actually used. */
int orc_partition_rows(int M, const int *irp, int parts, int *cut) {
      const double target = (double)irp[M] / parts;
      int used = 0;
      double running = 0.0;
      cut[0] = 0;
      for (int r = 0; r < M && used < parts - 1; ++r) {
            running += irp[r + 1] - irp[r];
            if (running >= target) {
                  cut[++used] = r + 1;
                  running = 0.0;
            }
      }
      cut[used + 1] = M;
      return used + 1;
}

/* --------------------------------------------- timed CPU baseline ("port") */
/* Used by bench.py when oracle/_ref is unavailable: the same loops, timed,
 * with `threads` OpenMP threads (1 = the serial path).  Returns ms. */
double orc_csr_spmv_timed(int M, const int *irp, const int *ja,
                          const double *as, const double *x, double *y,
                          int threads) {
#ifdef _OPENMP
      const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(guided) num_threads(threads) if (threads > 1)
      for (int r = 0; r < M; ++r) {
            double acc = 0.0;
            for (int k = irp[r]; k < irp[r + 1]; ++k)
                  acc += as[k] * x[ja[k]];
            y[r] = acc;
      }
      return (omp_get_wtime() - t0) * 1e3;
#else
      (void)threads;
      orc_csr_spmv(M, irp, ja, as, x, y);
      return 0.0;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
      return omp_get_max_threads();
#else
      return 1;
#endif
}

/* ------------------------------------------------ synthetic inputs (oracle side) */
/* bench.py --impl reference must not map any product library, so the BASELINE.json
 * inputs are restated here: same definitions as include/spmv_gen.h (the reference
 * itself has no generators -- it downloads SuiteSparse files,
 * scripts/download-matrices.py).  tests/test_generators.py checks that both sides
 * produce identical arrays.  Callers size the arrays with the *_nnz helpers. */

static uint64_t orc_mix64(uint64_t z) { /* splitmix64 finaliser */
      z += 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
}
static uint64_t orc_draw(uint64_t seed, uint64_t a, uint64_t b) {
      return orc_mix64(orc_mix64(seed ^ (a * 0xD1342543DE82EF95ull)) + b);
}
static double orc_pm1(uint64_t bits) {
      return ((double)(bits >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}
static int orc_span(int c, int n) { return 1 + (c > 0) + (c < n - 1); }

/* 27-point stencil, rows [row0,row1): entries of those rows (global columns). */
int64_t orc_stencil27_nnz(int nx, int ny, int nz, int64_t row0, int64_t row1) {
      const int64_t plane = (int64_t)nx * ny;
      int64_t nnz = 0;
#pragma omp parallel for reduction(+ : nnz) schedule(static)
      for (int64_t g = row0; g < row1; ++g)
            nnz += (int64_t)orc_span((int)(g % nx), nx) * orc_span((int)(g / nx % ny), ny) *
                   orc_span((int)(g / plane), nz);
      return nnz;
}

void orc_gen_stencil27_rows(int nx, int ny, int nz, int64_t row0, int64_t row1, int *irp, int *ja,
                            double *as) {
      const int64_t plane = (int64_t)nx * ny, rows = row1 - row0;
      irp[0] = 0;
      for (int64_t r = 0; r < rows; ++r) {
            const int64_t g = row0 + r;
            irp[r + 1] = irp[r] + orc_span((int)(g % nx), nx) * orc_span((int)(g / nx % ny), ny) *
                                      orc_span((int)(g / plane), nz);
      }
#pragma omp parallel for schedule(static)
      for (int64_t r = 0; r < rows; ++r) {
            const int64_t g = row0 + r;
            const int ix = (int)(g % nx), iy = (int)(g / nx % ny), iz = (int)(g / plane);
            int k = irp[r];
            for (int dz = -1; dz <= 1; ++dz)
                  for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                              if (iz + dz < 0 || iz + dz >= nz || iy + dy < 0 || iy + dy >= ny ||
                                  ix + dx < 0 || ix + dx >= nx)
                                    continue;
                              ja[k] = (int)(g + dz * plane + dy * nx + dx);
                              as[k] = (dz | dy | dx) ? -1.0 : 26.0;
                              ++k;
                        }
      }
}

/* 2D 5-point Laplacian (diag 4, neighbours -1), nnz = 5n - 2nx - 2ny. */
void orc_gen_poisson2d(int nx, int ny, int *irp, int *ja, double *as) {
      int k = 0;
      for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix) {
                  const int r = iy * nx + ix;
                  irp[r] = k;
                  if (iy > 0)
                        ja[k] = r - nx, as[k++] = -1.0;
                  if (ix > 0)
                        ja[k] = r - 1, as[k++] = -1.0;
                  ja[k] = r, as[k++] = 4.0;
                  if (ix < nx - 1)
                        ja[k] = r + 1, as[k++] = -1.0;
                  if (iy < ny - 1)
                        ja[k] = r + nx, as[k++] = -1.0;
            }
      irp[nx * ny] = k;
}

static int orc_cmp_int(const void *a, const void *b) {
      const int x = *(const int *)a, y = *(const int *)b;
      return (x > y) - (x < y);
}

/* Rows [row0,row1) of the n x n uniform-random matrix with k distinct sorted columns per row. */
void orc_gen_uniform_rows(int n, int k, uint64_t seed, int row0, int row1, int *irp, int *ja,
                          double *as) {
      for (int r = row0; r <= row1; ++r)
            irp[r - row0] = (int)((int64_t)(r - row0) * k);
#pragma omp parallel for schedule(static, 4096)
      for (int r = row0; r < row1; ++r) {
            int *cols = ja + (size_t)(r - row0) * k;
            double *vals = as + (size_t)(r - row0) * k;
            int have = 0;
            for (uint64_t t = 0; have < k; ++t) {
                  const int c = (int)(orc_draw(seed, (uint64_t)r, t) % (uint64_t)n);
                  int dup = 0;
                  for (int q = 0; q < have; ++q)
                        dup |= cols[q] == c;
                  if (!dup)
                        cols[have++] = c;
            }
            qsort(cols, (size_t)k, sizeof *cols, orc_cmp_int);
            for (int j = 0; j < k; ++j)
                  vals[j] = orc_pm1(orc_draw(seed ^ 0xA5A5A5A5A5A5A5A5ull, (uint64_t)r, (uint64_t)j));
      }
}

/* R-MAT, 2^scale vertices, edge_factor * 2^scale edges, duplicates kept, rows in edge order.
 * 0 or -ENOMEM. */
int orc_gen_rmat(int scale, int edge_factor, double a, double b, double c, uint64_t seed, int *irp,
                 int *ja, double *as) {
      const int64_t n = (int64_t)1 << scale, m = n * edge_factor;
      int *src = malloc((size_t)m * sizeof *src), *dst = malloc((size_t)m * sizeof *dst);
      int *cursor = calloc((size_t)n + 1, sizeof *cursor);
      if (!src || !dst || !cursor) {
            free(src), free(dst), free(cursor);
            return -ENOMEM;
      }
      const uint64_t ta = (uint64_t)(a * 4294967296.0), tab = (uint64_t)((a + b) * 4294967296.0),
                     tabc = (uint64_t)((a + b + c) * 4294967296.0);
#pragma omp parallel for schedule(static, 65536)
      for (int64_t e = 0; e < m; ++e) {
            int i = 0, j = 0;
            uint64_t bits = 0;
            for (int lvl = 0; lvl < scale; ++lvl) {
                  if ((lvl & 1) == 0)
                        bits = orc_draw(seed, (uint64_t)e, (uint64_t)(lvl >> 1));
                  const uint64_t u = bits & 0xFFFFFFFFull;
                  bits >>= 32;
                  const int q = (u >= ta) + (u >= tab) + (u >= tabc);
                  i = (i << 1) | (q >> 1);
                  j = (j << 1) | (q & 1);
            }
            src[e] = i, dst[e] = j;
      }
      for (int64_t e = 0; e < m; ++e)
            ++cursor[src[e] + 1];
      irp[0] = 0;
      for (int64_t r = 0; r < n; ++r) {
            irp[r + 1] = irp[r] + cursor[r + 1];
            cursor[r] = irp[r];
      }
      for (int64_t e = 0; e < m; ++e) {
            const int k = cursor[src[e]]++;
            ja[k] = dst[e];
            as[k] = orc_pm1(orc_draw(seed ^ 0x5A5A5A5A5A5A5A5Aull, (uint64_t)e, 0xC0FFEEull));
      }
      free(src), free(dst), free(cursor);
      return 0;
}
