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
