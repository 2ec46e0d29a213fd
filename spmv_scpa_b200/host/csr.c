/* csr.c -- Matrix Market -> CSR loader, CPU CSR SpMV paths, GPU trampolines.
 *
 * Fresh code with the observable behaviour of reference src/csr.c:
 *   io_load_csr        :31-171   (same CSR arrays bit for bit, same errors)
 *   csr_free           :173-180
 *   bench_csr_serial / _omp_guided / _omp_nnz_balancing   :201-380
 *   bench_csr_cuda_*   :382-415  (forward to libspmv_b200, cuda_csr.h)
 *
 * The loader reads the entry section of the file ONCE into memory and
 * tokenises it there -- in parallel over line-aligned chunks when the file is
 * large and regular, with a sequential walk as the fallback and as the
 * arbiter of every error (the reference runs fscanf over the file twice); the
 * arithmetic that decides the result -- strtol for indices, strtod for
 * values, both correctly rounded like fscanf -- and the order in which
 * entries land in each row (file order; a symmetric off-diagonal (i,j) is
 * followed immediately by its mirror (j,i)) are the same.
 */
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "csr.h"
#include "cuda_csr.h"
#include "err.h"
#include "mmio.h"
#include "utils.h"
#include "vector.h"

/* ------------------------------------------------------------------ loader */

/* "dir/foo.mtx" -> "foo" (at most MAX_NAME-1 chars). */
static void matrix_name_from_path(const char *path, char out[MAX_NAME]) {
      const char *slash = strrchr(path, '/');
      const char *base = slash ? slash + 1 : path;
      size_t n = strlen(base);
      if (n > 4 && memcmp(base + n - 4, ".mtx", 4) == 0)
            n -= 4;
      if (n > MAX_NAME - 1)
            n = MAX_NAME - 1;
      memcpy(out, base, n);
      out[n] = '\0';
}

/* Read the rest of `f` into a NUL-terminated heap buffer. */
static char *slurp_rest(FILE *f, size_t *len_out) {
      long here = ftell(f);
      if (here < 0 || fseek(f, 0, SEEK_END) != 0)
            return NULL;
      long end = ftell(f);
      if (end < here || fseek(f, here, SEEK_SET) != 0)
            return NULL;
      size_t want = (size_t)(end - here);
      char *buf = malloc(want + 1);
      if (!buf)
            return NULL;
      size_t got = fread(buf, 1, want, f);
      buf[got] = '\0';
      *len_out = got;
      return buf;
}

/* Cursor over the in-memory entry section.  Each reader mirrors one scanf
 * conversion: skip white space, convert, fail if nothing was consumed. */
struct cursor {
      const char *p;
};

static int read_int(struct cursor *c, int *out) {
      char *stop;
      while (isspace((unsigned char)*c->p))
            ++c->p;
      if (*c->p == '\0')
            return -1;
      long v = strtol(c->p, &stop, 10);
      if (stop == c->p)
            return -1;
      c->p = stop;
      *out = (int)v;
      return 0;
}

static int read_double(struct cursor *c, double *out) {
      char *stop;
      while (isspace((unsigned char)*c->p))
            ++c->p;
      if (*c->p == '\0')
            return -1;
      double v = strtod(c->p, &stop);
      if (stop == c->p)
            return -1;
      c->p = stop;
      *out = v;
      return 0;
}

/* ---- entry section: sequential reference walk ---------------------------
 * Parses up to n_decl entries in file order.  Returns the number of entries
 * accepted; *err is 0, or the errno of the first problem (-EIO for a token
 * that does not convert / early end, -ERANGE for an index outside M x N) --
 * the same first-problem-wins order as the reference's pass 1
 * (src/csr.c:68-95). */
static size_t parse_entries_serial(const char *text, size_t n_decl, int pattern,
                                   int M, int N, int *ei, int *ej, double *ev,
                                   int *err) {
      struct cursor cur = {text};
      *err = 0;
      for (size_t e = 0; e < n_decl; ++e) {
            int i, j;
            double v = 1.0;
            if (read_int(&cur, &i) || read_int(&cur, &j) ||
                (!pattern && read_double(&cur, &v))) {
                  *err = -EIO;
                  return e;
            }
            --i, --j;
            if (i < 0 || i >= M || j < 0 || j >= N) {
                  *err = -ERANGE;
                  return e;
            }
            ei[e] = i, ej[e] = j;
            if (!pattern)
                  ev[e] = v;
      }
      return n_decl;
}

/* ---- entry section: parallel walk ---------------------------------------
 * The text is cut at line ends into one chunk per thread.  A first sweep
 * counts white-space separated tokens per chunk; if every chunk starts on an
 * entry boundary (token index divisible by the tokens per entry) the chunks
 * are parsed independently, each writing its entries at their global index.
 * Anything unusual -- a chunk not aligned to entries, a number glued to the
 * next token (which fscanf would split differently), any conversion error --
 * makes the function return -1 and the caller falls back to the sequential
 * walk, so the result is always the one the reference's fscanf loop gives. */
static int parse_entries_parallel(const char *text, size_t len, size_t n_decl,
                                  int pattern, int M, int N, int *ei, int *ej,
                                  double *ev) {
#ifndef _OPENMP
      (void)text, (void)len, (void)n_decl, (void)pattern, (void)M, (void)N;
      (void)ei, (void)ej, (void)ev;
      return -1;
#else
      int nt = omp_get_max_threads();
      if (nt > 64)
            nt = 64;
      if (nt < 2 || len < ((size_t)1 << 20))
            return -1;
      const int tpe = pattern ? 2 : 3;
      size_t cut[65], tok[65];
      cut[0] = 0;
      for (int t = 1; t < nt; ++t) {
            size_t p = len / nt * t;
            if (p < cut[t - 1])
                  p = cut[t - 1];
            while (p < len && text[p] != '\n')
                  ++p;
            cut[t] = p < len ? p + 1 : len;
      }
      cut[nt] = len;

      int bad = 0;
#pragma omp parallel num_threads(nt) reduction(| : bad)
      {
            const int t = omp_get_thread_num();
            if (omp_get_num_threads() != nt)
                  bad |= 1; /* the runtime granted a smaller team: chunks would go unparsed */
            size_t n = 0;
            int in_tok = 0;
            for (size_t p = cut[t]; p < cut[t + 1]; ++p) {
                  const int sp = isspace((unsigned char)text[p]) || text[p] == '\0';
                  n += (!sp && !in_tok);
                  in_tok = !sp;
            }
            tok[t + 1] = n;
#pragma omp barrier
#pragma omp single
            {
                  tok[0] = 0;
                  for (int k = 0; k < nt; ++k)
                        tok[k + 1] += tok[k];
            }
            if (tok[t] % tpe != 0) {
                  bad |= 1;
            } else {
                  size_t e = tok[t] / tpe;
                  const size_t e_end = tok[t + 1] / tpe; /* whole entries in this chunk */
                  struct cursor cur = {text + cut[t]};
                  const char *stop_at = text + cut[t + 1];
                  for (; e < e_end && e < n_decl && !bad; ++e) {
                        int i, j;
                        double v = 1.0;
                        if (read_int(&cur, &i) || !(isspace((unsigned char)*cur.p)) ||
                            read_int(&cur, &j) ||
                            (!pattern && (!isspace((unsigned char)*cur.p) ||
                                          read_double(&cur, &v))) ||
                            !(isspace((unsigned char)*cur.p) || *cur.p == '\0') ||
                            cur.p > stop_at) {
                              bad |= 1;
                              break;
                        }
                        --i, --j;
                        if (i < 0 || i >= M || j < 0 || j >= N) {
                              bad |= 1; /* let the sequential walk pick the errno */
                              break;
                        }
                        ei[e] = i, ej[e] = j;
                        if (!pattern)
                              ev[e] = v;
                  }
            }
      }
      if (bad || tok[nt] / tpe < n_decl) /* short file: sequential walk reports it */
            return -1;
      return 0;
#endif
}

sparse_csr *io_load_csr(const char *path) {
      char name[MAX_NAME];
      MM_typecode tc;
      int M = 0, N = 0, nz_decl = 0;
      int err = 0;

      char *text = NULL;
      int *ei = NULL, *ej = NULL; /* parsed coordinates, file order */
      double *ev = NULL;
      int *fill = NULL; /* per-row counts, then insertion cursors */
      int *IRP = NULL, *JA = NULL;
      double *AS = NULL;
      sparse_csr *A = NULL;

      matrix_name_from_path(path, name);

      FILE *f = fopen(path, "r");
      if (!f)
            return ERR_PTR(-errno);

      if (mm_read_banner(f, &tc) != 0 || !mm_is_matrix(tc) ||
          !mm_is_sparse(tc) || !(mm_is_real(tc) || mm_is_pattern(tc)) ||
          mm_read_mtx_crd_size(f, &M, &N, &nz_decl) != 0) {
            err = -EINVAL;
            goto out;
      }
      const int symmetric = mm_is_symmetric(tc);
      const int pattern = mm_is_pattern(tc);

      size_t text_len = 0;
      text = slurp_rest(f, &text_len);
      const size_t n_decl = nz_decl > 0 ? (size_t)nz_decl : 0;
      ei = malloc((n_decl ? n_decl : 1) * sizeof *ei);
      ej = malloc((n_decl ? n_decl : 1) * sizeof *ej);
      ev = pattern ? NULL : malloc((n_decl ? n_decl : 1) * sizeof *ev);
      fill = calloc(M > 0 ? (size_t)M + 1 : 2, sizeof *fill);
      if (!text || !ei || !ej || (!pattern && !ev) || !fill) {
            err = -ENOMEM;
            goto out;
      }

      /* parse (parallel when the file is large and regular, else sequential) */
      if (parse_entries_parallel(text, text_len, n_decl, pattern, M, N, ei, ej,
                                 ev) != 0) {
            parse_entries_serial(text, n_decl, pattern, M, N, ei, ej, ev, &err);
            if (err)
                  goto out;
      }

      /* count per row (a symmetric off-diagonal also counts for its mirror) */
      long total = 0;
      for (size_t e = 0; e < n_decl; ++e) {
            ++fill[ei[e]], ++total;
            if (symmetric && ei[e] != ej[e])
                  ++fill[ej[e]], ++total;
      }

      IRP = aligned_malloc(((size_t)M + 1) * sizeof *IRP);
      if (!IRP) {
            err = -ENOMEM;
            goto out;
      }
      IRP[0] = 0;
      for (int r = 0; r < M; ++r)
            IRP[r + 1] = IRP[r] + fill[r];

      JA = aligned_malloc((size_t)total * sizeof *JA);
      AS = aligned_malloc((size_t)total * sizeof *AS);
      if (!JA || !AS) {
            err = -ENOMEM;
            goto out;
      }

      /* Scatter in file order.  Rows are dealt to threads in contiguous ranges
       * of ~equal entry count; every thread walks ALL entries and keeps those
       * of its rows, so the order inside a row is the file order whatever the
       * thread count (an entry, then its mirror). */
      {
#ifdef _OPENMP
            int nt = total > (1 << 20) ? omp_get_max_threads() : 1;
#else
            int nt = 1;
#endif
            if (nt > 64)
                  nt = 64;
            int row_cut[65];
            row_cut[0] = 0;
            for (int t = 1; t < nt; ++t) {
                  const long want = total / nt * t;
                  int lo = row_cut[t - 1], hi = M;
                  while (lo < hi) { /* first row whose offset reaches `want` */
                        const int mid = lo + (hi - lo) / 2;
                        if (IRP[mid] < want)
                              lo = mid + 1;
                        else
                              hi = mid;
                  }
                  row_cut[t] = lo;
            }
            row_cut[nt] = M;
            for (int r = 0; r < M; ++r)
                  fill[r] = IRP[r]; /* write cursor of row r */
#pragma omp parallel num_threads(nt)
            {
#ifdef _OPENMP
                  const int t = omp_get_thread_num();
#else
                  const int t = 0;
#endif
                  const int r_lo = row_cut[t], r_hi = row_cut[t + 1];
                  for (size_t e = 0; e < n_decl; ++e) {
                        const int i = ei[e], j = ej[e];
                        const double v = pattern ? 1.0 : ev[e];
                        if (i >= r_lo && i < r_hi) {
                              const int k = fill[i]++;
                              JA[k] = j, AS[k] = v;
                        }
                        if (symmetric && i != j && j >= r_lo && j < r_hi) {
                              const int k = fill[j]++;
                              JA[k] = i, AS[k] = v;
                        }
                  }
            }
      }

      A = malloc(sizeof *A);
      if (!A) {
            err = -ENOMEM;
            goto out;
      }
      init_csr(A, name, M, N, (int)total, IRP, JA, AS);

out:
      free(text);
      free(ei);
      free(ej);
      free(ev);
      free(fill);
      if (err) {
            free(IRP);
            free(JA);
            free(AS);
      }
      fclose(f);
      return err ? ERR_PTR(err) : A;
}

void csr_free(sparse_csr *A) {
      if (!A)
            return;
      free(A->IRP);
      free(A->JA);
      free(A->AS);
      free(A);
}

/* ------------------------------------------------------------ bench glue */

typedef double (*csr_spmv_fn)(const sparse_csr *, const double *, double *,
                              void *);

/* Allocate y, run one variant, convert its milliseconds to GFLOP/s
 * (reference compute_benchmark_csr, src/csr.c:182-199). */
static int run_variant(const sparse_csr *A, const double *x, bench *out,
                       void *arg, csr_spmv_fn fn) {
      vec y = vec_create((size_t)A->M);
      if (!y.data)
            return -ENOMEM;
      const double ms = fn(A, x, y.data, arg);
      out->duration_ms = ms;
      out->gflops = compute_gflops(ms, A->NZ);
      out->data = y;
      return 0;
}

/* ------------------------------------------------------------- CPU paths */
/* Kept so the CLI still emits serial.csv / omp.csv.  Not used by the GPU
 * path and not the parity oracle (that is oracle/, built from the reference
 * sources themselves). */

static inline double row_dot(const sparse_csr *A, const double *x, int r) {
      double acc = 0.0;
      for (int k = A->IRP[r], e = A->IRP[r + 1]; k < e; ++k)
            acc += A->AS[k] * x[A->JA[k]];
      return acc;
}

static double cpu_csr_serial(const sparse_csr *A, const double *x, double *y,
                             void *unused) {
      (void)unused;
      const double t0 = now();
      for (int r = 0; r < A->M; ++r)
            y[r] = row_dot(A, x, r);
      return now() - t0;
}

static double wall_ms(void) {
#ifdef _OPENMP
      return omp_get_wtime() * 1e3;
#else
      struct timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
#endif
}

static double cpu_csr_omp_guided(const sparse_csr *A, const double *x,
                                 double *y, void *arg) {
      const int nt = *(const int *)arg;
      const double t0 = wall_ms();
#pragma omp parallel for schedule(guided) num_threads(nt)
      for (int r = 0; r < A->M; ++r)
            y[r] = row_dot(A, x, r);
      return wall_ms() - t0;
}

/* Contiguous row ranges with ~equal nnz: a range is closed as soon as its
 * running nnz reaches total/parts (reference partition_csr_rows,
 * src/csr.c:218-276).  May use fewer parts than requested; *parts is updated.
 * Returns parts+1 boundaries (malloc) or NULL. */
static int *split_rows_by_nnz(const sparse_csr *A, int *parts) {
      const int want = *parts;
      int *cut = malloc(((size_t)want + 1) * sizeof *cut);
      if (!cut)
            return NULL;
      const double quota = (double)A->IRP[A->M] / want;
      int used = 0;
      double acc = 0.0;
      cut[0] = 0;
      for (int r = 0; r < A->M && used < want - 1; ++r) {
            acc += A->IRP[r + 1] - A->IRP[r];
            if (acc >= quota) {
                  cut[++used] = r + 1;
                  acc = 0.0;
            }
      }
      cut[++used] = A->M;
      *parts = used;
      return cut;
}

struct nnz_split {
      int parts;
      const int *cut;
};

static double cpu_csr_omp_nnz(const sparse_csr *A, const double *x, double *y,
                              void *arg) {
      const struct nnz_split *s = arg;
      const double t0 = wall_ms();
#pragma omp parallel num_threads(s->parts)
      {
#ifdef _OPENMP
            const int t = omp_get_thread_num();
            const int nth = omp_get_num_threads();
#else
            const int t = 0, nth = 1;
#endif
            /* if the runtime granted fewer threads, each takes several parts */
            for (int p = t; p < s->parts; p += nth)
                  for (int r = s->cut[p]; r < s->cut[p + 1]; ++r)
                        y[r] = row_dot(A, x, r);
      }
      return wall_ms() - t0;
}

int bench_csr_serial(const sparse_csr *A, const double *x, bench *out) {
      return run_variant(A, x, out, NULL, cpu_csr_serial);
}

int bench_csr_omp_guided(const sparse_csr *A, const double *x, bench_omp *out) {
      snprintf(out->name, sizeof out->name, "omp_guided");
      return run_variant(A, x, &out->bench, &out->num_threads,
                         cpu_csr_omp_guided);
}

int bench_csr_omp_nnz_balancing(const sparse_csr *A, const double *x,
                                bench_omp *out) {
      int *cut = split_rows_by_nnz(A, &out->num_threads);
      if (!cut)
            return -ENOMEM;
      struct nnz_split s = {out->num_threads, cut};
      const int rc = run_variant(A, x, &out->bench, &s, cpu_csr_omp_nnz);
      free(cut);
      snprintf(out->name, sizeof out->name, "omp_nnz");
      return rc;
}

/* ------------------------------------------------------------- GPU paths */
/* Two-call protocol of the reference (set wpb, then launch) kept as is. */

#define DEFINE_CSR_CUDA_BENCH(suffix)                                          \
      int bench_csr_cuda_##suffix(const sparse_csr *A, const double *x,        \
                                  bench_cuda *out) {                           \
            set_csr_warps_per_block(out->warps_per_block);                     \
            return run_variant(A, x, &out->bench, NULL,                        \
                               csr_spmv_cuda_##suffix);                        \
      }

DEFINE_CSR_CUDA_BENCH(thread_row)
DEFINE_CSR_CUDA_BENCH(warp_row)
DEFINE_CSR_CUDA_BENCH(halfwarp_row)
DEFINE_CSR_CUDA_BENCH(block_row)
DEFINE_CSR_CUDA_BENCH(halfwarp_row_text)
