/* support.c -- the small services of the host layer: log lines, aligned memory, the dense
 * vector, the validation gate, the usage text, the OpenMP warm-up.
 *
 * Observable behaviour follows the reference's src/utils.c:10-60 and src/vector.c:10-41
 * (same usage wording, same 64-byte alignment, same 0.1 threshold on the L2 distance, same
 * unseeded rand() stream for x); the code is this project's own.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "utils.h"
#include "vector.h"

/* ------------------------------------------------------------------ log lines */

void spmv_log_line(FILE *to, const char *level, const char *file, int line, const char *fmt, ...) {
    va_list ap;
    fprintf(to, "[%s] %s:%d: ", level, file, line);
    va_start(ap, fmt);
    vfprintf(to, fmt, ap);
    va_end(ap);
    fputc('\n', to);
}

void log_prog_usage(const char *prog) {
    static const char *const options[][2] = {
        {"-m, --matrix <file>", "Path to the Matrix Market file to process"},
        {"-o, --out <file>", "Path where benchmark csv files will be saved"},
        {"-d, --debug", "Validate results against serial implementation"},
        {"-h, --help", "Show this help message and exit"},
    };
    fprintf(stderr, "Usage: %s -m <matrix-file.mtx> -o <out-dir> [-d] [-h]\n", prog);
    for (size_t i = 0; i < ARRAY_SIZE(options); ++i)
        fprintf(stderr, "  %-22s%s\n", options[i][0], options[i][1]);
}

/* --------------------------------------------------------------------- memory */

void *aligned_malloc(size_t bytes) {
    void *block;
    if (posix_memalign(&block, ALIGNMENT, bytes) != 0)
        block = NULL;
    return block;
}

/* --------------------------------------------------------------------- vector */

vec vec_create(size_t count) {
    vec out;
    out.len = count;
    out.data = aligned_malloc(count * sizeof(double));
    if (out.data != NULL)
        memset(out.data, 0, count * sizeof(double));
    return out;
}

void vec_put(vec *self) {
    if (self == NULL)
        return;
    free(self->data);
    self->data = NULL;
}

void vec_fill(vec *self, double value) {
    if (self == NULL || self->data == NULL)
        return;
    for (double *p = self->data, *end = p + self->len; p != end; ++p)
        *p = value;
}

void vec_fill_random(vec *self) {
    if (self == NULL || self->data == NULL)
        return;
    /* glibc rand() with its default seed: identical to the reference's x in a fresh process */
    for (double *p = self->data, *end = p + self->len; p != end; ++p)
        *p = (double)rand() / RAND_MAX;
}

void print_result_vector(const vec y) {
    printf("Result vector y (length %zu)\n", y.len);
    for (size_t k = 0; k < y.len; ++k)
        printf("  y[%zu] = %.4f\n", k, y.data[k]);
    putchar('\n');
}

/* ----------------------------------------------------------------- validation */

int validation_vec_result(const vec expected, const vec got) {
    if (expected.len != got.len)
        return -1;
    double sum_sq = 0.0;
    for (size_t k = 0; k < got.len; ++k) {
        const double delta = expected.data[k] - got.data[k];
        sum_sq += delta * delta;
    }
    return sqrt(sum_sq) <= 0.1 ? 0 : -1;
}

/* -------------------------------------------------------------- OpenMP warm-up */

void omp_warmup(int num_threads) {
    volatile double sink = 0.0;
#pragma omp parallel for schedule(guided) num_threads(num_threads)
    for (int k = 0; k < 1000000; ++k) {
        if (k == -1) /* never: keeps the loop from being optimised away */
            sink = k * 0.5;
    }
    (void)sink;
}
