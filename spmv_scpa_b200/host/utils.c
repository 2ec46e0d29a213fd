/* utils.c -- usage text, aligned allocation, result validation.
 * Behaviour follows reference src/utils.c:10-60. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "utils.h"

void log_prog_usage(const char *prog) {
      /* The reference prints this as one unterminated line (src/utils.c:10-20);
       * the text is kept, line breaks are added for readability. */
      fprintf(stderr,
              "Usage: %s -m <matrix-file.mtx> -o <out-dir> [-d] [-h]\n"
              "  -m, --matrix <file>   Path to the Matrix Market file to "
              "process\n"
              "  -o, --out <file>      Path where benchmark csv files will be "
              "saved\n"
              "  -d, --debug           Validate results against serial "
              "implementation\n"
              "  -h, --help            Show this help message and exit\n",
              prog);
}

void print_result_vector(const vec res) {
      printf("Result vector y (length %zu)\n", res.len);
      for (size_t i = 0; i < res.len; ++i)
            printf("  y[%zu] = %.4f\n", i, res.data[i]);
      printf("\n");
}

void *aligned_malloc(size_t size) {
      void *p = NULL;
      return posix_memalign(&p, ALIGNMENT, size) == 0 ? p : NULL;
}

int validation_vec_result(const vec expected, const vec res) {
      if (expected.len != res.len)
            return -1;
      double acc = 0.0;
      for (size_t i = 0; i < res.len; ++i) {
            const double d = expected.data[i] - res.data[i];
            acc += d * d;
      }
      return sqrt(acc) > 1e-1 ? -1 : 0;
}

void omp_warmup(int num_threads) {
#pragma omp parallel num_threads(num_threads)
      {
#pragma omp for schedule(guided)
            for (int j = 0; j < 1000000; ++j) {
                  volatile double sink = j * 0.5;
                  (void)sink;
            }
      }
}
