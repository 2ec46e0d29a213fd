/* vector.c -- dense FP64 vector helpers (see include/vector.h).
 * Behaviour follows reference src/vector.c:10-41. */
#include <stdlib.h>
#include <string.h>

#include "utils.h"
#include "vector.h"

vec vec_create(size_t n) {
      vec v = {.len = n, .data = aligned_malloc(n * sizeof(double))};
      if (v.data)
            memset(v.data, 0, n * sizeof(double));
      return v;
}

void vec_put(vec *v) {
      if (v) {
            free(v->data);
            v->data = NULL;
      }
}

void vec_fill(vec *v, double value) {
      if (!v || !v->data)
            return;
      double *p = v->data;
      for (size_t n = v->len; n; --n)
            *p++ = value;
}

void vec_fill_random(vec *v) {
      if (!v || !v->data)
            return;
      /* glibc rand(), default seed: same stream as the reference binary */
      for (size_t i = 0; i < v->len; ++i)
            v->data[i] = (double)rand() / RAND_MAX;
}
