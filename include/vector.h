/* vector.h -- dense FP64 vector used for x and y.
 *
 * Same layout and entry points as the reference's include/vector.h:6-18
 * ({size_t len; double *data}); storage is 64-byte aligned and zero filled.
 */
#ifndef SPMV_B200_VECTOR_H
#define SPMV_B200_VECTOR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
      size_t len;
      double *data;
} vec;

/* Allocate n zeroed doubles (64-B aligned). data == NULL on failure. */
vec vec_create(size_t n);
/* Release the storage; safe on NULL / already released vectors. */
void vec_put(vec *v);
/* Set every element to `value`. */
void vec_fill(vec *v, double value);
/* x[i] = rand() / RAND_MAX with the C library generator, never seeded here:
 * the sequence is therefore identical to the reference's (src/vector.c:36-41)
 * when both run in a fresh process. */
void vec_fill_random(vec *v);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_VECTOR_H */
