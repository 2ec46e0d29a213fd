/* vector.h -- the dense FP64 vector that carries x and y.
 *
 * ABI: { size_t len; double *data; } exactly as the reference's `vec`
 * (include/vector.h:6-9).  Storage is 64-byte aligned.
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

vec vec_create(size_t count);            /* zero-filled; .data == NULL when out of memory     */
void vec_put(vec *self);                 /* releases the storage; NULL and double put are fine */
void vec_fill(vec *self, double value);  /* every element = value                              */
void vec_fill_random(vec *self);         /* element i = rand() / RAND_MAX, C library generator,
                                            never seeded here: a fresh process gets the same
                                            x as the reference binary (src/vector.c:36-41)     */

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_VECTOR_H */
