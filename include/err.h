/* err.h -- errno-in-pointer helpers and the stderr error line.
 *
 * ABI/behaviour mirror of the reference's include/err.h:10-18: a failing
 * constructor returns a small negative errno cast to a pointer; callers test
 * it with IS_ERR and decode it with PTR_ERR.  Written fresh for spmv-b200.
 */
#ifndef SPMV_B200_ERR_H
#define SPMV_B200_ERR_H

#include <stdint.h>
#include <stdio.h>

/* Highest errno the encoding can carry (same window as the reference). */
#define SPMV_MAX_ERRNO 4095

#define ERR_PTR(code) ((void *)(intptr_t)(code))
#define PTR_ERR(p) ((int)(intptr_t)(p))
#define IS_ERR(p) ((uintptr_t)(p) > (uintptr_t)(-(SPMV_MAX_ERRNO + 1)))
/* Convenience the reference lacks: NULL *or* encoded error. */
#define IS_ERR_OR_NULL(p) ((p) == NULL || IS_ERR(p))

#define LOG_ERR(fmt, ...)                                                      \
      do {                                                                     \
            fprintf(stderr, "[ERROR] %s:%d: " fmt "\n", __FILE__, __LINE__,    \
                    ##__VA_ARGS__);                                            \
      } while (0)

#endif /* SPMV_B200_ERR_H */
