/* spmv_errptr.h -- small negative errno values carried inside pointers.
 *
 * A constructor that can fail returns either a valid object or ERR_PTR(-errno);
 * the top 4095 addresses of the address space are never valid objects, so the
 * two cannot be confused.  Same encoding as the reference (include/err.h:10-12),
 * so objects can cross between the two code bases.
 */
#ifndef SPMV_B200_ERRPTR_H
#define SPMV_B200_ERRPTR_H

#include <stdint.h>
#include <stdio.h>

enum { SPMV_MAX_ERRNO = 4095 };

static inline void *spmv_err_ptr(long code) { return (void *)(intptr_t)code; }
static inline int spmv_ptr_err(const void *p) { return (int)(intptr_t)p; }
static inline int spmv_is_err(const void *p) {
    return (uintptr_t)p > (uintptr_t)(-(intptr_t)(SPMV_MAX_ERRNO + 1));
}

/* the spellings the reference's callers use */
#define ERR_PTR(code) spmv_err_ptr((long)(code))
#define PTR_ERR(p) spmv_ptr_err((const void *)(p))
#define IS_ERR(p) spmv_is_err((const void *)(p))
#define IS_ERR_OR_NULL(p) ((p) == NULL || IS_ERR(p))

/* one line on stderr: "[ERROR] file:line: message" */
#define LOG_ERR(...) spmv_log_line(stderr, "ERROR", __FILE__, __LINE__, __VA_ARGS__)
#define LOG_WARN(...) spmv_log_line(stdout, "WARN ", __FILE__, __LINE__, __VA_ARGS__)
#define LOG_INFO(...) spmv_log_line(stdout, "INFO ", __FILE__, __LINE__, __VA_ARGS__)

#ifdef __cplusplus
extern "C"
#endif
void spmv_log_line(FILE *to, const char *level, const char *file, int line, const char *fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 5, 6)))
#endif
    ;

#endif /* SPMV_B200_ERRPTR_H */
