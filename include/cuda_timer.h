/* cuda_timer.h -- CUDA-event stopwatch with C linkage.
 *
 * The reference keeps this API C++-mangled and private to its .cu files
 * (include/cuda_timer.cuh:6-21, src/cuda_timer.cu:3-26).  libspmv_b200
 * exports it with C linkage so plain-C hosts can time device work.  The
 * struct is two opaque event handles (cudaEvent_t is a pointer type, so the
 * layout equals the reference's {cudaEvent_t start, stop}); streams are passed
 * as void* (a cudaStream_t; NULL = default stream).
 */
#ifndef SPMV_B200_CUDA_TIMER_H
#define SPMV_B200_CUDA_TIMER_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
      void *start; /* cudaEvent_t */
      void *stop;  /* cudaEvent_t */
} cuda_timer;

/* 0 on success, -1 if an event could not be created. */
int timer_init(cuda_timer *t);
/* Record the start event on `stream`. */
void timer_start(cuda_timer *t, void *stream);
/* Record the stop event, wait for it, return elapsed milliseconds
 * (negative on error). */
double timer_stop(cuda_timer *t, void *stream);
void timer_destroy(cuda_timer *t);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_CUDA_TIMER_H */
