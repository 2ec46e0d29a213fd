/* err.h -- kept for source compatibility with code written against the reference's
 * include/err.h (ERR_PTR / PTR_ERR / IS_ERR / LOG_ERR); everything lives in spmv_errptr.h. */
#ifndef SPMV_B200_ERR_H
#define SPMV_B200_ERR_H
#include "spmv_errptr.h"
#endif
