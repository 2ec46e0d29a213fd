/* mmio.h -- the slice of the NIST Matrix Market I/O API the loader uses.
 *
 * The reference vendors the whole NIST library (include/mmio.h, src/mmio.c)
 * but calls only mm_read_banner, mm_read_mtx_crd_size and the mm_is_*
 * predicates (src/csr.c:48-59).  This is a fresh implementation of that
 * slice with the same names, typecode encoding (4 chars: object, format,
 * field, symmetry) and return codes, so code written against mmio.h builds.
 */
#ifndef SPMV_B200_MMIO_H
#define SPMV_B200_MMIO_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_MAX_LINE_LENGTH 1025
#define MM_MAX_TOKEN_LENGTH 64
#define MatrixMarketBanner "%%MatrixMarket"

typedef char MM_typecode[4];

/* return codes */
#define MM_COULD_NOT_READ_FILE 11
#define MM_PREMATURE_EOF 12
#define MM_NOT_MTX 13
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15
#define MM_LINE_TOO_LONG 16
#define MM_COULD_NOT_WRITE_FILE 17

/* typecode[0]: 'M' matrix */
#define mm_is_matrix(t) ((t)[0] == 'M')
/* typecode[1]: 'C' coordinate (sparse) | 'A' array (dense) */
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
/* typecode[2]: 'R' real | 'C' complex | 'P' pattern | 'I' integer */
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
/* typecode[3]: 'G' general | 'S' symmetric | 'K' skew | 'H' hermitian */
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

#define mm_clear_typecode(t)                                                   \
      ((*(t))[0] = (*(t))[1] = (*(t))[2] = ' ', (*(t))[3] = 'G')

/* Parse the "%%MatrixMarket matrix <format> <field> <symmetry>" line. */
int mm_read_banner(FILE *f, MM_typecode *matcode);
/* Skip comment lines, then read "M N nz". */
int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz);
/* Writers used by the synthetic generators. */
int mm_write_banner(FILE *f, MM_typecode matcode);
int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_MMIO_H */
