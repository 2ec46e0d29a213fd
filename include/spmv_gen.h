/* spmv_gen.h -- deterministic synthetic matrices (no reference counterpart).
 *
 * The reference benchmarks SuiteSparse downloads (scripts/download-matrices.py);
 * there is no network here, and BASELINE.json names synthetic inputs.  These
 * generators build a sparse_csr directly in host memory with exactly the
 * arrays io_load_csr would produce from the equivalent .mtx file written by
 * gen_write_mtx (row by row, entries in increasing column order unless noted)
 * -- tests/test_generators.py checks that equivalence on small sizes.
 * All results are released with csr_free().  NULL on failure (ENOMEM or a
 * matrix that does not fit int32 indices).
 */
#ifndef SPMV_B200_GEN_H
#define SPMV_B200_GEN_H

#include <stdint.h>

#include "csr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 2D 5-point Laplacian on nx*ny points, x fastest: diag 4, neighbours -1. */
sparse_csr *gen_poisson2d(int nx, int ny);
/* 3D 27-point stencil on nx*ny*nz points, x fastest: diag 26, others -1. */
sparse_csr *gen_stencil27(int nx, int ny, int nz);
/* Rows [row0,row1) of the same stencil, column indices left global. */
sparse_csr *gen_stencil27_rows(int nx, int ny, int nz, int64_t row0,
                               int64_t row1);
/* n*n, k distinct uniformly random columns per row (sorted), values uniform
 * in (-1,1); counter-based generator keyed by (seed,row): any row range can be
 * regenerated independently. */
sparse_csr *gen_uniform_random(int n, int k, uint64_t seed);
/* R-MAT / Kronecker graph: 2^scale vertices, edge_factor*2^scale directed
 * edges, quadrant probabilities (a,b,c,1-a-b-c); duplicates are KEPT (as
 * io_load_csr keeps them) and a row lists its entries in edge-generation
 * order; values uniform in (-1,1). */
sparse_csr *gen_rmat(int scale, int edge_factor, double a, double b, double c,
                     uint64_t seed);
/* Banded test matrix with ragged rows: row r has (r*7919 % (max_len+1))
 * entries at columns r-len/2.. clipped to [0,n), values from the counter
 * generator.  Used by the parity tests to cover every row-length bin. */
sparse_csr *gen_ragged(int n, int max_len, uint64_t seed);

/* Write A as "matrix coordinate real general", one entry per line in CSR
 * order, values with 17 significant digits (round-trips FP64 exactly).
 * 0 or -errno. */
int gen_write_mtx(const sparse_csr *A, const char *path);

/* The 64-bit mixing function all generators share (splitmix64 finaliser). */
uint64_t gen_mix64(uint64_t z);

#ifdef __cplusplus
}
#endif

#endif /* SPMV_B200_GEN_H */
