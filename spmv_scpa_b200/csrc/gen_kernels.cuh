// gen_kernels.cuh -- 27-point stencil rows generated directly in HBM.
//
// New code (no reference counterpart).  Needed for BASELINE.json configs[4]
// (512^3: 3.6e9 entries, not representable as one int-indexed sparse_csr and
// far too large for a .mtx file).  Produces exactly the arrays
// gen_stencil27_rows() (host/gen.c) produces for the same rows -- checked by
// tests/test_gpu_device.py::test_device_stencil_generator_matches_host through
// spmv_b200_csr_download().
#pragma once

#include "common.cuh"

namespace b200 {

struct StencilGeom {
      int nx, ny, nz;
      long long z0; // first plane of this shard
      long long col_offset;
};

__device__ __forceinline__ int axis_span(int c, int n) { return 1 + (c > 0) + (c < n - 1); }
// entries contributed by coordinates [0, c) on an axis of length n
__device__ __forceinline__ long long axis_prefix(int c, int n) {
      return c == 0 ? 0 : (c >= n ? 3ll * n - 2 : 3ll * c - 1);
}

__device__ __forceinline__ long long stencil_row_offset(const StencilGeom &g, int ix, int iy,
                                                        int iz) {
      const long long sx = 3ll * g.nx - 2, sy = 3ll * g.ny - 2;
      const long long before_planes = (axis_prefix(iz, g.nz) - axis_prefix((int)g.z0, g.nz)) * sy * sx;
      return before_planes +
             (long long)axis_span(iz, g.nz) *
                 (axis_prefix(iy, g.ny) * sx + (long long)axis_span(iy, g.ny) * axis_prefix(ix, g.nx));
}

template <typename OffT>
__global__ void stencil27_fill_kernel(StencilGeom g, long long rows, OffT *__restrict__ irp,
                                      int *__restrict__ ja, double *__restrict__ as) {
      const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
      if (r > rows)
            return;
      const long long plane = (long long)g.nx * g.ny;
      if (r == rows) { // closing offset
            const long long sx = 3ll * g.nx - 2, sy = 3ll * g.ny - 2;
            const int z1 = (int)(g.z0 + rows / plane);
            irp[r] = (OffT)((axis_prefix(z1, g.nz) - axis_prefix((int)g.z0, g.nz)) * sy * sx);
            return;
      }
      const long long grow = g.z0 * plane + r;
      const int ix = (int)(grow % g.nx), iy = (int)(grow / g.nx % g.ny), iz = (int)(grow / plane);
      long long k = stencil_row_offset(g, ix, iy, iz);
      irp[r] = (OffT)k;
      for (int dz = -1; dz <= 1; ++dz) {
            if (iz + dz < 0 || iz + dz >= g.nz)
                  continue;
            for (int dy = -1; dy <= 1; ++dy) {
                  if (iy + dy < 0 || iy + dy >= g.ny)
                        continue;
                  for (int dx = -1; dx <= 1; ++dx) {
                        if (ix + dx < 0 || ix + dx >= g.nx)
                              continue;
                        ja[k] = (int)(grow + dz * plane + dy * g.nx + dx - g.col_offset);
                        as[k] = (dz | dy | dx) ? -1.0 : 26.0;
                        ++k;
                  }
            }
      }
}

} // namespace b200
