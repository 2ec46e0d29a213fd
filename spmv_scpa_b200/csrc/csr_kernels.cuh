// csr_kernels.cuh -- FP64 CSR y = A*x kernels for sm_100a.
//
// New designs; the reference kernels they stand in for are
// src/cuda_csr.cu:19-178 (thread / warp / half-warp / block per row, plain
// LDG).  Families here:
//   csr_vec_kernel<LPR>     LPR = 1,2,4,...,32 lanes cooperate on one row
//                           (LPR=1 is thread-per-row, LPR=32 warp-per-row),
//                           shuffle reduction; matrix streams bypass L1.
//   csr_block_row_kernel    one CTA per row, shuffle + shared-memory reduce.
//   csr_split_kernel        rows too long for one CTA are cut into chunks,
//   csr_combine_kernel      partial sums combined deterministically.
//   csr_stream_kernel       persistent CTAs; the value/index streams of a
//                           tile of consecutive rows are staged in shared
//                           memory by cp.async.bulk (TMA) through an
//                           mbarrier ring, then each thread walks its own row
//                           in shared memory, so the x gather of a warp hits
//                           neighbouring columns (ELL-like access on CSR).
// Every kernel takes the matrix as (irp, ja, as) with OffT row offsets
// (int32, or int64 for shards beyond 2^31 entries) and column indices
// already relative to the local x slice.
#pragma once

#include "common.cuh"

namespace b200 {

struct Tuning {
      int stream_hints; // 1: no_allocate + evict_first on matrix streams
      int x_evict_last; // 1: evict_last policy on x gathers
};

// ------------------------------------------------------------------------
// LPR lanes per row.  `rowlist` == nullptr: rows [row0, row0+nrows) in order;
// otherwise rows rowlist[0..nrows).  Rows longer than `max_len` are skipped
// (they belong to another bin's launch); max_len < 0 disables the test.
// ------------------------------------------------------------------------
template <int LPR, typename OffT, bool HINTS>
__global__ void __launch_bounds__(1024)
    csr_vec_kernel(const OffT *__restrict__ irp, const int *__restrict__ ja,
                   const double *__restrict__ as, long long row0, long long nrows,
                   const int *__restrict__ rowlist, long long max_len,
                   const double *__restrict__ x, double *__restrict__ y, PushArgs push) {
      const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
      const long long g = gid / LPR;
      const int sub = (int)(gid % LPR);
      // whole groups leave together, so the shuffles below stay convergent
      if (g >= nrows)
            return;
      const long long row = rowlist ? (long long)rowlist[g] : row0 + g;
      const OffT s = irp[row], e = irp[row + 1];
      if (max_len >= 0 && (long long)(e - s) > max_len)
            return;

      const uint64_t pol_s = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      double acc = 0.0;
#pragma unroll 4
      for (OffT k = s + sub; k < e; k += LPR) {
            double a;
            int c;
            if (HINTS) {
                  a = ld_stream_f64(as + k, pol_s);
                  c = ld_stream_s32(ja + k, pol_s);
            } else {
                  a = __ldg(as + k);
                  c = __ldg(ja + k);
            }
            acc = fma(a, ld_x(x + c, pol_x), acc);
      }
      acc = group_sum<LPR>(acc);
      if (sub == 0)
            store_y(y, row, acc, push);
}

// ------------------------------------------------------------------------
// One CTA per row (blockDim.x = 32*wpb).  Row = rowlist[blockIdx.x] or
// row0 + blockIdx.x.
// ------------------------------------------------------------------------
template <typename OffT>
__global__ void __launch_bounds__(1024)
    csr_block_row_kernel(const OffT *__restrict__ irp, const int *__restrict__ ja,
                         const double *__restrict__ as, long long row0,
                         const int *__restrict__ rowlist, const double *__restrict__ x,
                         double *__restrict__ y, PushArgs push) {
      __shared__ double warp_part[32];
      const long long row = rowlist ? (long long)rowlist[blockIdx.x] : row0 + blockIdx.x;
      const OffT s = irp[row], e = irp[row + 1];
      const uint64_t pol_s = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();

      double acc = 0.0;
#pragma unroll 4
      for (OffT k = s + threadIdx.x; k < e; k += blockDim.x)
            acc = fma(ld_stream_f64(as + k, pol_s), ld_x(x + ld_stream_s32(ja + k, pol_s), pol_x),
                      acc);
      acc = group_sum<32>(acc);

      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int nwarps = blockDim.x >> 5;
      if (lane == 0)
            warp_part[warp] = acc;
      __syncthreads();
      if (warp == 0) {
            double v = lane < nwarps ? warp_part[lane] : 0.0;
            v = group_sum<32>(v);
            if (lane == 0)
                  store_y(y, row, v, push);
      }
}

// ------------------------------------------------------------------------
// Very long rows: chunk c covers entries [chunk_k[c], chunk_k[c+1]) of one
// row and produces partial[c]; csr_combine_kernel then adds the partials of
// each split row in chunk order (deterministic, no atomics).
// ------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
    csr_split_kernel(const long long *__restrict__ chunk_k0, const long long *__restrict__ chunk_k1,
                     const int *__restrict__ ja, const double *__restrict__ as,
                     const double *__restrict__ x, double *__restrict__ partial) {
      __shared__ double warp_part[32];
      const long long s = chunk_k0[blockIdx.x], e = chunk_k1[blockIdx.x];
      const uint64_t pol_s = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      double acc = 0.0;
#pragma unroll 4
      for (long long k = s + threadIdx.x; k < e; k += blockDim.x)
            acc = fma(ld_stream_f64(as + k, pol_s), ld_x(x + ld_stream_s32(ja + k, pol_s), pol_x),
                      acc);
      acc = group_sum<32>(acc);
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      if (lane == 0)
            warp_part[warp] = acc;
      __syncthreads();
      if (warp == 0) {
            double v = lane < (int)(blockDim.x >> 5) ? warp_part[lane] : 0.0;
            v = group_sum<32>(v);
            if (lane == 0)
                  partial[blockIdx.x] = v;
      }
}

__global__ void csr_combine_kernel(const int *__restrict__ split_row,
                                   const int *__restrict__ split_first_chunk, int n_split,
                                   const double *__restrict__ partial, double *__restrict__ y,
                                   PushArgs push) {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n_split)
            return;
      double acc = 0.0;
      for (int c = split_first_chunk[i]; c < split_first_chunk[i + 1]; ++c)
            acc += partial[c];
      store_y(y, split_row[i], acc, push);
}

// ------------------------------------------------------------------------
// TMA-staged row tiles.
//
// Plan (host): tile t = rows [tile_row[t], tile_row[t+1]) with at most
// MAXROWS rows and (after rounding the entry range out to multiples of 4)
// at most CAP entries; tile_k[t] = irp[tile_row[t]].  A tile that is a single
// row with more than CAP entries is skipped here (long-row launch).
//
// Kernel: persistent CTAs, CTA b owns tiles b, b+G, b+2G, ...  A ring of
// STAGES shared-memory buffers; thread 0 arms stage s with expect_tx and
// issues two cp.async.bulk copies (values, indices); everybody waits on the
// stage's mbarrier, computes from shared memory, __syncthreads, and thread 0
// refills the stage with the tile STAGES ahead.
//
// Compute: LPR lanes per row (LPR = 1: thread per row).  Rows of a warp are
// neighbours, so for banded matrices the 32 x-gathers of one step fall in a
// few contiguous sectors.  For even row lengths the walk through the row is
// rotated by the row's index so lanes start in different banks.
// ------------------------------------------------------------------------
template <int THREADS, int LPR, int STAGES, int CAP, int PASSES, typename OffT>
struct StreamCfg {
      static constexpr int kThreads = THREADS;
      static constexpr int kRowsPerPass = THREADS / LPR;
      static constexpr int kMaxRows = kRowsPerPass * PASSES;
      static constexpr int kCap = CAP; // entries per stage (multiple of 4)
      static constexpr size_t kSmem = (size_t)STAGES * CAP * 12 + STAGES * 8 + 16;
};

template <int THREADS, int LPR, int STAGES, int CAP, int PASSES, typename OffT>
__global__ void __launch_bounds__(THREADS)
    csr_stream_kernel(const OffT *__restrict__ irp, const int *__restrict__ ja,
                      const double *__restrict__ as, const int *__restrict__ tile_row,
                      const long long *__restrict__ tile_k, int tile0, int n_tiles,
                      const double *__restrict__ x, double *__restrict__ y, PushArgs push) {
      extern __shared__ __align__(128) unsigned char smem_raw[];
      double *s_as = reinterpret_cast<double *>(smem_raw);
      int *s_ja = reinterpret_cast<int *>(smem_raw + (size_t)STAGES * CAP * 8);
      uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * CAP * 12);

      constexpr int RPP = THREADS / LPR;
      const int tid = threadIdx.x;
      const int gi = tid / LPR, sub = tid % LPR;
      const uint64_t pol_s = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();

      if (tid == 0) {
            for (int s = 0; s < STAGES; ++s)
                  mbar_init(&bars[s], 1);
            mbar_fence_init();
      }
      __syncthreads();

      // number of tiles this CTA owns
      const int first = tile0 + blockIdx.x;
      const int last = tile0 + n_tiles;
      const int stride = gridDim.x;

      auto issue = [&](int t, int stage) {
            const long long k0 = tile_k[t] & ~3ll;
            const long long k1 = (tile_k[t + 1] + 3) & ~3ll;
            const long long cnt = k1 - k0;
            if (cnt > 0 && cnt <= CAP) {
                  mbar_expect_tx(&bars[stage], (uint32_t)(cnt * 12));
                  bulk_g2s(s_as + (size_t)stage * CAP, as + k0, (uint32_t)(cnt * 8), &bars[stage],
                           pol_s);
                  bulk_g2s(s_ja + (size_t)stage * CAP, ja + k0, (uint32_t)(cnt * 4), &bars[stage],
                           pol_s);
            } else {
                  // nothing to copy (empty rows only, or a long row handled
                  // elsewhere): complete the phase with a zero-byte arrival
                  mbar_expect_tx(&bars[stage], 0);
            }
      };

      if (tid == 0) {
            int t = first;
            for (int s = 0; s < STAGES && t < last; ++s, t += stride)
                  issue(t, s);
      }

      int it = 0;
      for (int t = first; t < last; t += stride, ++it) {
            const int stage = it % STAGES;
            const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
            const int r0 = tile_row[t], r1 = tile_row[t + 1];
            const long long kbase = tile_k[t] & ~3ll;
            const bool staged = ((tile_k[t + 1] + 3) & ~3ll) - kbase <= CAP;

            // row extent of pass 0 is fetched before waiting on the copy
            int row = r0 + gi;
            long long ks = 0, ke = 0;
            if (row < r1) {
                  ks = (long long)irp[row] - kbase;
                  ke = (long long)irp[row + 1] - kbase;
            }
            mbar_wait(&bars[stage], parity);

            if (staged) {
                  const double *tas = s_as + (size_t)stage * CAP;
                  const int *tja = s_ja + (size_t)stage * CAP;
#pragma unroll 1
                  for (int p = 0; p < PASSES; ++p) {
                        if (p > 0) {
                              if (r0 + p * RPP >= r1)
                                    break; // uniform over the CTA
                              row = r0 + p * RPP + gi;
                              if (row < r1) {
                                    ks = (long long)irp[row] - kbase;
                                    ke = (long long)irp[row + 1] - kbase;
                              }
                        }
                        const int len = row < r1 ? (int)(ke - ks) : 0;
                        const int base = (int)ks;
                        // rotate the walk for even lengths (bank spreading)
                        int start = 0;
                        if (LPR == 1 && len > 1 && (len & 1) == 0)
                              start = gi % len;
                        double acc = 0.0;
#pragma unroll 4
                        for (int j = sub; j < len; j += LPR) {
                              int jj = j + start;
                              jj = jj >= len ? jj - len : jj;
                              const double a = tas[base + jj];
                              const int c = tja[base + jj];
                              acc = fma(a, ld_x(x + c, pol_x), acc);
                        }
                        if (LPR > 1)
                              acc = group_sum<LPR>(acc);
                        if (sub == 0 && row < r1)
                              store_y(y, row, acc, push);
                  }
            }
            __syncthreads();
            if (tid == 0) {
                  const int tn = t + STAGES * stride;
                  if (tn < last)
                        issue(tn, stage);
            }
      }
}

} // namespace b200
