// kernels/potrf_panel_dmma.cuh -- fp64 batched Cholesky for n > 32 with the left-looking trailing
// update on the FP64 tensor path (mma.sync m8n8k4 = DMMA), one CTA per matrix (sm_100a).
//
// Same panel algorithm as kernels/potrf_panel.cuh (which stays the fp32 path): for the panel of
// columns j0 .. j0+31 every warp owns a 32-row x 32-column block of the panel.
//   1. update: acc(32x32) = L[rows, 0:j0] * L[j0:j0+32, 0:j0]^T as 4x4 DMMA tiles, 16 independent
//      accumulators per warp (hides the ~150-cycle DMMA latency), operands fetched straight from
//      global/L2 in fragment order: both the A fragment (rows of this warp) and the B fragment
//      (rows j0.. of the panel, shared by all warps through L1) are "element (base + lane/4,
//      k + lane%4)" of the same column-major matrix.  No shared-memory broadcast traffic at all:
//      the FFMA/DFMA version needs 16 LDS.128 per 64 FMAs and is MIO-bound.
//   2. the accumulators go through (padded) shared memory into a row-per-thread layout,
//      p = A[row, j0:j0+32] - acc;
//   3. warp 0 factors the diagonal block (row per lane), 4. the other rows solve against it,
//   5. rows are stored -- identical to the fp32 kernel.
// Measured on B200: DMMA peaks at 63.5 FMA/clk/SM, the same as the DFMA pipe
// (profiles/r01_microbench_pipes.txt); it wins by freeing issue slots and the MIO pipe.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"

namespace kblasx {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int THREADS>
struct PanelDmmaSmem {
  static constexpr int NB = 32;
  static constexpr int LD = 33;  // odd row stride: conflict-free row-per-thread reads
  static constexpr int warps = THREADS / 32;
  static constexpr size_t bytes = sizeof(double) * (NB * NB + NB + (size_t)warps * NB * LD);
};

#ifndef KX_PANEL_WARPS_PER_SM
#define KX_PANEL_WARPS_PER_SM 8
#endif
template <int THREADS, bool STRIDED>
__global__ void __launch_bounds__(THREADS, (32 * KX_PANEL_WARPS_PER_SM) / THREADS)  // resident warps per SM
potrf_panel_dmma_kernel(const int n, BatchRef<double, STRIDED> Aref, const int lda, const int batchCount,
                        int *__restrict__ info, const int info_mode) {
  typedef double T;
  constexpr int NB = 32;
  constexpr int LD = PanelDmmaSmem<THREADS>::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *Lkk = reinterpret_cast<T *>(smem_raw);  // factored diagonal block, column-major, identity padded
  T *invd = Lkk + NB * NB;                    // 1 / diag(L_JJ)
  T *accs = invd + NB;                        // per warp: 32 x LD transpose buffer

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane
  T *__restrict__ A = Aref.at(blockIdx.x);
  T *acc_w = accs + warp * NB * LD;
  int bad = 0;

  for (int j0 = 0; j0 < n; j0 += NB) {
    const int jb = (n - j0 < NB) ? (n - j0) : NB;
    const int m = n - j0;  // rows of this panel
    for (int r0 = 0; r0 < m; r0 += THREADS) {
      const int wrow0 = j0 + r0 + warp * 32;  // first row of this warp's block
      const bool warp_has_rows = wrow0 < n;   // warp-uniform
      const int row = wrow0 + lane;
      const bool valid = row < n;
      T p[NB];

      // ---- 1. acc = L[rows, 0:j0] * L[j0:j0+32, 0:j0]^T on the FP64 tensor path ---------------
      if (warp_has_rows && j0 > 0) {
        T acc[4][4][2];
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
        // rows beyond n read row n-1 instead (finite data, results discarded)
        int arow[4], brow[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          arow[b] = wrow0 + 8 * b + fr;
          arow[b] = arow[b] < n ? arow[b] : n - 1;
          brow[b] = j0 + 8 * b + fr;
          brow[b] = brow[b] < n ? brow[b] : n - 1;
        }
        const T *col = A + (long)fk * lda;
#pragma unroll 4
        for (int k = 0; k < j0; k += 4) {
          T af[4], bf[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            af[b] = col[arow[b]];
            bf[b] = col[brow[b]];
          }
#pragma unroll
          for (int rb = 0; rb < 4; ++rb)
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af[rb], bf[cb]);
          col += 4 * (long)lda;
        }
        // C fragment: lane holds (row fr, cols 2*fk, 2*fk+1) of each 8x8 tile
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk] = acc[rb][cb][0];
            acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk + 1] = acc[rb][cb][1];
          }
      }
      __syncwarp();

      // ---- 2. my row of the panel: p = A[row, j0 : j0+32] - acc -------------------------------
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        p[c] = 0.0;
        ldg_stream_if(p[c], A + row + (long)(j0 + c) * lda, valid && c < jb);
      }
      if (j0 > 0) {
#pragma unroll
        for (int c = 0; c < NB; ++c) p[c] -= acc_w[lane * LD + c];
      }
      __syncwarp();

      // ---- 3. diagonal block: rows j0 .. j0+31 are warp 0's rows in the first slab -------------
      if (r0 == 0) {
        if (warp == 0) {
          if (lane >= jb) {  // identity padding of a ragged last panel
#pragma unroll
            for (int c = 0; c < NB; ++c) p[c] = (c == lane) ? 1.0 : 0.0;
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const T d = shfl_seg<32>(p[j], j);
            if (info_mode && bad == 0 && j < jb && !(d > 0.0)) bad = j0 + j + 1;
            const T r = rsqrt(d);
            p[j] *= r;
            Lkk[lane + j * NB] = p[j];
            if (lane == j) invd[j] = r;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NB; ++k) p[k] = fma(-p[j], lds_one(Lkk + k + j * NB), p[k]);
          }
        }
        __syncthreads();  // L_JJ and invd are published
      }

      // ---- 4. forward substitution of the rows below the diagonal block -----------------------
      if (!(r0 == 0 && warp == 0)) tri_forward<T, NB>(p, Lkk, invd);

      // ---- 5. store --------------------------------------------------------------------------
#pragma unroll
      for (int c = 0; c < NB; ++c)
        stg_stream_if(A + row + (long)(j0 + c) * lda, p[c], valid && c < jb && row >= j0 + c);
    }
    // the factored panel must be visible to the whole CTA before panel J+1 reads it from global
    __threadfence_block();
    __syncthreads();
  }
  if (info_mode && tid == 0) info[blockIdx.x] = bad;
}

}  // namespace kblasx
