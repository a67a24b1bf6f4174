// kernels/potrf_panel.cuh -- batched Cholesky for n > 32: one CTA per matrix, left-looking over
// 32-column panels, ONE launch per call (sm_100a).
//
// Replaces the reference's host recursion for n > 16 (Xpotrf_batch_drivers.cuh:91-135: potrf ->
// trsm -> syrk -> potrf, each level more launches; 108 launches and 28 cuBLAS batched GEMMs at
// n = 256) and its per-level global-memory round trips.
//
// Mapping.  Every thread owns R = 2 rows of the current panel and keeps them in registers
// (2 x 32 values).  For panel J (columns j0 .. j0+31, rows j0 .. n-1):
//   1. the thread loads its rows of the panel (coalesced: consecutive threads = consecutive rows);
//   2. left-looking update  P -= L[rows, 0:j0] * L[j0:j0+32, 0:j0]^T : the thread streams its own
//      row of the already factored columns from global memory (L2 resident, coalesced) and gets
//      the 32 x 32 tile of L[j0:j0+32, k0:k0+32] from shared memory as warp-uniform LDS.128
//      broadcasts: 64 FMAs per thread for every 16 shared-memory loads, no cross-lane traffic;
//   3. warp 0 factors the 32 x 32 diagonal block (row per lane, right-looking, column broadcast
//      through shared memory) and leaves L_JJ + reciprocal diagonal in shared memory;
//   4. every other row does its forward substitution against L_JJ (kernels/trsm_small.cuh);
//   5. rows are stored (only row >= col: the strict upper triangle is never written).
// Matrices taller than THREADS*R rows are processed in row slabs that reuse L_JJ, so any n works.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"

namespace kblasx {

template <typename T, int THREADS, int R, bool STRIDED>
__global__ void __launch_bounds__(THREADS)
potrf_panel_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount, int *__restrict__ info,
                   const int info_mode) {
  constexpr int NB = 32;             // panel width; R = rows per thread
  constexpr int SLAB = THREADS * R;  // rows per slab
  typedef typename Vec2T<T>::type V2;

  __shared__ __align__(16) T Lkk[NB * NB];  // factored diagonal block, column-major, identity padded
  __shared__ __align__(16) T invd[NB];      // 1 / diag(L_JJ)
  __shared__ __align__(16) T S[NB * NB];    // S[kk*NB + c] = L[j0 + c][k0 + kk]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  T *__restrict__ A = Aref.at(blockIdx.x);
  int bad = 0;

  for (int j0 = 0; j0 < n; j0 += NB) {
    const int jb = (n - j0 < NB) ? (n - j0) : NB;
    const int m = n - j0;  // rows of this panel
    for (int r0 = 0; r0 < m; r0 += SLAB) {
      int row[R];
      bool valid[R];
      T p[R][NB];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        row[q] = j0 + r0 + tid + q * THREADS;
        valid[q] = row[q] < n;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          p[q][c] = T(0);
          ldg_stream_if(p[q][c], A + row[q] + (long)(j0 + c) * lda, valid[q] && c < jb);  // branch-free
        }
      }

      // ---- 2. left-looking update with every previously factored block column ---------------
      for (int k0 = 0; k0 < j0; k0 += NB) {
        __syncthreads();  // S is about to be overwritten
        {
          // all loads of the tile in flight, then the (conflict-free) stores
          constexpr int PER = NB * NB / THREADS;
          T tv[PER];
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int e = tid + i * THREADS, c = e % NB, kk = e / NB;
            tv[i] = T(0);
            ldg_stream_if(tv[i], A + (j0 + c) + (long)(k0 + kk) * lda, c < jb);
          }
          sched_fence();
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int e = tid + i * THREADS;
            S[(e / NB) * NB + (e % NB)] = tv[i];
          }
        }
        __syncthreads();
#pragma unroll
        for (int kc = 0; kc < NB; kc += 8) {
          // my rows' entries of 8 already factored columns at a time (loads batched ahead of the FMAs)
          T na[R][8];
#pragma unroll
          for (int q = 0; q < R; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              na[q][i] = T(0);
              ldg_stream_if(na[q][i], A + row[q] + (long)(k0 + kc + i) * lda, valid[q]);
            }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int c = 0; c < NB; c += 2) {
              const V2 s2 = lds_pair(S + (kc + i) * NB + c);
#pragma unroll
              for (int q = 0; q < R; ++q) {
                p[q][c] = fma_t(-na[q][i], s2.x, p[q][c]);
                p[q][c + 1] = fma_t(-na[q][i], s2.y, p[q][c + 1]);
              }
            }
          }
        }
      }

      // ---- 3. diagonal block: rows j0 .. j0+31 live in p[0] of warp 0 (first slab only) -------
      if (r0 == 0) {
        if (warp == 0) {
          if (lane >= jb) {  // identity padding of a ragged last panel
#pragma unroll
            for (int c = 0; c < NB; ++c) p[0][c] = (c == lane) ? T(1) : T(0);
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const T d = shfl_seg<32>(p[0][j], j);
            if (info_mode && bad == 0 && j < jb && !(d > T(0))) bad = j0 + j + 1;
            const T r = rsqrt_t(d);
            p[0][j] *= r;
            Lkk[lane + j * NB] = p[0][j];
            if (lane == j) invd[j] = r;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NB; ++k) {
              // warp-uniform broadcast of L[k][j]; rows above the diagonal pick up garbage that
              // is never read or stored
              p[0][k] = fma_t(-p[0][j], lds_one(Lkk + k + j * NB), p[0][k]);
            }
          }
        }
        __syncthreads();  // L_JJ and invd are published
      }

      // ---- 4. forward substitution of the rows below the diagonal block -----------------------
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const bool is_diag_row = (r0 == 0) && (warp == 0) && (q == 0);
        if (!is_diag_row) tri_forward<T, NB>(p[q], Lkk, invd);
      }

      // ---- 5. store --------------------------------------------------------------------------
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          stg_stream_if(A + row[q] + (long)(j0 + c) * lda, p[q][c], valid[q] && c < jb && row[q] >= j0 + c);
        }
      }
      // rows of later slabs reuse Lkk; the next panel's S / Lkk writes are fenced by the
      // __syncthreads at the top of the k0 loop and below
    }
    // the factored panel must be visible to the whole CTA before panel J+1 reads it from global
    __threadfence_block();
    __syncthreads();
  }
  if (info_mode && tid == 0) info[blockIdx.x] = bad;
}

}  // namespace kblasx
