// kernels/trsm_blocked.cuh -- batched triangular solves with a k x k lower factor, any k (used for
// k > 32), all four (side, trans) variants and the fused POTRS, ONE launch per call (sm_100a).
//
// Replaces the reference's recursion TRSM -> GEMM(cuBLAS batched) -> TRSM
// (Xtrsm_batch_drivers.cuh:127-266) and the 4 x TRSM + 2 x GEMM composition of POTRS
// (Xpotrs_batch_drivers.cuh:94-171): 30 (strided) to 92 (pointer array) launches at n = 256.
//
// One warp owns one (matrix, 32-vector slab) task; every lane owns one right-hand-side vector
// (side R: a row of B, side L: a column of B) and walks it in blocks of 32 entries held in
// registers.  Blocked substitution over the 32 x 32 blocks L[J][K] of the factor:
//   forward  (R/T, L/N):  x_J = ( a b_J - sum_{K<J} L[J][K] x_K ) solved against L[J][J]
//   backward (R/N, L/T):  x_J = ( a b_J - sum_{K>J} L[K][J]^T x_K ) solved against L[J][J]^T
// The warp stages each L block into shared memory once (coalesced) and reads it back as
// warp-uniform broadcasts; the already solved x_K are re-read from B by the lane that wrote them.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"

namespace kblasx {

// element e of vector v of B:  side R: B[v + e*ldb],  side L: B[e + v*ldb]
template <bool LEFT>
__device__ __forceinline__ long b_index(int v, int e, int ldb) {
  return LEFT ? ((long)e + (long)v * ldb) : ((long)v + (long)e * ldb);
}

// Stage a 32 x 32 tile of the factor in MEMORY order: Ts[i*NB + lane] = A[(r0 + lane) + (c0 + i)*lda]
// (zero outside nr x nc).  32 predicated loads in flight, then 32 conflict-free stores.
template <typename T>
__device__ __forceinline__ void stage_tile(T *Ts, const T *__restrict__ A, int lda, int r0, int c0, int nr, int nc,
                                           int lane) {
  constexpr int NB = 32;
  T v[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    v[i] = T(0);
    ldg_stream_if(v[i], A + (r0 + lane) + (long)(c0 + i) * lda, lane < nr && i < nc);
  }
  sched_fence();
#pragma unroll
  for (int i = 0; i < NB; ++i) Ts[i * NB + lane] = v[i];
  __syncwarp();
}

template <typename T, bool LEFT, bool FORWARD>
__device__ __forceinline__ void tri_blocked_pass(const int k, const T alpha, const T *__restrict__ A, const int lda,
                                                 T *__restrict__ B, const int ldb, const int my, const bool have,
                                                 T *Lkk, T *invd, T *S, const int lane) {
  constexpr int NB = 32;
  typedef typename Vec2T<T>::type V2;
  const int nblk = (k + NB - 1) / NB;
  for (int bi = 0; bi < nblk; ++bi) {
    const int J = FORWARD ? bi : (nblk - 1 - bi);
    const int j0 = J * NB;
    const int jb = (k - j0 < NB) ? (k - j0) : NB;
    T x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      x[c] = T(0);
      ldg_stream_if(x[c], B + b_index<LEFT>(my, j0 + c, ldb), have && c < jb);
    }
    sched_fence();
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] *= alpha;

    // ---- contributions of the blocks already solved -------------------------------------------
    for (int bk = 0; bk < bi; ++bk) {
      const int K = FORWARD ? bk : (nblk - 1 - bk);
      const int k0 = K * NB;
      const int kb = (k - k0 < NB) ? (k - k0) : NB;
      // my already solved entries of block K (negated)
      T nx[NB];
#pragma unroll
      for (int kk = 0; kk < NB; ++kk) {
        nx[kk] = T(0);
        ldg_stream_if(nx[kk], B + b_index<LEFT>(my, k0 + kk, ldb), have && kk < kb);
      }
      __syncwarp();  // previous users of S are done
      if (FORWARD) {
        // coefficient of x_K[kk] in equation c: L[j0 + c][k0 + kk] = Ts[kk*NB + c]  (axpy form)
        stage_tile<T>(S, A, lda, j0, k0, jb, kb, lane);
#pragma unroll 8
        for (int kk = 0; kk < NB; ++kk) {
          const T m1 = -nx[kk];
#pragma unroll
          for (int c = 0; c < NB; c += 2) {
            const V2 s2 = lds_pair(S + kk * NB + c);
            x[c] = fma_t(m1, s2.x, x[c]);
            x[c + 1] = fma_t(m1, s2.y, x[c + 1]);
          }
        }
      } else {
        // coefficient of x_K[kk] in equation c: L[k0 + kk][j0 + c] = Ts[c*NB + kk]  (dot form)
        stage_tile<T>(S, A, lda, k0, j0, kb, jb, lane);
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
          for (int kk = 0; kk < NB; kk += 2) {
            const V2 s2 = lds_pair(S + c * NB + kk);
            acc[kk & 3] = fma_t(nx[kk], s2.x, acc[kk & 3]);
            acc[(kk + 1) & 3] = fma_t(nx[kk + 1], s2.y, acc[(kk + 1) & 3]);
          }
          x[c] -= (acc[0] + acc[1]) + (acc[2] + acc[3]);
        }
      }
    }

    // ---- solve against the diagonal block ------------------------------------------------------
    __syncwarp();
    stage_factor<T, NB>(A + j0 + (long)j0 * lda, lda, jb, Lkk, invd, lane);
    if (FORWARD) tri_forward<T, NB>(x, Lkk, invd);
    else tri_backward<T, NB>(x, Lkk, invd);
#pragma unroll
    for (int c = 0; c < NB; ++c) stg_stream_if(B + b_index<LEFT>(my, j0 + c, ldb), x[c], have && c < jb);
  }
}

// OP: TRI_FORWARD / TRI_BACKWARD / TRI_BOTH (forward with alpha, then backward with 1)
template <typename T, bool LEFT, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
tri_solve_blocked_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                         BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount, const int slabs) {
  constexpr int NB = 32;
  __shared__ __align__(16) T smem[WARPS * (2 * NB * NB + NB)];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *Lkk = smem + warp * (2 * NB * NB + NB);
  T *S = Lkk + NB * NB;
  T *invd = S + NB * NB;

  const long task = (long)blockIdx.x * WARPS + warp;
  if (task >= (long)batchCount * slabs) return;  // warp-uniform
  const long mat = task / slabs;
  const int my = (int)(task % slabs) * 32 + lane;
  const bool have = my < vec;
  const T *__restrict__ A = Aref.at(mat);
  T *__restrict__ B = Bref.at(mat);

  if (OP == TRI_FORWARD || OP == TRI_BOTH)
    tri_blocked_pass<T, LEFT, true>(k, alpha, A, lda, B, ldb, my, have, Lkk, invd, S, lane);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH)
    tri_blocked_pass<T, LEFT, false>(k, OP == TRI_BOTH ? T(1) : alpha, A, lda, B, ldb, my, have, Lkk, invd, S, lane);
}

}  // namespace kblasx
