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
    for (int c = 0; c < NB; ++c) x[c] = (have && c < jb) ? alpha * B[b_index<LEFT>(my, j0 + c, ldb)] : T(0);

    // ---- contributions of the blocks already solved -------------------------------------------
    for (int bk = 0; bk < bi; ++bk) {
      const int K = FORWARD ? bk : (nblk - 1 - bk);
      const int k0 = K * NB;
      const int kb = (k - k0 < NB) ? (k - k0) : NB;
      __syncwarp();
      // S[kk*NB + c] = coefficient of x_K[kk] in equation c of block J:
      //   forward: L[j0 + c][k0 + kk]     backward: L[k0 + kk][j0 + c]
      for (int e = lane; e < NB * NB; e += 32) {
        int c, kk;
        long src;
        if (FORWARD) { c = e % NB; kk = e / NB; src = (long)(j0 + c) + (long)(k0 + kk) * lda; }
        else         { kk = e % NB; c = e / NB; src = (long)(k0 + kk) + (long)(j0 + c) * lda; }
        S[kk * NB + c] = (c < jb && kk < kb) ? A[src] : T(0);
      }
      __syncwarp();
#pragma unroll 4
      for (int kk = 0; kk < NB; ++kk) {
        const T nx = (have && kk < kb) ? -B[b_index<LEFT>(my, k0 + kk, ldb)] : T(0);
#pragma unroll
        for (int c = 0; c < NB; c += 2) {
          const V2 s2 = lds_pair(S + kk * NB + c);
          x[c] = fma_t(nx, s2.x, x[c]);
          x[c + 1] = fma_t(nx, s2.y, x[c + 1]);
        }
      }
    }

    // ---- solve against the diagonal block ------------------------------------------------------
    __syncwarp();
    stage_factor<T, NB>(A + j0 + (long)j0 * lda, lda, jb, Lkk, invd, lane);
    if (FORWARD) tri_forward<T, NB>(x, Lkk, invd);
    else tri_backward<T, NB>(x, Lkk, invd);
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (have && c < jb) B[b_index<LEFT>(my, j0 + c, ldb)] = x[c];
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
