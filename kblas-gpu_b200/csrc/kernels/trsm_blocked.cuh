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
                                           int lane, int flags = 0) {
  constexpr int NB = 32;
  T v[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    v[i] = T(0);
    // upper storage: L[r0 + lane][c0 + i] = U[c0 + i][r0 + lane]
    const T *p = (flags & TRI_FLAG_UPPER) ? A + (c0 + i) + (long)(r0 + lane) * lda : A + (r0 + lane) + (long)(c0 + i) * lda;
    ldg_stream_if(v[i], p, lane < nr && i < nc);
  }
  sched_fence();
#pragma unroll
  for (int i = 0; i < NB; ++i) Ts[i * NB + lane] = v[i];
  __syncwarp();
}

// L2 prefetch of the 32 x 32 block at (r0, c0) of a k x k factor (hint only; coordinates clamped into the matrix)
template <typename T>
__device__ __forceinline__ void prefetch_tile_l2(const T *__restrict__ A, int lda, int k, int r0, int c0, int lane) {
  constexpr int LPC = (32 * (int)sizeof(T)) / 128 > 0 ? (32 * (int)sizeof(T)) / 128 : 1;  // 128-byte lines per column segment
#pragma unroll
  for (int l = lane; l < 32 * LPC; l += 32) {
    int r = r0 + (l % LPC) * (128 / (int)sizeof(T)), c = c0 + l / LPC;
    r = r < k ? r : k - 1;
    c = c < k ? c : k - 1;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + r + (long)c * lda));
  }
}

// GP = lanes per matrix: 32 (one matrix per warp, vec may span several 32-vector slabs) or 16 / 8
// (vec <= GP right-hand sides: 2 / 4 matrices per warp, each lane group with its own staged tiles --
// posv with 16 right-hand-side rows otherwise leaves half of every warp idle).
template <typename T, bool LEFT, bool FORWARD, int GP>
__device__ __forceinline__ void tri_blocked_pass(const int k, const T alpha, const T *const (&Aq)[32 / GP], const int lda,
                                                 T *__restrict__ B, const int ldb, const int my, const bool have,
                                                 T *Lkk_all, T *S_all, const int lane, const int flags) {
  constexpr int NB = 32;
  constexpr int MPW = 32 / GP;
  constexpr int FSZ = NB * NB + NB;
  typedef typename Vec2T<T>::type V2;
  const int g = lane / GP;
  T *Lkk = Lkk_all + g * FSZ, *invd = Lkk + NB * NB, *S = S_all + g * FSZ;
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
      // my already solved entries of block K
      T nx[NB];
#pragma unroll
      for (int kk = 0; kk < NB; ++kk) {
        nx[kk] = T(0);
        ldg_stream_if(nx[kk], B + b_index<LEFT>(my, k0 + kk, ldb), have && kk < kb);
      }
      __syncwarp();  // previous users of S are done
      {
        // pull the block that is needed next into L2 while this one is staged and used (the solve is a chain of
        // dependent block steps: without the hint each one starts with a DRAM / far-L2 round trip)
        const bool more = bk + 1 < bi;
        const int kn = FORWARD ? (K + 1) * NB : (K - 1) * NB;
        const int pr = FORWARD ? j0 : (more ? kn : j0), pc = FORWARD ? (more ? kn : j0) : j0;
#pragma unroll
        for (int q = 0; q < MPW; ++q) prefetch_tile_l2<T>(Aq[q], lda, k, pr, pc, lane);
      }
      if (FORWARD) {
        // coefficient of x_K[kk] in equation c: L[j0 + c][k0 + kk] = Ts[kk*NB + c]  (axpy form)
#pragma unroll
        for (int q = 0; q < MPW; ++q) stage_tile<T>(S_all + q * FSZ, Aq[q], lda, j0, k0, jb, kb, lane, flags);
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
#pragma unroll
        for (int q = 0; q < MPW; ++q) stage_tile<T>(S_all + q * FSZ, Aq[q], lda, k0, j0, kb, jb, lane, flags);
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
#pragma unroll
    for (int q = 0; q < MPW; ++q)
      stage_factor<T, NB>(Aq[q] + j0 + (long)j0 * lda, lda, jb, Lkk_all + q * FSZ, Lkk_all + q * FSZ + NB * NB, lane, flags);
    if (FORWARD) tri_forward<T, NB>(x, Lkk, invd);
    else tri_backward<T, NB>(x, Lkk, invd);
#pragma unroll
    for (int c = 0; c < NB; ++c) stg_stream_if(B + b_index<LEFT>(my, j0 + c, ldb), x[c], have && c < jb);
  }
}

template <typename T, int GP>
struct TriBlockedSmem {
  static constexpr int NB = 32;
  // the update tiles (S) and the diagonal block (Lkk + invd) are never live at the same time: one
  // region serves both, which doubles the resident warps (shared memory is the occupancy limiter)
  static constexpr int per_warp = (32 / GP) * (NB * NB + NB);  // elements of T
};

// OP: TRI_FORWARD / TRI_BACKWARD / TRI_BOTH (forward with alpha, then backward with 1)
template <typename T, bool LEFT, int OP, int GP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
tri_solve_blocked_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                         BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount, const int slabs, const int flags) {
  constexpr int MPW = 32 / GP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *Lkk_all = reinterpret_cast<T *>(smem_raw) + warp * TriBlockedSmem<T, GP>::per_warp;
  T *S_all = Lkk_all;  // aliased (see TriBlockedSmem); tile q lives at S_all + q*FSZ

  // task = (warp-batch of MPW matrices, slab); slabs == 1 when packed
  const long task = (long)blockIdx.x * WARPS + warp;
  const long wbatches = ((long)batchCount + MPW - 1) / MPW;
  if (task >= wbatches * slabs) return;  // warp-uniform
  const long mat0 = (task / slabs) * MPW;
  const int g = lane / GP;
  const int my = (int)(task % slabs) * 32 + (lane % GP);
  const long last = (long)batchCount - 1;
  const T *Aq[MPW];
#pragma unroll
  for (int q = 0; q < MPW; ++q) Aq[q] = Aref.at(mat0 + q < last ? mat0 + q : last);
  const bool have = (mat0 + g <= last) && (my < vec);
  T *__restrict__ B = Bref.at(mat0 + g < last ? mat0 + g : last);

  if (OP == TRI_FORWARD || OP == TRI_BOTH)
    tri_blocked_pass<T, LEFT, true, GP>(k, alpha, Aq, lda, B, ldb, my, have, Lkk_all, S_all, lane, flags);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH)
    tri_blocked_pass<T, LEFT, false, GP>(k, OP == TRI_BOTH ? T(1) : alpha, Aq, lda, B, ldb, my, have, Lkk_all, S_all, lane, flags);
}

}  // namespace kblasx
