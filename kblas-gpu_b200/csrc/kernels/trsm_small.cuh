// kernels/trsm_small.cuh -- batched triangular solves with a k x k lower factor, k <= 32.
//
// Replaces the reference's register kernels K5-K9 (Xtrsm_batch_kernels.cuh:36-898), its
// host recursion with cuBLAS batched GEMM for k in 17..32 (Xtrsm_batch_drivers.cuh:127-266)
// and, for POTRS, the 4 x TRSM + 2 x GEMM composition (Xpotrs_batch_drivers.cuh:94-171)
// by ONE launch that reads the factor once and makes one pass over B.
//
// One warp owns one (matrix, 32-vector slab) task:
//   * the factor is staged once into shared memory (column-major, identity-padded to NP,
//     reciprocal diagonal precomputed) and is only ever read back as warp-wide broadcasts
//     (LDS.128 = two factor entries for all 32 lanes);
//   * every lane owns ONE right-hand-side vector, held in NP registers, so no cross-lane
//     traffic is needed during the substitution and all 32 lanes do useful FMAs:
//       side R: vector = row of B  (lanes = consecutive rows -> direct coalesced access),
//       side L: vector = column of B (B tile is transposed through padded shared memory,
//               instead of the reference's stride-ldb per-lane loads, kernels.cuh:580-589).
// Substitution forms (L = lower factor):
//   forward  (R/Trans: X L^T = aB;  L/NoTrans: L X = aB):  x_j = b_j / L_jj ; b_k -= x_j L_kj (k>j)
//   backward (R/NoTrans: X L = aB;  L/Trans: L^T X = aB):  x_j = (b_j - sum_{k>j} x_k L_kj) / L_jj
// both walk column j of L below the diagonal, which is contiguous in the staged copy.
// POTRS (side R) = forward then backward on the same registers (reference order:
// trsm(R,L,T) then trsm(R,L,N), Xpotrs_batch_drivers.cuh:94-171).
#pragma once

#include "common.cuh"

namespace kblasx {

enum TriOp { TRI_FORWARD = 0, TRI_BACKWARD = 1, TRI_BOTH = 2 };

// ptxas hoists the (volatile) shared-memory loads of later columns far ahead of the FMAs that
// consume them; unbounded, that costs > 200 registers or even kilobytes of spills for the fused
// forward+backward solve.  A warp-level sync every KX_TRI_FENCE columns bounds the look-ahead.
#ifndef KX_TRI_FENCE
#define KX_TRI_FENCE 8
#endif
__device__ __forceinline__ void sched_fence() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }

// Stage the k x k lower factor of one matrix into shared memory: Ls[row + col*NP], padded
// with the identity; invd[j] = 1 / L_jj.  One warp, coalesced column reads.
template <typename T, int NP>
__device__ __forceinline__ void stage_factor(const T *__restrict__ A, int lda, int k, T *__restrict__ Ls,
                                             T *__restrict__ invd, int lane) {
  constexpr int SE = SectorElems<T>::value;
#pragma unroll
  for (int col = 0; col < NP; ++col) {
    T v = (lane == col) ? T(1) : T(0);
    if (lane < NP) {
      if (lane < k && col < k) {
        v = T(0);
        if ((lane | (SE - 1)) >= col) v = ldg_stream(A + lane + (long)col * lda);
      }
      Ls[lane + col * NP] = v;
    }
    if (lane == col) invd[col] = T(1) / v;
  }
  __syncwarp();
}

// forward substitution on the NP-vector x (see file header)
template <typename T, int NP>
__device__ __forceinline__ void tri_forward(T (&x)[NP], const T *__restrict__ Ls, const T *__restrict__ invd) {
  typedef typename Vec2T<T>::type V2;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j % KX_TRI_FENCE == 0) sched_fence();
    x[j] *= lds_one(invd + j);
    const T nx = -x[j];
#pragma unroll
    for (int p = (j + 1) / 2; p < NP / 2; ++p) {
      const V2 l2 = lds_pair(Ls + 2 * p + j * NP);
      if (2 * p > j) x[2 * p] = fma_t(nx, l2.x, x[2 * p]);
      x[2 * p + 1] = fma_t(nx, l2.y, x[2 * p + 1]);
    }
  }
}

// backward substitution on the NP-vector x (two partial sums shorten the FMA chain)
template <typename T, int NP>
__device__ __forceinline__ void tri_backward(T (&x)[NP], const T *__restrict__ Ls, const T *__restrict__ invd) {
  typedef typename Vec2T<T>::type V2;
#pragma unroll
  for (int j = NP - 1; j >= 0; --j) {
    if (j % KX_TRI_FENCE == KX_TRI_FENCE - 1) sched_fence();
    T acc0 = x[j], acc1 = T(0);
#pragma unroll
    for (int p = (j + 1) / 2; p < NP / 2; ++p) {
      const V2 l2 = lds_pair(Ls + 2 * p + j * NP);
      if (2 * p > j) acc0 = fma_t(-x[2 * p], l2.x, acc0);
      acc1 = fma_t(-x[2 * p + 1], l2.y, acc1);
    }
    x[j] = (acc0 + acc1) * lds_one(invd + j);
  }
}

// shared memory per warp, in elements of T
template <int NP, bool LEFT>
struct TriSmem {
  static constexpr int factor = NP * NP + NP;          // Ls + invd
  static constexpr int tile_stride = NP + 1;            // odd stride: conflict-free transposed reads
  static constexpr int tile = LEFT ? 32 * tile_stride : 0;  // k x 32 tile of B, stored [column][row]
  static constexpr int per_warp = factor + tile;
};

// k = order of the triangular factor (n for side R, m for side L); vec = the other dimension
// of B (number of independent vectors).  OP selects forward / backward / both (potrs).
template <typename T, int NP, bool LEFT, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, LEFT ? 3 : 4)
tri_solve_small_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                       BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount, const int slabs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *Ls = reinterpret_cast<T *>(smem_raw) + warp * TriSmem<NP, LEFT>::per_warp;
  T *invd = Ls + NP * NP;
  T *tile = invd + NP;
  constexpr int TS = TriSmem<NP, LEFT>::tile_stride;

  const long task = (long)blockIdx.x * WARPS + warp;  // (matrix, slab)
  if (task >= (long)batchCount * slabs) return;        // warp-uniform
  const long mat = task / slabs;
  const int v0 = (int)(task % slabs) * 32;             // first vector of this slab

  const T *__restrict__ A = Aref.at(mat);
  T *__restrict__ B = Bref.at(mat);
  stage_factor<T, NP>(A, lda, k, Ls, invd, lane);

  T x[NP];
  const int my = v0 + lane;  // my vector
  if (!LEFT) {
    // vector = row `my` of B; element j at B[my + j*ldb]
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = (my < vec && j < k) ? alpha * ldg_stream(B + my + (long)j * ldb) : T(0);
  } else {
    // vector = column `my` of B; stage the k x 32 tile as tile[c*TS + i] (coalesced reads of B)
    for (int c = 0; c < 32; ++c) {
      const int colB = v0 + c;
      if (lane < k && colB < vec) tile[c * TS + lane] = ldg_stream(B + lane + (long)colB * ldb);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = (my < vec && j < k) ? alpha * tile[lane * TS + j] : T(0);
  }

  if (OP == TRI_FORWARD || OP == TRI_BOTH) tri_forward<T, NP>(x, Ls, invd);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) tri_backward<T, NP>(x, Ls, invd);

  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (my < vec && j < k) stg_stream(B + my + (long)j * ldb, x[j]);
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) tile[lane * TS + j] = x[j];
    __syncwarp();
    for (int c = 0; c < 32; ++c) {
      const int colB = v0 + c;
      if (lane < k && colB < vec) stg_stream(B + lane + (long)colB * ldb, tile[c * TS + lane]);
    }
  }
}

}  // namespace kblasx
