// kernels/trsm_left_vec.cuh -- side-L triangular solves (and the fused side-L POTRS) with a k x k lower factor, k <= 32,
// strided batches with 16-byte aligned operands: every global access is a 16-byte one (sm_100a).
//
// Why: on config 3 (side L, n = nrhs = 32, 2^20 problems) the element-wise kernels are not bound by DRAM but by the
// instructions AROUND the memory operations -- ncu of tri_solve_small<float,32,LEFT>
// (profiles/r02_ncu_strsm32_LLN_tri_small.json): 1930 warp-instructions per problem of which 498 are FFMAs, issue slots 69 %
// busy, LSU wavefronts 81 %, `not_selected` the top stall: one predicated 4-byte LDG / STS / LDS / STG plus its 64-bit
// address arithmetic and predicate per element, twice (in and out).  Here
//   * the factor and the k x 32 slab of B go global -> shared memory as 16-byte cp.async (LDGSTS.128, zero-filled where
//     predicated off): one instruction per 4 floats / 2 doubles, no staging registers, 512 contiguous bytes per instruction;
//   * a lane's right-hand side is one COLUMN of B, i.e. one padded row of the tile: it is read, and the solution written
//     back, with LDS.128 / STS.128 (row pitch NP + 16 bytes: the eight 16-byte accesses of a quarter-warp hit eight different
//     bank groups);
//   * the solution leaves through LDS.128 + 16-byte streaming stores, again 512 contiguous bytes per instruction.
// The substitution itself is the broadcast-LDS.128 scheme of kernels/trsm_small.cuh (one vector per lane: 64 registers of
// fp64 state, so 12 warps per SM instead of the 6 of the two-vector kernel).
// Replaces the reference's side-L path: K8/K9 with stride-ldb per-lane accesses (Xtrsm_batch_kernels.cuh:551-723) and the
// recursion above them (Xtrsm_batch_drivers.cuh:127-266).
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"  // sched_fence
#include "vec16.cuh"
#include "trsm_dual.cuh"   // prefetch_solve_task_l2

namespace kblasx {

#ifndef KX_TLV_FENCE_F64
#define KX_TLV_FENCE_F64 2
#endif
#ifndef KX_TLV_MAXB
#define KX_TLV_MAXB 6  // CTAs per SM: 6 x 4 warps leave 80 registers per thread
#endif
template <typename T, int NP>
struct TriLeftVecSmem {
  static constexpr int VW = 16 / (int)sizeof(T);
  static constexpr int TS = NP + VW;                            // tile row pitch: 16-byte aligned, conflict-free LDS.128 / STS.128
  static constexpr int per_warp = NP * NP + NP + 32 * TS;       // factor, reciprocal diagonal, 32 columns of B
  static constexpr int warps = (sizeof(T) == 8 && NP == 32) ? 2 : 4;
  static constexpr int by_smem = (227 * 1024) / (warps * per_warp * (int)sizeof(T) + 1024);
  static constexpr int ctas_per_sm = by_smem < KX_TLV_MAXB ? by_smem : KX_TLV_MAXB;
  static_assert((NP * sizeof(T)) % 16 == 0, "factor columns and the tile must stay 16-byte aligned");
};

// forward / backward substitution of one NP-vector per lane against the staged factor (column-major Ls, reciprocal diagonal
// invd): broadcast LDS.128 = VW factor entries for all 32 lanes
template <typename T, int NP, int OP>
__device__ __forceinline__ void tri_vec_substitute(T (&x)[NP], const T *Ls, const T *invd) {
  constexpr int VW = 16 / (int)sizeof(T);
  constexpr int NV = NP / VW;
  constexpr int FENCE = sizeof(T) == 8 ? 4 : 8;    // columns between scheduling fences (bounds ptxas' LDS look-ahead), backward
  constexpr int FENCE_F = sizeof(T) == 8 ? KX_TLV_FENCE_F64 : 4;  // forward: every hoisted column costs NP registers
  if (OP == TRI_FORWARD || OP == TRI_BOTH) {
    T dv[VW];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (j % FENCE_F == 0) sched_fence();
      if (j % VW == 0) lds_vec(dv, invd + j);
      x[j] *= dv[j % VW];
      const T nx = -x[j];
#pragma unroll
      for (int v = (j + 1) / VW; v < NV; ++v) {
        T c[VW];
        lds_vec(c, Ls + v * VW + j * NP);
#pragma unroll
        for (int e = 0; e < VW; ++e)
          if (v * VW + e > j) x[v * VW + e] = fma_t(nx, c[e], x[v * VW + e]);
      }
    }
  }
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) {
    T dv[VW];
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
      if ((NP - 1 - j) % FENCE == 0) sched_fence();
      if (j % VW == VW - 1) lds_vec(dv, invd + j - (VW - 1));
      T acc[4] = {x[j], T(0), T(0), T(0)};
#pragma unroll
      for (int v = (j + 1) / VW; v < NV; ++v) {
        T c[VW];
        lds_vec(c, Ls + v * VW + j * NP);
#pragma unroll
        for (int e = 0; e < VW; ++e)
          if (v * VW + e > j) acc[(v * VW + e) & 3] = fma_t(-x[v * VW + e], c[e], acc[(v * VW + e) & 3]);
      }
      x[j] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * dv[j % VW];
    }
  }
}

// k = order of the factor (a multiple of 16 / sizeof(T), <= NP), vec = columns of B.  lda, ldb, the batch strides and both
// base pointers are multiples of 16 bytes (the launcher checks).
template <typename T, int NP, int OP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
tri_left_vec_kernel(const int k, const int vec, const T alpha, const T *__restrict__ A0, const int lda, const long strideA,
                    T *__restrict__ B0, const int ldb, const long strideB, const int batchCount, const int slabs, const int ahead) {
  constexpr int VW = TriLeftVecSmem<T, NP>::VW;
  constexpr int NV = NP / VW;
  constexpr int TS = TriLeftVecSmem<T, NP>::TS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *Ls = reinterpret_cast<T *>(smem_raw) + warp * TriLeftVecSmem<T, NP>::per_warp;
  T *invd = Ls + NP * NP;
  T *tile = invd + NP;

  const long task = (long)blockIdx.x * WARPS + warp;  // (matrix, 32-column slab)
  if (task >= (long)batchCount * slabs) return;        // warp-uniform
  const long mat = task / slabs;
  const int v0 = (int)(task % slabs) * 32;
  const int ncol = (vec - v0 < 32) ? (vec - v0) : 32;
  const T *__restrict__ A = A0 + mat * strideA;
  T *__restrict__ B = B0 + mat * strideB + (long)v0 * ldb;

  // ---- factor and slab: global -> shared, asynchronously.  Chunk i = (column i / NV, rows (i % NV) * VW ...) ----------
#pragma unroll
  for (int i0 = 0; i0 < NP * NV; i0 += 32) {
    const int i = i0 + lane, c = i / NV, r0 = (i % NV) * VW;
    if ((NP * NV) % 32 == 0 || c < NP)
      cp_async16_if(Ls + c * NP + r0, A + (long)c * lda + r0, c < k && r0 < k && r0 + VW - 1 >= c);  // chunks above the diagonal: zeros
  }
#pragma unroll
  for (int i0 = 0; i0 < 32 * NV; i0 += 32) {
    const int i = i0 + lane, c = i / NV, r0 = (i % NV) * VW;
    cp_async16_if(tile + c * TS + r0, B + (long)c * ldb + r0, c < ncol && r0 < k);
  }
  if (ahead > 0) {  // the operands of the task a later CTA will own: into L2 (see prefetch_solve_task_l2)
    const long ptask = task + (long)ahead * WARPS;
    if (ptask < (long)batchCount * slabs) {
      const long pmat = ptask / slabs;
      prefetch_solve_task_l2<T, NP, true>(A0 + pmat * strideA, lda, B0 + pmat * strideB, ldb, (int)(ptask % slabs) * 32, vec, lane, 32);
    }
  }
  cp_async_wait_all();
  __syncwarp();
  if (lane < NP) invd[lane] = lane < k ? T(1) / Ls[lane + lane * NP] : T(1);
  __syncwarp();

  // ---- my vector = column `lane` of the slab = row `lane` of the tile ------------------------------------------------
  T x[NP];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    T c[VW];
    lds_vec(c, tile + lane * TS + q * VW);
#pragma unroll
    for (int e = 0; e < VW; ++e) x[q * VW + e] = alpha * c[e];
  }

  tri_vec_substitute<T, NP, OP>(x, Ls, invd);

  // ---- back through the tile, then 16-byte streaming stores ------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    T c[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) c[e] = x[q * VW + e];
    sts_vec(tile + lane * TS + q * VW, c);
  }
  __syncwarp();
  T *Bs = launder(B);
#pragma unroll
  for (int i0 = 0; i0 < 32 * NV; i0 += 32) {
    const int i = i0 + lane, c = i / NV, r0 = (i % NV) * VW;
    if (c < ncol && r0 < k) {
      T o[VW];
      lds_vec(o, tile + c * TS + r0);
      stg_vec_stream(Bs + (long)c * ldb + r0, o);
    }
  }
}

// ---- side R: vector = row of B, lanes = consecutive rows, so B needs no staging at all; only the factor goes through
// shared memory (16-byte cp.async).  One vector per lane keeps the kernel at ~100 registers: 16 warps per SM.
template <typename T, int NP>
struct TriRightVecSmem {
  static constexpr int per_warp = NP * NP + NP;
  static constexpr int warps = 4;
  static constexpr int ctas_per_sm = 4;
};

template <typename T, int NP, int OP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
tri_right_vec_kernel(const int k, const int vec, const T alpha, const T *__restrict__ A0, const int lda, const long strideA,
                     T *__restrict__ B0, const int ldb, const long strideB, const int batchCount, const int slabs, const int ahead) {
  constexpr int VW = 16 / (int)sizeof(T);
  constexpr int NV = NP / VW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *Ls = reinterpret_cast<T *>(smem_raw) + warp * TriRightVecSmem<T, NP>::per_warp;
  T *invd = Ls + NP * NP;

  const long task = (long)blockIdx.x * WARPS + warp;  // (matrix, 32-row slab)
  if (task >= (long)batchCount * slabs) return;        // warp-uniform
  const long mat = task / slabs;
  const int my = (int)(task % slabs) * 32 + lane;      // my row of B
  const T *__restrict__ A = A0 + mat * strideA;
  T *__restrict__ B = B0 + mat * strideB;

#pragma unroll
  for (int i0 = 0; i0 < NP * NV; i0 += 32) {
    const int i = i0 + lane, c = i / NV, r0 = (i % NV) * VW;
    if ((NP * NV) % 32 == 0 || c < NP)
      cp_async16_if(Ls + c * NP + r0, A + (long)c * lda + r0, c < k && r0 < k && r0 + VW - 1 >= c);
  }
  T x[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    x[j] = T(0);
    ldg_stream_if(x[j], B + my + (long)j * ldb, my < vec && j < k);
  }
  if (ahead > 0) {
    const long ptask = task + (long)ahead * WARPS;
    if (ptask < (long)batchCount * slabs) {
      const long pmat = ptask / slabs;
      prefetch_solve_task_l2<T, NP, false>(A0 + pmat * strideA, lda, B0 + pmat * strideB, ldb, (int)(ptask % slabs) * 32, vec, lane, 32);
    }
  }
  cp_async_wait_all();
  __syncwarp();
  if (lane < NP) invd[lane] = lane < k ? T(1) / Ls[lane + lane * NP] : T(1);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NP; ++j) x[j] *= alpha;

  tri_vec_substitute<T, NP, OP>(x, Ls, invd);

  T *Bs = launder(B);
#pragma unroll
  for (int j = 0; j < NP; ++j) stg_stream_if(Bs + my + (long)j * ldb, x[j], my < vec && j < k);
}

}  // namespace kblasx
