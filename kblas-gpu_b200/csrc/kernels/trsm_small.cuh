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


// ptxas hoists the (volatile) shared-memory loads of later columns far ahead of the FMAs that
// consume them; unbounded, that costs > 200 registers or even kilobytes of spills for the fused
// forward+backward solve.  A warp-level sync every KX_TRI_FENCE columns bounds the look-ahead.
#ifndef KX_TRI_FENCE
#define KX_TRI_FENCE 8
#endif
__device__ __forceinline__ void sched_fence() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }

// Stage the k x k lower factor of one matrix into shared memory: Ls[row + col*NP], padded
// with the identity; invd[j] = 1 / L_jj.  One warp, coalesced column reads.  Two phases so that
// all NP predicated loads are in flight before the first shared-memory store waits on one.
template <typename T, int NP>
__device__ __forceinline__ void stage_factor_load(T (&v)[NP], const T *__restrict__ A, int lda, int k, int lane, int flags = 0) {
  constexpr int SE = SectorElems<T>::value;
  if (flags & TRI_FLAG_UPPER) {
#pragma unroll
    for (int col = 0; col < NP; ++col) {
      const bool inside = (lane < k) && (col < k);
      v[col] = (!inside && lane == col) ? T(1) : T(0);
      ldg_stream_if(v[col], A + col + (long)lane * lda, inside && lane >= col);
    }
    return;
  }
#pragma unroll
  for (int col = 0; col < NP; ++col) {
    const bool inside = (lane < k) && (col < k);
    v[col] = (!inside && lane == col) ? T(1) : T(0);
    ldg_stream_if(v[col], A + lane + (long)col * lda, inside && ((lane | (SE - 1)) >= col));
  }
}
template <typename T, int NP>
__device__ __forceinline__ void stage_factor_store(const T (&v)[NP], T *__restrict__ Ls, T *__restrict__ invd, int lane,
                                                   int flags = 0) {
  // every load of the batch is issued before the first store can wait on one (ptxas otherwise
  // interleaves LDG / STS to save registers and the in-order warp eats one memory latency per column)
  sched_fence();
  T dg = T(1);  // my own diagonal entry (select chain: ONE division per lane, all lanes at once;
                // a per-column `if (lane == col) 1/v` serialises NP single-lane divisions)
#pragma unroll
  for (int col = 0; col < NP; ++col) {
    if (lane < NP) Ls[lane + col * NP] = v[col];
    dg = (lane == col) ? v[col] : dg;
  }
  if (lane < NP) invd[lane] = (flags & TRI_FLAG_UNIT) ? T(1) : T(1) / dg;
  __syncwarp();
}
template <typename T, int NP>
__device__ __forceinline__ void stage_factor(const T *__restrict__ A, int lda, int k, T *__restrict__ Ls,
                                             T *__restrict__ invd, int lane, int flags = 0) {
  T v[NP];
  stage_factor_load<T, NP>(v, A, lda, k, lane, flags);
  stage_factor_store<T, NP>(v, Ls, invd, lane, flags);
}

// Column j of the staged factor, rows > j, as NP/2 register pairs (pair p = rows 2p, 2p+1).
// Issued one column AHEAD of its use (software pipeline): the FMAs of column j hide the
// shared-memory latency of column j+1; the explicit double buffer gives ptxas the ILP that the
// ordered (volatile) loads would otherwise deny it.
template <typename T, int NP>
__device__ __forceinline__ void load_factor_col(typename Vec2T<T>::type (&buf)[NP / 2], const T *Ls, int j) {
#pragma unroll
  for (int p = 0; p < NP / 2; ++p)
    if (p >= (j + 1) / 2) buf[p] = lds_pair(Ls + 2 * p + j * NP);
}

// forward substitution on the NP-vector x (see file header)
template <typename T, int NP>
__device__ __forceinline__ void tri_forward(T (&x)[NP], const T *__restrict__ Ls, const T *__restrict__ invd) {
  typedef typename Vec2T<T>::type V2;
  V2 cur[NP / 2], nxt[NP / 2];
  load_factor_col<T, NP>(cur, Ls, 0);
  T dinv = lds_one(invd);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j % KX_TRI_FENCE == 0) sched_fence();
    T dnext = T(0);
    if (j + 1 < NP) {
      load_factor_col<T, NP>(nxt, Ls, j + 1);
      dnext = lds_one(invd + j + 1);
    }
    x[j] *= dinv;
    const T nx = -x[j];
#pragma unroll
    for (int p = (j + 1) / 2; p < NP / 2; ++p) {
      if (2 * p > j) x[2 * p] = fma_t(nx, cur[p].x, x[2 * p]);
      x[2 * p + 1] = fma_t(nx, cur[p].y, x[2 * p + 1]);
    }
#pragma unroll
    for (int p = 0; p < NP / 2; ++p) cur[p] = nxt[p];
    dinv = dnext;
  }
}

// backward substitution on the NP-vector x (four partial sums shorten the FMA chain)
template <typename T, int NP>
__device__ __forceinline__ void tri_backward(T (&x)[NP], const T *__restrict__ Ls, const T *__restrict__ invd) {
  typedef typename Vec2T<T>::type V2;
  V2 cur[NP / 2], nxt[NP / 2];
  load_factor_col<T, NP>(cur, Ls, NP - 1);
  T dinv = lds_one(invd + NP - 1);
#pragma unroll
  for (int j = NP - 1; j >= 0; --j) {
    if (j % KX_TRI_FENCE == KX_TRI_FENCE - 1) sched_fence();
    T dnext = T(0);
    if (j > 0) {
      load_factor_col<T, NP>(nxt, Ls, j - 1);
      dnext = lds_one(invd + j - 1);
    }
    T acc[4] = {x[j], T(0), T(0), T(0)};
#pragma unroll
    for (int p = (j + 1) / 2; p < NP / 2; ++p) {
      if (2 * p > j) acc[(2 * p) & 3] = fma_t(-x[2 * p], cur[p].x, acc[(2 * p) & 3]);
      acc[(2 * p + 1) & 3] = fma_t(-x[2 * p + 1], cur[p].y, acc[(2 * p + 1) & 3]);
    }
    x[j] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * dinv;
#pragma unroll
    for (int p = 0; p < NP / 2; ++p) cur[p] = nxt[p];
    dinv = dnext;
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
__global__ void __launch_bounds__(WARPS * 32, 3)
tri_solve_small_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                       BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount, const int slabs, const int flags) {
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
  T fv[NP];
  stage_factor_load<T, NP>(fv, A, lda, k, lane, flags);  // NP predicated loads in flight
  stage_factor_store<T, NP>(fv, Ls, invd, lane, flags);

  T x[NP];
  const int my = v0 + lane;  // my vector
  if (!LEFT) {
    // vector = row `my` of B; element j at B[my + j*ldb]
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      x[j] = T(0);
      ldg_stream_if(x[j], B + my + (long)j * ldb, my < vec && j < k);
    }
    sched_fence();
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] *= alpha;
  } else {
    // vector = column `my` of B; stage the k x 32 tile as tile[c*TS + i] (coalesced reads of B)
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {  // 16 predicated loads in flight, then 16 stores
      T tv[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        tv[c] = T(0);
        ldg_stream_if(tv[c], B + lane + (long)(v0 + c0 + c) * ldb, lane < k && (v0 + c0 + c) < vec);
      }
      sched_fence();
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (lane < NP) tile[(c0 + c) * TS + lane] = tv[c];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = (my < vec && j < k) ? alpha * tile[lane * TS + j] : T(0);
  }

  if (OP == TRI_FORWARD || OP == TRI_BOTH) tri_forward<T, NP>(x, Ls, invd);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) tri_backward<T, NP>(x, Ls, invd);

  T *Bs = launder(B);  // fresh store addresses (common.cuh: launder)
  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j) stg_stream_if(Bs + my + (long)j * ldb, x[j], my < vec && j < k);
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) tile[lane * TS + j] = x[j];
    __syncwarp();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const int colB = v0 + c;
      const T out = (lane < NP) ? tile[c * TS + lane] : T(0);  // lanes >= NP would read past the tile
      stg_stream_if(Bs + lane + (long)colB * ldb, out, lane < k && colB < vec);
    }
  }
}

// ---- packed variant: vec <= 16 right-hand-side vectors per matrix -> several matrices per warp ------
// A group of GP (8 or 16) lanes owns one matrix; lane lg of the group owns vector lg.  Each group has
// its own staged factor, so the substitution's shared-memory reads are 32/GP-address broadcasts.
// (With one matrix per warp an 8 x 8 problem kept 24 of 32 lanes idle: 0.15 of the HBM roofline,
// slower than the reference's 8-lane register kernels, Xtrsm_batch_kernels.cuh:36-133.)
template <typename T, int NP, int GP, bool LEFT, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
tri_solve_packed_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                        BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount) {
  constexpr int MPW = 32 / GP;                 // matrices per warp
  constexpr int FSZ = NP * NP + NP;            // staged factor + reciprocal diagonal
  __shared__ __align__(16) T smem[WARPS * MPW * FSZ];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane / GP, lg = lane % GP;
  T *Lw = smem + warp * MPW * FSZ;

  const long wtask = (long)blockIdx.x * WARPS + warp;  // warp-batch of MPW matrices
  const long mat0 = wtask * MPW;
  if (mat0 >= (long)batchCount) return;  // warp-uniform

  // ---- all global loads of this warp-batch are issued before anything waits: the MPW factors (the
  //      whole warp loads one factor at a time, lane = row) and, below, my row of B ---------------------
  T fv[MPW][NP];
#pragma unroll
  for (int q = 0; q < MPW; ++q) {
    const long mq = (mat0 + q < (long)batchCount) ? mat0 + q : (long)batchCount - 1;
    stage_factor_load<T, NP>(fv[q], Aref.at(mq), lda, k, lane);
  }
  const long mat = mat0 + g;
  T *__restrict__ B = Bref.at(mat < (long)batchCount ? mat : (long)batchCount - 1);
  const T *Ls = Lw + g * FSZ;
  const T *invd = Ls + NP * NP;

  // Coalesced access for both sides: lane lg of a group reads ROW lg of B, one column per instruction
  // (GP consecutive elements = one contiguous segment per matrix).  Side R: that row is my vector.
  // Side L: my vector is COLUMN lg, so the rows are transposed through a padded shared-memory tile
  // (stride NP+1: conflict-free both ways) instead of per-lane strided global accesses.
  __shared__ T tiles[LEFT ? WARPS * MPW * GP * (NP + 1) : 1];
  T *tile = tiles + (LEFT ? (warp * MPW + g) * GP * (NP + 1) : 0);
  const int nrow = LEFT ? k : vec;   // rows of B
  const int ncol = LEFT ? vec : k;   // columns of B
  const bool hrow = (mat < (long)batchCount) && (lg < nrow);
  constexpr int NC = LEFT ? GP : NP; // columns held per lane while loading
  T x[NP];
  {
    T t[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      t[c] = T(0);
      ldg_stream_if(t[c], B + (long)lg + (long)c * ldb, hrow && c < ncol);
    }
#pragma unroll
    for (int q = 0; q < MPW; ++q) stage_factor_store<T, NP>(fv[q], Lw + q * FSZ, Lw + q * FSZ + NP * NP, lane);  // fences first
    if (!LEFT) {
#pragma unroll
      for (int j = 0; j < NP; ++j) x[j] = alpha * t[j < NC ? j : 0];
    } else {
      // tile[c*(NP+1) + r] = B[r][c]  (r = lg < GP rows fit because k <= NP <= ... see dispatch)
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (lg < NP) tile[c * (NP + 1) + lg] = t[c];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NP; ++j) x[j] = alpha * tile[lg * (NP + 1) + j];
    }
  }

  if (OP == TRI_FORWARD || OP == TRI_BOTH) tri_forward<T, NP>(x, Ls, invd);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) tri_backward<T, NP>(x, Ls, invd);

  T *Bs = launder(B);  // fresh store addresses (common.cuh: launder)
  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j) stg_stream_if(Bs + (long)lg + (long)j * ldb, x[j], hrow && j < ncol);
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) tile[lg * (NP + 1) + j] = x[j];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const T out = (lg < NP) ? tile[c * (NP + 1) + lg] : T(0);
      stg_stream_if(Bs + (long)lg + (long)c * ldb, out, hrow && c < ncol);
    }
  }
}

}  // namespace kblasx
