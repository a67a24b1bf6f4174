// kernels/trsm_bcast.cuh -- batched triangular solves for k <= 16 and at most 16 right-hand-side
// vectors with the factor read as L1-broadcast vector loads: 2 (GP = 16) or 4 (GP = 8) problems per
// warp, no shared-memory staging, no shuffles in the substitution (sm_100a).
//
// The register kernel (kernels/trsm_reg.cuh, the reference's own mapping,
// Xtrsm_batch_kernels.cuh:36-133) keeps one factor row per lane and broadcasts L[i][j] with a
// shuffle per (i, j): 272 SHFL.32 per fp64 16 x 16 problem pair, 116 registers, 14 resident warps;
// ncu (profiles/r01_ncu_dtrsm16_RLN_tri_reg.json): LSU pipe 53 % busy, 53 % of the stalls on the
// scoreboard, 69 % of the DRAM peak.  Here every lane owns ONE right-hand-side vector x (side R: a row of
// B; side L: a column of B, transposed through shared memory) and reads the factor column by
// column straight from global memory: all lanes of a problem read the SAME 16 bytes, so one
// LDG.128 delivers 2 (fp64) / 4 (fp32) entries of L to the whole lane group out of L1 -- 64
// loads instead of 272 shuffles, and no factor registers (x is the only array).  The diagonal loads
// that feed the reciprocals touch every line of the factor first, so the DRAM fetches of all of
// its lines are in flight together with the loads of B.
//   forward  (R/T, L/N):  x_j *= 1/L_jj ;  x_i -= x_j * L[i][j]  (i > j)
//   backward (R/N, L/T):  x_j = (x_j - sum_{i>j} x_i * L[i][j]) / L_jj
// Vector loads need k == NP and 16-byte aligned columns: the host dispatch sends ragged k / odd lda to
// the register kernel; a pointer-array entry that turns out to be unaligned takes the (out-of-line,
// slow, correct) scalar path of this kernel.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"  // TriOp, sched_fence

#ifndef KX_BCAST_FP_SCALE
#define KX_BCAST_FP_SCALE 1
#endif

namespace kblasx {

// L1-cached (read-only path) loads of the factor.  asm volatile: the forward and the backward pass
// of POTRS read the same entries, and a plain load lets the compiler keep all of them in registers.
// (the byte offset is an immediate: the only address register is the column pointer)
template <int OFF>
__device__ __forceinline__ void ldg_bcast_vec(double (&v)[2], const double *p) {
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2 + %3];" : "=d"(v[0]), "=d"(v[1]) : "l"(p), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void ldg_bcast_vec(float (&v)[4], const float *p) {
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4 + %5];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "n"(OFF));
}
__device__ __forceinline__ void ldg_cached_if(double &v, const double *p, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.nc.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg_cached_if(float &v, const float *p, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.nc.f32 %0, [%1]; }" : "+f"(v) : "l"(p), "r"((int)pred));
}

// entries v*VW .. v*VW+VW-1 of column j of the factor (zero outside the k x k lower triangle in the scalar path;
// the vector path returns whatever the strict upper triangle holds -- callers only use i > j)
// `col` = address of column j
template <typename T, bool VEC, int V>
__device__ __forceinline__ void load_factor_chunk(T (&c)[16 / sizeof(T)], const T *col, const int k, const int j) {
  constexpr int VW = 16 / (int)sizeof(T);
  if constexpr (VEC) {
    ldg_bcast_vec<V * 16>(c, col);
  } else {
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const int i = V * VW + e;
      c[e] = T(0);
      ldg_cached_if(c[e], col + i, i > j && i < k && j < k);
    }
  }
}

// x[i] -= xj * L[i][j] for every i > j of chunks V .. NV-1 (compile-time recursion: V feeds an immediate)
template <typename T, int NP, bool VEC, int V>
__device__ __forceinline__ void column_axpy(T (&x)[NP], const T nx, const T *col, const int k, const int j) {
  constexpr int VW = 16 / (int)sizeof(T);
  if constexpr (V < NP / VW) {
    if (V >= (j + 1) / VW) {  // resolved at compile time once the column loop is unrolled
      T c[VW];
      load_factor_chunk<T, VEC, V>(c, col, k, j);
#pragma unroll
      for (int e = 0; e < VW; ++e)
        if (V * VW + e > j) x[V * VW + e] = fma_t(nx, c[e], x[V * VW + e]);
    }
    column_axpy<T, NP, VEC, V + 1>(x, nx, col, k, j);
  }
}
// acc -= sum_{i > j} x[i] * L[i][j] over chunks V .. NV-1, two accumulators
template <typename T, int NP, bool VEC, int V>
__device__ __forceinline__ void column_dot(const T (&x)[NP], T &acc0, T &acc1, const T *col, const int k, const int j) {
  constexpr int VW = 16 / (int)sizeof(T);
  if constexpr (V < NP / VW) {
    if (V >= (j + 1) / VW) {
      T c[VW];
      load_factor_chunk<T, VEC, V>(c, col, k, j);
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const int i = V * VW + e;
        if (i > j) {
          if ((i - j) & 1) acc0 = fma_t(-x[i], c[e], acc0);
          else acc1 = fma_t(-x[i], c[e], acc1);
        }
      }
    }
    column_dot<T, NP, VEC, V + 1>(x, acc0, acc1, col, k, j);
  }
}

template <typename T, int NP, int GP, int OP, bool VEC>
__device__ __forceinline__ void tri_bcast_solve_body(T (&x)[NP], const T *__restrict__ A, const int lda, const int k, const T inv) {
  constexpr int VW = 16 / (int)sizeof(T);
  constexpr int NV = NP / VW;
  // ptxas hoists every load of the unrolled substitution to the top (250+ registers, or KBs of spills
  // under a cap): a scheduling fence every FP columns bounds the loads in flight to ~32 registers
  constexpr int FP = ((8 / NV) > 0 ? (8 / NV) : 1) * KX_BCAST_FP_SCALE;
  if (OP == TRI_FORWARD || OP == TRI_BOTH) {
    const T *col = launder(A);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (j % FP == 0) sched_fence();
      x[j] *= shfl_seg<GP>(inv, j);
      column_axpy<T, NP, VEC, 0>(x, -x[j], col, k, j);
      col += lda;
    }
  }
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) {
    const T *col = launder(A) + (long)(NP - 1) * lda;
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
      if ((NP - 1 - j) % FP == 0) sched_fence();
      T acc0 = x[j], acc1 = T(0);
      column_dot<T, NP, VEC, 0>(x, acc0, acc1, col, k, j);
      x[j] = (acc0 + acc1) * shfl_seg<GP>(inv, j);
      col -= lda;
    }
  }
}

template <typename T, int NP, int GP, int OP>
__device__ __forceinline__ void tri_bcast_solve_vec(T (&x)[NP], const T *__restrict__ A, const int lda, const T inv) {
  tri_bcast_solve_body<T, NP, GP, OP, true>(x, A, lda, NP, inv);
}
// unaligned pointer-array entry: kept out of line so that it costs the hot path no registers
template <typename T, int NP, int GP, int OP>
__device__ __noinline__ void tri_bcast_solve_scalar(T *x_io, const T *__restrict__ A, const int lda, const int k, const T inv) {
  T x[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) x[j] = x_io[j];
  tri_bcast_solve_body<T, NP, GP, OP, false>(x, A, lda, k, inv);
#pragma unroll
  for (int j = 0; j < NP; ++j) x_io[j] = x[j];
}

template <typename T, int NP, int GP, bool LEFT, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, (sizeof(T) * NP > 64 ? 24 : 32) / WARPS)  // <= 80 / 64 registers: 24 / 32 resident warps
tri_solve_bcast_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                       BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount) {
  static_assert(NP <= GP, "the lane group also covers the k rows of B (side L) and the k diagonal entries");
  constexpr int MPW = 32 / GP;  // problems per warp
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane / GP, lg = lane % GP;
  const long mat = ((long)blockIdx.x * WARPS + warp) * MPW + g;
  const bool live = mat < (long)batchCount;
  const long msafe = live ? mat : (long)batchCount - 1;
  const T *__restrict__ A = Aref.at(msafe);
  T *__restrict__ B = Bref.at(msafe);

  // side L: padded transpose tile per problem (stride NP+1)
  __shared__ T tiles[LEFT ? WARPS * MPW * GP * (NP + 1) : 1];
  T *tile = tiles + (LEFT ? (warp * MPW + g) * GP * (NP + 1) : 0);

  // ---- loads that go to DRAM, all up front: my diagonal entry (touches every line of the factor), my row of B
  T dg = T(1);  // identity padding for k < NP
  ldg_cached_if(dg, A + (long)lg + (long)lg * lda, lg < k);
  const int nrow = LEFT ? k : vec;  // rows / columns of B
  const int ncol = LEFT ? vec : k;
  const bool hrow = live && (lg < nrow);
  constexpr int NC = LEFT ? GP : NP;
  T t[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    t[c] = T(0);
    ldg_stream_if(t[c], B + (long)lg + (long)c * ldb, hrow && c < ncol);
  }
  sched_fence();
  const T inv = T(1) / dg;

  T x[NP];
  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = alpha * t[j < NC ? j : 0];
  } else {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (lg < NP) tile[c * (NP + 1) + lg] = t[c];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = alpha * tile[lg * (NP + 1) + j];
  }

  // vector loads need full columns and 16-byte alignment (warp-uniform decision)
  const bool aligned = (k == NP) && ((reinterpret_cast<unsigned long long>(A) | ((unsigned long long)lda * sizeof(T))) & 15ull) == 0;
  if (__all_sync(0xffffffffu, aligned)) tri_bcast_solve_vec<T, NP, GP, OP>(x, A, lda, inv);
  else tri_bcast_solve_scalar<T, NP, GP, OP>(x, A, lda, k, inv);

  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j) stg_stream_if(B + (long)lg + (long)j * ldb, x[j], hrow && j < ncol);
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) tile[lg * (NP + 1) + j] = x[j];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const T out = (lg < NP) ? tile[c * (NP + 1) + lg] : T(0);
      stg_stream_if(B + (long)lg + (long)c * ldb, out, hrow && c < ncol);
    }
  }
}

}  // namespace kblasx
