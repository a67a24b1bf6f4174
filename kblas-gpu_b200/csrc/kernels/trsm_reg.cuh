// kernels/trsm_reg.cuh -- register-resident batched triangular solves for k <= 16 and at most 16
// right-hand-side vectors: 2 (GP = 16) or 4 (GP = 8) problems per warp, no shared-memory staging of
// the factor (sm_100a).
//
// Same lane mapping as the reference's small kernels (Xtrsm_batch_kernels.cuh:36-133: lane = row of
// the factor, width-TX shuffles broadcast L[i][j]) -- for these sizes the whole problem is a few
// cache lines and the shortest instruction stream wins -- with the things that kept the reference
// from the memory floor on B200 removed: one fused launch for potrs (forward + backward on the same
// registers), reciprocal diagonal computed once per lane instead of a division per column, all
// global loads predicated and in flight before the first use, and the side-L right-hand sides
// transposed through shared memory instead of stride-ldb loads (kernels.cuh:580-589).
//   lane lg of a group holds  a[c] = L[lg][c]  (row lg of the factor) and its own vector x.
//   forward  (R/T, L/N):  x_j *= 1/L_jj ;  x_i -= x_j * L[i][j]  (i > j),  L[i][j] = shfl(a[j], i)
//   backward (R/N, L/T):  x_j = (x_j - sum_{i>j} x_i * L[i][j]) / L_jj
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"  // TriOp, sched_fence

namespace kblasx {

// FULL: k == NP and every lane of a group owns a row (side R: vec == GP; side L: NP == GP == vec) -- no bounds predicates
// at all.  These kernels are issue-bound (one warp = 4 problems of 8 x 8 is ~500 instructions for 34 FMAs per lane), so
// every predicate and 64-bit address multiply that goes away shows up in the time; the unmodified reference was 4-14 %
// faster on fp32 n = 8 with the generic form (profiles/r02_ours_vs_reference.txt).
template <typename T, int NP, int GP, bool LEFT, int OP, int WARPS, bool STRIDED, bool FULL = false>
__global__ void __launch_bounds__(WARPS * 32, OP == TRI_BOTH ? 1 : (sizeof(T) * NP > 64 ? 24 : 32) / WARPS)  // <= 80 / 64 registers
tri_solve_reg_kernel(const int k, const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                     BatchRef<T, STRIDED> Bref, const int ldb, const int batchCount, const int ahead) {
  static_assert(NP <= GP, "one lane per factor row");
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

  // ---- every global load up front: my row of the factor, my row of B -------------------------------
  constexpr int SE = SectorElems<T>::value;
  T a[NP];
  const int nrow = LEFT ? k : vec;  // rows / columns of B
  const int ncol = LEFT ? vec : k;
  const bool hrow = FULL ? live : (live && (lg < nrow));
  constexpr int NC = LEFT ? GP : NP;
  T t[NC];
  if (FULL) {
    // whole sectors of my factor row (entries above the diagonal are never used), my whole row of B; column pointers advance
    // by the leading dimension instead of a 64-bit multiply per element
    const T *pa = A + lg;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      a[c] = T(0);
      if (NP <= SE) a[c] = ldg_stream(pa);
      else ldg_stream_if(a[c], pa, (lg | (SE - 1)) >= c);
      pa += lda;
    }
    const T *pb = B + lg;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      t[c] = ldg_stream(pb);
      pb += ldb;
    }
  } else {
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      a[c] = (c == lg) ? T(1) : T(0);  // identity padding for k < NP
      ldg_stream_if(a[c], A + (long)lg + (long)c * lda, lg < k && c < k && c <= lg);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      t[c] = T(0);
      ldg_stream_if(t[c], B + (long)lg + (long)c * ldb, hrow && c < ncol);
    }
  }
  if (ahead > 0) {
    // the matrix that the same lane group of a later CTA will own: the lines its factor and its B span, into L2 (hints)
    const long pmat = mat + (long)ahead * WARPS * MPW;
    if (pmat < (long)batchCount) {
      constexpr int ES = (int)sizeof(T);
      const char *pa = reinterpret_cast<const char *>(Aref.at(pmat));
      const char *pb = reinterpret_cast<const char *>(Bref.at(pmat));
      const int abytes = ((k - 1) * lda + k) * ES, bbytes = ((ncol - 1) * ldb + nrow) * ES;
#pragma unroll 1
      for (int off = lg * 128; off < abytes; off += GP * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + off));
#pragma unroll 1
      for (int off = lg * 128; off < bbytes; off += GP * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + off));
    }
  }
  sched_fence();

  // reciprocal of my own diagonal entry (one division per lane, all lanes at once)
  T dg = T(1);
#pragma unroll
  for (int c = 0; c < NP; ++c) dg = (c == lg) ? a[c] : dg;
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

  if (OP == TRI_FORWARD || OP == TRI_BOTH) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      x[j] *= shfl_seg<GP>(inv, j);
      const T nx = -x[j];
#pragma unroll
      for (int i = j + 1; i < NP; ++i) x[i] = fma_t(nx, shfl_seg<GP>(a[j], i), x[i]);
    }
  }
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) {
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
      T acc0 = x[j], acc1 = T(0);
#pragma unroll
      for (int i = j + 1; i < NP; ++i) {
        const T lij = shfl_seg<GP>(a[j], i);
        if ((i - j) & 1) acc0 = fma_t(-x[i], lij, acc0);
        else acc1 = fma_t(-x[i], lij, acc1);
      }
      x[j] = (acc0 + acc1) * shfl_seg<GP>(inv, j);
    }
  }

  T *Bs = launder(B);  // fresh store addresses: the 16 load addresses would otherwise stay live through the solve
  if (!LEFT) {
    T *ps = Bs + lg;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      stg_stream_if(ps, x[j], FULL ? hrow : (hrow && j < ncol));
      ps += ldb;
    }
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) tile[lg * (NP + 1) + j] = x[j];
    __syncwarp();
    T *ps = Bs + lg;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const T out = (lg < NP) ? tile[c * (NP + 1) + lg] : T(0);
      stg_stream_if(ps, out, FULL ? hrow : (hrow && c < ncol));
      ps += ldb;
    }
  }
}

}  // namespace kblasx
