// trsm_dispatch.cuh -- kernel selection of the triangular solves (trsm, potrs, posv), shared by the per-(precision,
// layout, side) instantiation units trsm_inst_*.cu: the kernels are heavily unrolled templates, so the instantiations
// are spread over eight translation units that compile in parallel.
// Counterpart of the reference's kernel table + recursion, Xtrsm_batch_drivers.cuh:54-272.  Here: one launch per call.
#pragma once
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/trsm_small.cuh"
#include "kernels/trsm_blocked.cuh"
#include "kernels/trsm_reg.cuh"
#include "kernels/trsm_dual.cuh"
#include "kernels/trsm_left_vec.cuh"
#include "kernels/trsm_mma.cuh"
#include "tri_batch.h"

namespace kblasx {

template <typename T, int NP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_small(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                            int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4;
  const int slabs = (vec + 31) / 32;
  const long tasks = (long)batchCount * slabs;
  const long grid = (tasks + WARPS - 1) / WARPS;
  const size_t smem = (size_t)WARPS * TriSmem<NP, LEFT>::per_warp * sizeof(T);
  auto kern = tri_solve_small_kernel<T, NP, LEFT, OP, WARPS, STRIDED>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, slabs, h->tri_flags);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// full NP x NP factor: two vectors per lane, two problems per warp, cp.async staging (kernels/trsm_dual.cuh)
template <typename T, int NP, bool LEFT, int OP, bool STRIDED, bool VEC16 = false>
static int launch_tri_dual(KBlasHandle *h, const char *name, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                           BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = LEFT ? 2 : 4;  // side L carries a transpose tile per problem: 2-warp CTAs keep 3 CTAs per SM
  const int slabs = (vec + 31) / 32;
  const long tasks = (long)batchCount * slabs;
  const long grid = (tasks + 2 * WARPS - 1) / (2 * WARPS);
  const size_t smem = (size_t)WARPS * TriDualSmem<T, NP, LEFT>::per_warp * sizeof(T);
  auto kern = tri_solve_dual_kernel<T, NP, LEFT, OP, WARPS, STRIDED, VEC16>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  // CTAs resident on the whole GPU = how far ahead the kernel prefetches into L2 (variant 46: no prefetch, A/B)
  const bool pf = h->variant_override != 46 && (NP > 16 || sizeof(T) == 4);  // measured: see prefetch_solve_task_l2
  const int ahead = pf ? h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, smem, 2) : 0;
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(vec, alpha, A, lda, B, ldb, batchCount, slabs, ahead);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// side L, strided, everything 16-byte aligned: 16-byte accesses only (kernels/trsm_left_vec.cuh)
template <typename T>
static bool tri_left_vec_ok(int k, const BatchRef<const T, true> &A, int lda, const BatchRef<T, true> &B, int ldb) {
  constexpr int VW = 16 / (int)sizeof(T);
  return k % VW == 0 && lda % VW == 0 && ldb % VW == 0 && A.stride % VW == 0 && B.stride % VW == 0 &&
         ((unsigned long long)A.base | (unsigned long long)B.base) % 16 == 0;
}
template <typename T>
static bool tri_left_vec_ok(int, const BatchRef<const T, false> &, int, const BatchRef<T, false> &, int) { return false; }

template <typename T, int NP, int OP>
static int launch_tri_left_vec(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, true> A, int lda,
                               BatchRef<T, true> B, int ldb, int batchCount) {
  constexpr int WARPS = TriLeftVecSmem<T, NP>::warps, MINB = TriLeftVecSmem<T, NP>::ctas_per_sm;
  const int slabs = (vec + 31) / 32;
  const long tasks = (long)batchCount * slabs;
  const long grid = (tasks + WARPS - 1) / WARPS;
  const size_t smem = (size_t)WARPS * TriLeftVecSmem<T, NP>::per_warp * sizeof(T);
  auto kern = tri_left_vec_kernel<T, NP, OP, WARPS, MINB>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  const int ahead = h->variant_override == 46 ? 0 : h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, smem, MINB);
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A.base, lda, A.stride, B.base, ldb, B.stride, batchCount, slabs,
                                                        ahead);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}
template <typename T, int NP, int OP>
static int launch_tri_left_vec(KBlasHandle *, const char *, int, int, T, BatchRef<const T, false>, int, BatchRef<T, false>, int, int) {
  return KBLAS_UnknownError;  // never selected: tri_left_vec_ok is false for pointer arrays
}

// side R, strided, 16-byte aligned factor: one vector per lane, factor staged with 16-byte cp.async (kernels/trsm_left_vec.cuh)
template <typename T>
static bool tri_right_vec_ok(int k, const BatchRef<const T, true> &A, int lda) {
  constexpr int VW = 16 / (int)sizeof(T);
  return k % VW == 0 && lda % VW == 0 && A.stride % VW == 0 && (unsigned long long)A.base % 16 == 0;
}
template <typename T>
static bool tri_right_vec_ok(int, const BatchRef<const T, false> &, int) { return false; }

template <typename T, int NP, int OP>
static int launch_tri_right_vec(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, true> A, int lda,
                                BatchRef<T, true> B, int ldb, int batchCount) {
  constexpr int WARPS = TriRightVecSmem<T, NP>::warps, MINB = TriRightVecSmem<T, NP>::ctas_per_sm;
  const int slabs = (vec + 31) / 32;
  const long tasks = (long)batchCount * slabs;
  const long grid = (tasks + WARPS - 1) / WARPS;
  const size_t smem = (size_t)WARPS * TriRightVecSmem<T, NP>::per_warp * sizeof(T);
  auto kern = tri_right_vec_kernel<T, NP, OP, WARPS, MINB>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  const int ahead = h->variant_override == 46 ? 0 : h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, smem, MINB);
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A.base, lda, A.stride, B.base, ldb, B.stride, batchCount, slabs,
                                                        ahead);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}
template <typename T, int NP, int OP>
static int launch_tri_right_vec(KBlasHandle *, const char *, int, int, T, BatchRef<const T, false>, int, BatchRef<T, false>, int, int) {
  return KBLAS_UnknownError;
}

// k <= 16 and vec <= 16: register-resident, 2 / 4 problems per warp (kernels/trsm_reg.cuh)
template <typename T, int NP, int GP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_reg(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                          int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / GP;
  const long wtasks = ((long)batchCount + MPW - 1) / MPW;
  const long grid = (wtasks + WARPS - 1) / WARPS;
  const bool full = k == NP && vec == GP && (!LEFT || NP == GP);
  if (full) {
    auto kern = tri_solve_reg_kernel<T, NP, GP, LEFT, OP, WARPS, STRIDED, true>;
    const int ahead = h->variant_override == 46 ? 0 : h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, 0, 4);
    kern<<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, ahead);
  } else {
    auto kern = tri_solve_reg_kernel<T, NP, GP, LEFT, OP, WARPS, STRIDED, false>;
    const int ahead = h->variant_override == 46 ? 0 : h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, 0, 4);
    kern<<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, ahead);
  }
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// vec <= 16: several matrices per warp (kernels/trsm_small.cuh, packed variant)
template <typename T, int NP, int GP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_packed(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                             int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / GP;
  const long wtasks = ((long)batchCount + MPW - 1) / MPW;
  const long grid = (wtasks + WARPS - 1) / WARPS;
  tri_solve_packed_kernel<T, NP, GP, LEFT, OP, WARPS, STRIDED>
      <<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// k > 32: blocked substitution, one warp per (matrix, 32-vector slab), or 2 / 4 matrices per warp when
// there are at most 16 / 8 right-hand-side vectors (kernels/trsm_blocked.cuh)
template <typename T, bool LEFT, int OP, int GP, bool STRIDED>
static int launch_tri_blocked_gp(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                                 int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int MPW = 32 / GP;
  constexpr size_t per_warp = TriBlockedSmem<T, GP>::per_warp * sizeof(T);
  constexpr int WARPS = (per_warp * 4 <= 70000) ? 4 : (per_warp * 2 <= 70000) ? 2 : 1;
  const int slabs = (GP == 32) ? (vec + 31) / 32 : 1;
  const long tasks = (((long)batchCount + MPW - 1) / MPW) * slabs;
  const long grid = (tasks + WARPS - 1) / WARPS;
  auto kern = tri_solve_blocked_kernel<T, LEFT, OP, GP, WARPS, STRIDED>;
  const size_t smem = per_warp * WARPS;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, slabs, h->tri_flags);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// k > 32, side R, fp64: the off-diagonal products on DMMA (kernels/trsm_mma.cuh)
template <int OP, int GP, bool STRIDED>
static int launch_tri_mma(KBlasHandle *h, const char *name, int k, int vec, double alpha, BatchRef<const double, STRIDED> A, int lda,
                          BatchRef<double, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / GP;
  const int slabs = (GP == 32) ? (vec + 31) / 32 : 1;
  const long tasks = (((long)batchCount + MPW - 1) / MPW) * slabs;
  const long grid = (tasks + WARPS - 1) / WARPS;
  const size_t smem = (size_t)WARPS * TriMmaSmem::per_warp * sizeof(double);
  auto kern = tri_solve_mma_kernel<OP, GP, WARPS, STRIDED>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, slabs);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, bool LEFT, int OP, bool STRIDED>
static int launch_tri_blocked(KBlasHandle *h, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                              BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  if constexpr (sizeof(T) == 8 && !LEFT) {
    // variant 44 keeps the FMA kernels (A/B); Upper / Unit factors are staged by the generic kernels only
    if (h->variant_override != 44 && vec > 8 && h->tri_flags == 0) {
      if (vec <= 16) return launch_tri_mma<OP, 16, STRIDED>(h, "tri_mma<GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
      return launch_tri_mma<OP, 32, STRIDED>(h, "tri_mma<GP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
    }
  }
  if (vec <= 8)
    return launch_tri_blocked_gp<T, LEFT, OP, 8, STRIDED>(h, "tri_blocked<GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (vec <= 16)
    return launch_tri_blocked_gp<T, LEFT, OP, 16, STRIDED>(h, "tri_blocked<GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  return launch_tri_blocked_gp<T, LEFT, OP, 32, STRIDED>(h, "tri_blocked<GP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
}

template <typename T, bool LEFT, int OP, bool STRIDED>
static int tri_small_np(KBlasHandle *h, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                        BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  if (h->tri_flags != 0) {  // Upper storage / unit diagonal: only the generic kernel stages those (kernels/trsm_small.cuh)
    if (k <= 8) return launch_tri_small<T, 8, LEFT, OP, STRIDED>(h, "tri_small<NP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 16) return launch_tri_small<T, 16, LEFT, OP, STRIDED>(h, "tri_small<NP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 24) return launch_tri_small<T, 24, LEFT, OP, STRIDED>(h, "tri_small<NP=24>", k, vec, alpha, A, lda, B, ldb, batchCount);
    return launch_tri_small<T, 32, LEFT, OP, STRIDED>(h, "tri_small<NP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
  }
  // measured (B200, 2^20 problems, side L): dual vs one-vector kernel  fp64 k=32 4.7-5.1 vs 5.1-5.4 ms, k=24 3.0-3.4 vs
  // 3.6-3.9; fp32 k=24 2.07 vs 2.35 but k=32 2.96-3.07 vs 2.77-2.83 -> fp32 k=32 side L stays on the older kernel
  if constexpr (LEFT) {
    // 16-byte kernel vs the element-wise ones, measured on B200 (2^20 problems, vec = k, ms; profiles/r02_trsm_left_vec.txt):
    //   fp64 k=32  trsm N 3.93 vs 4.70, T 3.78 vs 4.81, potrs 5.65 vs 6.47     fp32 k=32  2.18 vs 2.77, 1.82 vs 2.68, 2.72 vs 3.62
    //   fp64 k=24  2.78 vs 2.93, 2.58 vs 3.01, potrs 3.92 vs 3.51 (stays)      fp32 k=24  1.27 vs 1.65, 1.24 vs 1.68, 1.77 vs 2.18
    //   k=16: only the fused potrs wins (16 vectors: fp64 1.59 vs 2.54, fp32 0.91 vs 1.49); k=8 and few vectors (idle
    //   lanes) stay on the register kernels.  Variant 40 forces it wherever it is eligible, 41 switches it off.
    const int vo = h->variant_override;
    if constexpr (OP == TRI_BOTH && STRIDED) {
      // fused side-L potrs, k = 24 / 32: two vectors per lane (half the factor broadcasts of the one-vector kernel, which is
      // bound by the shared-memory pipe here) with the slab of B staged in 16-byte chunks.  Measured (B200, 2^20 problems, ms,
      // against what it replaces): fp64 k = 24 3.79 vs 4.21, fp32 k = 32 2.80 vs 3.26, k = 24 1.92 vs 2.03; fp64 k = 32 6.47 vs
      // 6.28 -- 34 KB of shared memory per warp leave 6 warps per SM there, so that one stays on the one-vector kernel.
      // Variant 53 switches it off.
      if ((k == 24 || (k == 32 && sizeof(T) == 4)) && vec > 16 && vo != 53 && vo != 40 && vo != 41 && tri_left_vec_ok<T>(k, A, lda, B, ldb)) {
        if (k == 24) return launch_tri_dual<T, 24, LEFT, OP, STRIDED, true>(h, "tri_dual16<NP=24>", vec, alpha, A, lda, B, ldb, batchCount);
        return launch_tri_dual<T, 32, LEFT, OP, STRIDED, true>(h, "tri_dual16<NP=32>", vec, alpha, A, lda, B, ldb, batchCount);
      }
    }
    const bool pays = k > 16 ? (vec > 16 && !(sizeof(T) == 8 && OP == TRI_BOTH && k <= 24)) : (k > 8 && vec > 8 && OP == TRI_BOTH);
    const bool want = vo == 40 || (vo != 41 && pays);
    if (want && tri_left_vec_ok<T>(k, A, lda, B, ldb)) {
      if (k <= 8) return launch_tri_left_vec<T, 8, OP>(h, "tri_left_vec<NP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
      if (k <= 16) return launch_tri_left_vec<T, 16, OP>(h, "tri_left_vec<NP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
      if (k <= 24) return launch_tri_left_vec<T, 24, OP>(h, "tri_left_vec<NP=24>", k, vec, alpha, A, lda, B, ldb, batchCount);
      return launch_tri_left_vec<T, 32, OP>(h, "tri_left_vec<NP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
    }
  }
  if constexpr (!LEFT && OP == TRI_BACKWARD && sizeof(T) == 8) {
    // side R, X L = alpha B in fp64 with a full 32-row slab: one vector per lane at 16 warps per SM beats the two-vector
    // kernel (2^20 problems, k = vec = 32: 3.52 vs 4.06 ms = 0.94 vs 0.81 of the HBM roofline).  The forward form of the
    // same kernel spills at 128 registers and loses (4.12 vs 3.91), fp32 loses (2.49 vs 2.38), fewer than 32 rows leave
    // lanes idle (k = vec = 24: 2.47 vs 2.14): those stay where they were.  Variant 43 switches it off.
    if (k > 24 && vec >= 32 && h->variant_override != 43 && tri_right_vec_ok<T>(k, A, lda))
      return launch_tri_right_vec<T, 32, OP>(h, "tri_right_vec<NP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
  }
  const bool dual_ok = !LEFT || !(sizeof(T) == 4 && k == 32);
  if (dual_ok) {
    // k = 16 with <= 16 vectors leaves the second vector of every lane idle and still wins on side R (measured, ms per
    // 2^20: fp64 potrs 1.07 vs 1.6-2.7, trsm R 0.95 vs 1.01-1.08; fp32 trsm R 0.57-0.59 vs 0.58-0.64); side L and fp32
    // potrs stay on the register / packed kernels (dual: 1.4-1.5 vs 1.13-1.2; 0.76 vs 0.73)
    const bool dual16 = vec > 16 || (!LEFT && (sizeof(T) == 8 || OP != TRI_BOTH));
    if (k == 16 && dual16) return launch_tri_dual<T, 16, LEFT, OP, STRIDED>(h, "tri_dual<NP=16>", vec, alpha, A, lda, B, ldb, batchCount);
    if (k == 24) return launch_tri_dual<T, 24, LEFT, OP, STRIDED>(h, "tri_dual<NP=24>", vec, alpha, A, lda, B, ldb, batchCount);
    if (k == 32) return launch_tri_dual<T, 32, LEFT, OP, STRIDED>(h, "tri_dual<NP=32>", vec, alpha, A, lda, B, ldb, batchCount);
  }
  // few right-hand sides and a small factor: register kernel, 4 / 2 problems per warp
  // (measured: the shared-memory packed kernel stays ahead only for fp32, side R, 8 < k <= 16)
  if (!(sizeof(T) == 4 && !LEFT && k > 8)) {
    if (k <= 8 && vec <= 8) return launch_tri_reg<T, 8, 8, LEFT, OP, STRIDED>(h, "tri_reg<NP=8,GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 8 && vec <= 16) return launch_tri_reg<T, 8, 16, LEFT, OP, STRIDED>(h, "tri_reg<NP=8,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 16 && vec <= 16) return launch_tri_reg<T, 16, 16, LEFT, OP, STRIDED>(h, "tri_reg<NP=16,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  }
  if (k <= 8 && vec <= 8) return launch_tri_packed<T, 8, 8, LEFT, OP, STRIDED>(h, "tri_packed<NP=8,GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (k <= 8 && vec <= 16) return launch_tri_packed<T, 8, 16, LEFT, OP, STRIDED>(h, "tri_packed<NP=8,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  // (side L reads B by rows, one lane per row: the lane group must cover all k rows)
  if constexpr (!LEFT) {
    if (k <= 16 && vec <= 8) return launch_tri_packed<T, 16, 8, LEFT, OP, STRIDED>(h, "tri_packed<NP=16,GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
  }
  if (k <= 16 && vec <= 16) return launch_tri_packed<T, 16, 16, LEFT, OP, STRIDED>(h, "tri_packed<NP=16,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (k <= 8) return launch_tri_small<T, 8, LEFT, OP, STRIDED>(h, "tri_small<NP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (k <= 16) return launch_tri_small<T, 16, LEFT, OP, STRIDED>(h, "tri_small<NP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (k <= 24) return launch_tri_small<T, 24, LEFT, OP, STRIDED>(h, "tri_small<NP=24>", k, vec, alpha, A, lda, B, ldb, batchCount);
  return launch_tri_small<T, 32, LEFT, OP, STRIDED>(h, "tri_small<NP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
}

// one side of tri_solve_core: LEFT fixed at compile time, op dispatched here
template <typename T, bool STRIDED, bool LEFT>
int tri_solve_side(KBlasHandle *h, int op, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                   BatchRef<T, STRIDED> B, int ldb, int batchCount) {
#define KX_TRI(O_)                                                                                   \
  (k <= 32 ? tri_small_np<T, LEFT, O_, STRIDED>(h, k, vec, alpha, A, lda, B, ldb, batchCount)         \
           : launch_tri_blocked<T, LEFT, O_, STRIDED>(h, k, vec, alpha, A, lda, B, ldb, batchCount))
  if (op == TRI_FORWARD) return KX_TRI(TRI_FORWARD);
  if (op == TRI_BACKWARD) return KX_TRI(TRI_BACKWARD);
  return KX_TRI(TRI_BOTH);
#undef KX_TRI
}

}  // namespace kblasx
