// trsm_batch.cu -- kblas_trsm_batch (+ the tri-solve dispatch shared with potrs/posv).
//
// Counterpart of reference src/batch_triangular/Xtrsm_batch.cu:42-257 (entry points,
// workspace check) and Xtrsm_batch_drivers.cuh:54-272 (driver: kernel table + recursion
// through cuBLAS batched GEMM).  Here: one launch per call.
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/trsm_small.cuh"
#include "kernels/trsm_blocked.cuh"
#include "kernels/trsm_reg.cuh"
#include "kernels/trsm_bcast.cuh"
#include "kernels/trsm_dual.cuh"
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
  // per instantiation AND per device: the attribute belongs to the device's context (one process may drive
  // several GPUs, one handle each, as the reference harness does)
  static bool attr_set[64] = {};
  const int dev = (h->device_id >= 0 && h->device_id < 64) ? h->device_id : 0;
  if (!attr_set[dev]) {
    check_error_ret(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    KBLAS_CUDA_Error);
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, slabs);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// full NP x NP factor: two vectors per lane, two problems per warp, cp.async staging (kernels/trsm_dual.cuh)
template <typename T, int NP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_dual(KBlasHandle *h, const char *name, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                           BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = LEFT ? 2 : 4;  // side L carries a transpose tile per problem: 2-warp CTAs keep 3 CTAs per SM
  const int slabs = (vec + 31) / 32;
  const long tasks = (long)batchCount * slabs;
  const long grid = (tasks + 2 * WARPS - 1) / (2 * WARPS);
  const size_t smem = (size_t)WARPS * TriDualSmem<T, NP, LEFT>::per_warp * sizeof(T);
  auto kern = tri_solve_dual_kernel<T, NP, LEFT, OP, WARPS, STRIDED>;
  // per instantiation AND per device: the attribute belongs to the device's context (one process may drive
  // several GPUs, one handle each, as the reference harness does)
  static bool attr_set[64] = {};
  const int dev = (h->device_id >= 0 && h->device_id < 64) ? h->device_id : 0;
  if (!attr_set[dev]) {
    check_error_ret(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), KBLAS_CUDA_Error);
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(vec, alpha, A, lda, B, ldb, batchCount, slabs);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// k <= 16 and vec <= 16: register-resident, 2 / 4 problems per warp (kernels/trsm_reg.cuh)
template <typename T, int NP, int GP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_reg(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                          int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / GP;
  const long wtasks = ((long)batchCount + MPW - 1) / MPW;
  const long grid = (wtasks + WARPS - 1) / WARPS;
  tri_solve_reg_kernel<T, NP, GP, LEFT, OP, WARPS, STRIDED>
      <<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// k <= 16 and vec <= 16: factor read as L1-broadcast vector loads, 2 / 4 problems per warp (kernels/trsm_bcast.cuh)
template <typename T, int NP, int GP, bool LEFT, int OP, bool STRIDED>
static int launch_tri_bcast(KBlasHandle *h, const char *name, int k, int vec, T alpha, BatchRef<const T, STRIDED> A,
                            int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / GP;  // measured: 2 / 4 / 8 warps per CTA within +-3 %
  const long wtasks = ((long)batchCount + MPW - 1) / MPW;
  const long grid = (wtasks + WARPS - 1) / WARPS;
  tri_solve_bcast_kernel<T, NP, GP, LEFT, OP, WARPS, STRIDED>
      <<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount);
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
  // per instantiation AND per device: the attribute belongs to the device's context (one process may drive
  // several GPUs, one handle each, as the reference harness does)
  static bool attr_set[64] = {};
  const int dev = (h->device_id >= 0 && h->device_id < 64) ? h->device_id : 0;
  if (!attr_set[dev]) {
    check_error_ret(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), KBLAS_CUDA_Error);
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(k, vec, alpha, A, lda, B, ldb, batchCount, slabs);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, bool LEFT, int OP, bool STRIDED>
static int launch_tri_blocked(KBlasHandle *h, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                              BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  if (vec <= 8 && h->variant_override != 9)
    return launch_tri_blocked_gp<T, LEFT, OP, 8, STRIDED>(h, "tri_blocked<GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
  if (vec <= 16 && h->variant_override != 9)
    return launch_tri_blocked_gp<T, LEFT, OP, 16, STRIDED>(h, "tri_blocked<GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  return launch_tri_blocked_gp<T, LEFT, OP, 32, STRIDED>(h, "tri_blocked<GP=32>", k, vec, alpha, A, lda, B, ldb, batchCount);
}

template <typename T, bool LEFT, int OP, bool STRIDED>
static int tri_small_np(KBlasHandle *h, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                        BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  // few right-hand sides and a small factor: register kernel, 4 / 2 problems per warp
  // (measured: the shared-memory packed kernel stays ahead only for fp32, side R, 8 < k <= 16)
  // Full 8 / 16 columns with 16-byte aligned columns can read the factor as L1-broadcast vector loads
  // (kernels/trsm_bcast.cuh; pointer-array entries are checked in the kernel).  Measured on B200 (batch 2^20) it
  // loses by 3-10 % to the shuffle kernel at <= 80 registers and, for fp64 potrs k = 16 (1.6 ms), to the dual
  // kernel (1.07 ms): it is kept as variant 7 only.
  bool vec_ok = (k == 8 || k == 16) && ((size_t)lda * sizeof(T)) % 16 == 0;
  if constexpr (STRIDED) vec_ok = vec_ok && (reinterpret_cast<size_t>(A.base) % 16 == 0) && ((size_t)A.stride * sizeof(T)) % 16 == 0;
  if (vec_ok && h->variant_override == 7) {
    if (k <= 8 && vec <= 8) return launch_tri_bcast<T, 8, 8, LEFT, OP, STRIDED>(h, "tri_bcast<NP=8,GP=8>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 8 && vec <= 16) return launch_tri_bcast<T, 8, 16, LEFT, OP, STRIDED>(h, "tri_bcast<NP=8,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
    if (k <= 16 && vec <= 16) return launch_tri_bcast<T, 16, 16, LEFT, OP, STRIDED>(h, "tri_bcast<NP=16,GP=16>", k, vec, alpha, A, lda, B, ldb, batchCount);
  }
  // measured (B200, 2^20 problems, side L): dual vs one-vector kernel  fp64 k=32 4.7-5.1 vs 5.1-5.4 ms, k=24 3.0-3.4 vs
  // 3.6-3.9; fp32 k=24 2.07 vs 2.35 but k=32 2.96-3.07 vs 2.77-2.83 -> fp32 k=32 side L stays on the older kernel
  const bool dual_ok = !LEFT || !(sizeof(T) == 4 && k == 32);
  if (h->variant_override != 9 && h->variant_override != 8 && dual_ok && (!LEFT || h->variant_override != 5)) {  // 5 = side L on the older kernel
    // k = 16 with <= 16 vectors leaves the second vector of every lane idle and still wins on side R (measured, ms per
    // 2^20: fp64 potrs 1.07 vs 1.6-2.7, trsm R 0.95 vs 1.01-1.08; fp32 trsm R 0.57-0.59 vs 0.58-0.64); side L and fp32
    // potrs stay on the register / packed kernels (dual: 1.4-1.5 vs 1.13-1.2; 0.76 vs 0.73)
    const bool dual16 = vec > 16 || (!LEFT && (sizeof(T) == 8 || OP != TRI_BOTH));
    if (k == 16 && dual16) return launch_tri_dual<T, 16, LEFT, OP, STRIDED>(h, "tri_dual<NP=16>", vec, alpha, A, lda, B, ldb, batchCount);
    if (k == 24) return launch_tri_dual<T, 24, LEFT, OP, STRIDED>(h, "tri_dual<NP=24>", vec, alpha, A, lda, B, ldb, batchCount);
    if (k == 32) return launch_tri_dual<T, 32, LEFT, OP, STRIDED>(h, "tri_dual<NP=32>", vec, alpha, A, lda, B, ldb, batchCount);
  }
  // 8 = the register/shuffle kernel (A/B comparisons)
  if (h->variant_override != 9 && (h->variant_override == 6 || !(sizeof(T) == 4 && !LEFT && k > 8))) {
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

// Solve with the k x k lower factor in A; `left` selects which side of B it acts on and `op`
// forward / backward / both (see kernels/trsm_small.cuh).
template <typename T, bool STRIDED>
int tri_solve_core(KBlasHandle *h, bool left, int op, int m, int n, T alpha, BatchRef<const T, STRIDED> A, int lda,
                   BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  const int k = left ? m : n, vec = left ? n : m;
  if (batchCount <= 0) {
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);  // reference: empty grid
  }
  if (vec <= 0) return KBLAS_Success;
#define KX_TRI(L_, O_)                                                                              \
  (k <= 32 ? tri_small_np<T, L_, O_, STRIDED>(h, k, vec, alpha, A, lda, B, ldb, batchCount)          \
           : launch_tri_blocked<T, L_, O_, STRIDED>(h, k, vec, alpha, A, lda, B, ldb, batchCount))
  if (left) {
    if (op == TRI_FORWARD) return KX_TRI(true, TRI_FORWARD);
    if (op == TRI_BACKWARD) return KX_TRI(true, TRI_BACKWARD);
    return KX_TRI(true, TRI_BOTH);
  }
  if (op == TRI_FORWARD) return KX_TRI(false, TRI_FORWARD);
  if (op == TRI_BACKWARD) return KX_TRI(false, TRI_BACKWARD);
  return KX_TRI(false, TRI_BOTH);
#undef KX_TRI
}

#define KX_INST(T, S)                                                                                      \
  template int tri_solve_core<T, S>(KBlasHandle *, bool, int, int, int, T, BatchRef<const T, S>, int,      \
                                    BatchRef<T, S>, int, int);
KX_INST(float, true)
KX_INST(float, false)
KX_INST(double, true)
KX_INST(double, false)
#undef KX_INST

// Xtrsm_batch_core of the reference (Xtrsm_batch_drivers.cuh:54-272)
template <typename T, bool STRIDED>
static int trsm_batch_core(KBlasHandle *h, char side, char uplo, char trans, char diag, int m, int n, T alpha,
                           BatchRef<const T, STRIDED> A, int lda, BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  if (uplo == KBLAS_Upper || diag == KBLAS_Unit) {
    printf("(Upper | Unit) TRSM_BATCH is not implemented yet\n");  // reference drivers.cuh:65
    return KBLAS_NotImplemented;
  }
  const bool left = (side == KBLAS_Left);
  if (!left && side != KBLAS_Right) return KBLAS_NotImplemented;
  // the reference falls through to "should not reach this" when the triangular dimension is 0
  if ((left ? m : n) <= 0) return KBLAS_NotImplemented;  // drivers.cuh:267-270
  const bool notrans = (trans == KBLAS_NoTrans);
  // forward: (R, T) and (L, N); backward: (R, N) and (L, T)
  const int op = (left == notrans) ? TRI_FORWARD : TRI_BACKWARD;
  return tri_solve_core<T, STRIDED>(h, left, op, m, n, alpha, A, lda, B, ldb, batchCount);
}

static int trsm_ws_check(KBlasHandle *h, bool strided, char side, int m, int n, int batchCount) {
  KBlasWorkspaceState need;
  trsm_batch_wsquery_core(strided, batchCount, side, m, n, &need);  // reference Xtrsm_batch.cu:196-203
  return need.isSufficient(&h->work_space.allocated_ws_state) ? KBLAS_Success : KBLAS_InsufficientWorkspace;
}

template <typename T>
int trsm_batch_strided(KBlasHandle *h, char side, char uplo, char trans, char diag, int m, int n, T alpha,
                       const T *A, int lda, long strideA, T *B, int ldb, long strideB, int batchCount) {
  if (trsm_ws_check(h, true, side, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<const T, true> a = {A, strideA};
  BatchRef<T, true> b = {B, strideB};
  return trsm_batch_core<T, true>(h, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb, batchCount);
}

template <typename T>
int trsm_batch_ptrs(KBlasHandle *h, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T **A,
                    int lda, T **B, int ldb, int batchCount) {
  if (trsm_ws_check(h, false, side, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<const T, false> a = {A, 0};
  BatchRef<T, false> b = {B, 0};
  return trsm_batch_core<T, false>(h, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb, batchCount);
}

}  // namespace kblasx

// ---- public API (reference Xtrsm_batch.cu:60-112 pointer array, 218-257 strided)
#define KX_TRSM_API(P, T)                                                                                      \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, const int m,         \
                       const int n, const T alpha, const T **A, int lda, T **B, int ldb, int batchCount) {     \
    return kblasx::trsm_batch_ptrs<T>(handle, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb,            \
                                      batchCount);                                                             \
  }                                                                                                            \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, const int m,         \
                       const int n, const T alpha, const T *A, int lda, long strideA, T *B, int ldb,           \
                       long strideB, int batchCount) {                                                         \
    return kblasx::trsm_batch_strided<T>(handle, side, uplo, trans, diag, m, n, alpha, A, lda, strideA, B,     \
                                         ldb, strideB, batchCount);                                            \
  }                                                                                                            \
  extern "C" int kblas##P##trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag,       \
                                      const int m, const int n, const T alpha, const T **A, int lda, T **B,    \
                                      int ldb, int batchCount) {                                               \
    return kblasx::trsm_batch_ptrs<T>(handle, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb,            \
                                      batchCount);                                                             \
  }                                                                                                            \
  extern "C" int kblas##P##trsm_batch_strided(kblasHandle_t handle, char side, char uplo, char trans,          \
                                              char diag, const int m, const int n, const T alpha, const T *A,  \
                                              int lda, long strideA, T *B, int ldb, long strideB,              \
                                              int batchCount) {                                                \
    return kblasx::trsm_batch_strided<T>(handle, side, uplo, trans, diag, m, n, alpha, A, lda, strideA, B,     \
                                         ldb, strideB, batchCount);                                            \
  }
KX_TRSM_API(S, float)
KX_TRSM_API(D, double)
