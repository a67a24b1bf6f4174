// trsm_batch.cu -- kblas_trsm_batch (+ the tri-solve entry shared with potrs/posv).
//
// Counterpart of reference src/batch_triangular/Xtrsm_batch.cu:42-257 (entry points,
// workspace check) and Xtrsm_batch_drivers.cuh:54-272 (driver: kernel table + recursion
// through cuBLAS batched GEMM).  Here: one launch per call; the kernel selection is in
// trsm_dispatch.cuh, instantiated by trsm_inst_*.cu.
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/common.cuh"
#include "tri_batch.h"

namespace kblasx {

// Solve with the k x k lower factor in A; `left` selects which side of B it acts on and `op`
// forward / backward / both (see kernels/trsm_small.cuh).  The per-side bodies live in trsm_inst_*.cu.
template <typename T, bool STRIDED, bool LEFT>
int tri_solve_side(KBlasHandle *h, int op, int k, int vec, T alpha, BatchRef<const T, STRIDED> A, int lda,
                   BatchRef<T, STRIDED> B, int ldb, int batchCount);

template <typename T, bool STRIDED>
int tri_solve_core(KBlasHandle *h, bool left, int op, int m, int n, T alpha, BatchRef<const T, STRIDED> A, int lda,
                   BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  const int k = left ? m : n, vec = left ? n : m;
  if (batchCount <= 0) {
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);  // reference: empty grid
  }
  if (vec <= 0) return KBLAS_Success;
  return left ? tri_solve_side<T, STRIDED, true>(h, op, k, vec, alpha, A, lda, B, ldb, batchCount)
              : tri_solve_side<T, STRIDED, false>(h, op, k, vec, alpha, A, lda, B, ldb, batchCount);
}

#define KX_INST(T, S)                                                                                      \
  extern template int tri_solve_side<T, S, true>(KBlasHandle *, int, int, int, T, BatchRef<const T, S>, int, \
                                                 BatchRef<T, S>, int, int);                                \
  extern template int tri_solve_side<T, S, false>(KBlasHandle *, int, int, int, T, BatchRef<const T, S>, int, \
                                                  BatchRef<T, S>, int, int);                               \
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
  // Upper and Unit are KBLAS_NotImplemented in the reference (Xtrsm_batch_drivers.cuh:64-67); here they are served by the
  // generic kernels (documented extension): an upper factor U is staged as L = U^T, so op(U) = op'(L)
  const bool upper = (uplo == KBLAS_Upper), unit = (diag == KBLAS_Unit);
  const bool left = (side == KBLAS_Left);
  if (!left && side != KBLAS_Right) return KBLAS_NotImplemented;
  // the reference falls through to "should not reach this" when the triangular dimension is 0
  if ((left ? m : n) <= 0) return KBLAS_NotImplemented;  // drivers.cuh:267-270
  const bool notrans = (trans == KBLAS_NoTrans) != upper;
  // forward: (R, T) and (L, N); backward: (R, N) and (L, T)
  const int op = (left == notrans) ? TRI_FORWARD : TRI_BACKWARD;
  h->tri_flags = (upper ? TRI_FLAG_UPPER : 0) | (unit ? TRI_FLAG_UNIT : 0);
  const int rc = tri_solve_core<T, STRIDED>(h, left, op, m, n, alpha, A, lda, B, ldb, batchCount);
  h->tri_flags = 0;
  return rc;
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
                    long a_off, int lda, T **B, long b_off, int ldb, int batchCount) {
  if (trsm_ws_check(h, false, side, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<const T, false> a = {A, a_off};
  BatchRef<T, false> b = {B, b_off};
  return trsm_batch_core<T, false>(h, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb, batchCount);
}

}  // namespace kblasx

// ---- public API (reference Xtrsm_batch.cu:60-112 pointer array, 218-257 strided)
#define KX_TRSM_API(P, T)                                                                                      \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, const int m,         \
                       const int n, const T alpha, const T **A, int lda, T **B, int ldb, int batchCount) {     \
    return kblasx::trsm_batch_ptrs<T>(handle, side, uplo, trans, diag, m, n, alpha, A, 0, lda, B, 0, ldb,      \
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
    return kblasx::trsm_batch_ptrs<T>(handle, side, uplo, trans, diag, m, n, alpha, A, 0, lda, B, 0, ldb,      \
                                      batchCount);                                                             \
  }                                                                                                            \
  extern "C" int kblas##P##trsm_batch_strided(kblasHandle_t handle, char side, char uplo, char trans,          \
                                              char diag, const int m, const int n, const T alpha, const T *A,  \
                                              int lda, long strideA, T *B, int ldb, long strideB,              \
                                              int batchCount) {                                                \
    return kblasx::trsm_batch_strided<T>(handle, side, uplo, trans, diag, m, n, alpha, A, lda, strideA, B,     \
                                         ldb, strideB, batchCount);                                            \
  }
// uniform-size internal C++ entry points with sub-matrix offsets: called by the reference's own test program
// (testing/batch_triangular/test_Xtrsm_batch.cpp:317) and by trtri (reference Xtrsm_batch.cu:42-58 pointer array --
// strideA / strideB ignored there too -- and 189-216 strided; src/Xblas_core.ch:194-213)
#define KX_TRSM_OFFSET_API(T)                                                                                   \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int m, int n, T alpha,     \
                  T **A, int A_row_off, int A_col_off, int lda, long /*strideA*/, T **B, int B_row_off,         \
                  int B_col_off, int ldb, long /*strideB*/, int batchCount) {                                   \
    return kblasx::trsm_batch_ptrs<T>(handle, side, uplo, trans, diag, m, n, alpha, (const T **)A,              \
                                      A_row_off + (long)A_col_off * lda, lda, B,                                \
                                      B_row_off + (long)B_col_off * ldb, ldb, batchCount);                      \
  }                                                                                                             \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int m, int n, T alpha,     \
                  T *A, int A_row_off, int A_col_off, int lda, long strideA, T *B, int B_row_off,               \
                  int B_col_off, int ldb, long strideB, int batchCount) {                                       \
    return kblasx::trsm_batch_strided<T>(handle, side, uplo, trans, diag, m, n, alpha,                          \
                                         A + A_row_off + (long)A_col_off * lda, lda, strideA,                   \
                                         B + B_row_off + (long)B_col_off * ldb, ldb, strideB, batchCount);      \
  }
KX_TRSM_OFFSET_API(float)
KX_TRSM_OFFSET_API(double)

// Non-uniform batches (per-matrix m, n, lda, ldb arrays): trsm_nonuniform.cu

KX_TRSM_API(S, float)
KX_TRSM_API(D, double)
