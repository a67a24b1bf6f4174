// potrs_batch.cu -- kblas_potrs_batch: X (L L^T) = B (side R, the reference's only form) or (L L^T) X = B (side L, extension).
//
// Counterpart of reference src/batch_triangular/Xpotrs_batch.cu:42-164 and
// Xpotrs_batch_drivers.cuh:32-174.  The reference composes 4 TRSM + 2 GEMM launches (its
// fused register kernels are compiled out, drivers.cuh:44-78); here the forward and the
// backward substitution run back to back on register-resident rows of B in one launch.
#include "kblas.h"
#include "kblas_common.h"
#include "tri_batch.h"

namespace kblasx {

template <typename T, bool STRIDED>
int potrs_batch_core(KBlasHandle *h, char side, char uplo, int m, int n, BatchRef<const T, STRIDED> A, int lda,
                     BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  // uplo = Upper (A = U^T U) is KBLAS_NotImplemented in the reference (drivers.cuh:40-43); here L = U^T is staged by the
  // generic solve kernels (extension)
  // side L -- A X = B, A = L L^T of order m, B is m x n -- is KBLAS_NotImplemented in the reference (drivers.cuh:40-43);
  // here it is the same fused forward + backward substitution with the factor acting from the left (SURVEY.md §8(f)3,
  // documented extension): L Y = B, L^T X = Y.
  const bool left = (side == KBLAS_Left);
  if (!left && side != KBLAS_Right) return KBLAS_NotImplemented;
  // The reference splits n = n1 + n2 and hands n1 = CLOSEST_REG_SIZE(1) = 0 columns to TRSM,
  // which answers KBLAS_NotImplemented (drivers.cuh:85-98, Xtrsm_batch_drivers.cuh:267-270).
  // A 1 x 1 factor is a perfectly good problem: solved here (documented deviation).
  if ((left ? m : n) <= 0) return KBLAS_NotImplemented;
  h->tri_flags = (uplo == KBLAS_Upper) ? TRI_FLAG_UPPER : 0;
  const int rc = tri_solve_core<T, STRIDED>(h, left, TRI_BOTH, m, n, T(1), A, lda, B, ldb, batchCount);
  h->tri_flags = 0;
  return rc;
}

#define KX_INST(T, S)                                                                                      \
  template int potrs_batch_core<T, S>(KBlasHandle *, char, char, int, int, BatchRef<const T, S>, int,      \
                                      BatchRef<T, S>, int, int);
KX_INST(float, true)
KX_INST(float, false)
KX_INST(double, true)
KX_INST(double, false)
#undef KX_INST

static int potrs_ws_check(KBlasHandle *h, bool strided, int m, int n, int batchCount) {
  KBlasWorkspaceState need;
  potrs_batch_wsquery_core(strided, m, n, batchCount, &need);  // reference Xpotrs_batch.cu:50-56
  return need.isSufficient(&h->work_space.allocated_ws_state) ? KBLAS_Success : KBLAS_InsufficientWorkspace;
}

template <typename T>
int potrs_batch_strided(KBlasHandle *h, char side, char uplo, int m, int n, const T *A, int lda, long strideA,
                               T *B, int ldb, long strideB, int batchCount) {
  if (potrs_ws_check(h, true, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<const T, true> a = {A, strideA};
  BatchRef<T, true> b = {B, strideB};
  return potrs_batch_core<T, true>(h, side, uplo, m, n, a, lda, b, ldb, batchCount);
}

template <typename T>
int potrs_batch_ptrs(KBlasHandle *h, char side, char uplo, int m, int n, const T **A, long a_off, int lda, T **B,
                     long b_off, int ldb, int batchCount) {
  if (potrs_ws_check(h, false, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<const T, false> a = {A, a_off};
  BatchRef<T, false> b = {B, b_off};
  return potrs_batch_core<T, false>(h, side, uplo, m, n, a, lda, b, ldb, batchCount);
}

}  // namespace kblasx

// ---- public API (reference Xpotrs_batch.cu:60-98 pointer array, 130-164 strided)
#define KX_POTRS_API(P, T)                                                                                     \
  int kblas_potrs_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n, const T **A,     \
                        int lda, T **B, int ldb, int batchCount) {                                             \
    return kblasx::potrs_batch_ptrs<T>(handle, side, uplo, m, n, A, 0, lda, B, 0, ldb, batchCount);                  \
  }                                                                                                            \
  int kblas_potrs_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n, const T *A,      \
                        int lda, long strideA, T *B, int ldb, long strideB, int batchCount) {                  \
    return kblasx::potrs_batch_strided<T>(handle, side, uplo, m, n, A, lda, strideA, B, ldb, strideB,          \
                                          batchCount);                                                         \
  }                                                                                                            \
  extern "C" int kblas##P##potrs_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,   \
                                       const T **A, int lda, T **B, int ldb, int batchCount) {                 \
    return kblasx::potrs_batch_ptrs<T>(handle, side, uplo, m, n, A, 0, lda, B, 0, ldb, batchCount);                  \
  }                                                                                                            \
  extern "C" int kblas##P##potrs_batch_strided(kblasHandle_t handle, char side, char uplo, const int m,        \
                                               const int n, const T *A, int lda, long strideA, T *B, int ldb,  \
                                               long strideB, int batchCount) {                                 \
    return kblasx::potrs_batch_strided<T>(handle, side, uplo, m, n, A, lda, strideA, B, ldb, strideB,          \
                                          batchCount);                                                         \
  }
// internal C++ entry points with sub-matrix offsets (reference Xpotrs_batch.cu:42-58, 104-127; src/Xblas_core.ch:264-277)
#define KX_POTRS_OFFSET_API(T)                                                                                  \
  int Xpotrs_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n, const T **A,    \
                          int A_row_off, int A_col_off, int lda, T **B, int B_row_off, int B_col_off, int ldb,  \
                          int batchCount) {                                                                     \
    return kblasx::potrs_batch_ptrs<T>(handle, side, uplo, m, n, A, A_row_off + (long)A_col_off * lda, lda, B,  \
                                       B_row_off + (long)B_col_off * ldb, ldb, batchCount);                     \
  }                                                                                                             \
  int Xpotrs_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n, const T *A,     \
                          int A_row_off, int A_col_off, int lda, long strideA, T *B, int B_row_off,             \
                          int B_col_off, int ldb, long strideB, int batchCount) {                               \
    return kblasx::potrs_batch_strided<T>(handle, side, uplo, m, n, A + A_row_off + (long)A_col_off * lda, lda, \
                                          strideA, B + B_row_off + (long)B_col_off * ldb, ldb, strideB,         \
                                          batchCount);                                                          \
  }
KX_POTRS_OFFSET_API(float)
KX_POTRS_OFFSET_API(double)

KX_POTRS_API(S, float)
KX_POTRS_API(D, double)
