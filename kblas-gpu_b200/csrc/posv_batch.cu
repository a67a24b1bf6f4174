// posv_batch.cu -- kblas_posv_batch: factor A = L L^T, then solve X (L L^T) = B (side R only).
//
// Counterpart of reference src/batch_triangular/Xposv_batch.cu:42-178 and
// Xposv_batch_drivers.cuh:32-117 (POTRF then POTRS, its fused kernels are compiled out,
// drivers.cuh:45-79).  Two launches here (factor, fused forward+backward solve) against
// the reference's ~10 (n <= 32) to ~217 (n = 256, pointer array).
#include "kblas.h"
#include "kblas_common.h"
#include "potrf_batch.h"
#include "tri_batch.h"

namespace kblasx {

template <typename T, bool STRIDED>
static int posv_batch_core(KBlasHandle *h, char side, char uplo, int m, int n, BatchRef<T, STRIDED> A, int lda,
                           BatchRef<T, STRIDED> B, int ldb, int batchCount, int *info) {
  // uplo = Upper is KBLAS_NotImplemented in the reference (drivers.cuh:41-44); potrf and potrs implement it here
  // side L (A of order m, A X = B) is an extension: the reference returns KBLAS_NotImplemented (drivers.cuh:41-44)
  if (side != KBLAS_Left && side != KBLAS_Right) return KBLAS_NotImplemented;
  check_ret_error((potrf_batch_core<T, STRIDED>(h, uplo, side == KBLAS_Left ? m : n, A, lda, batchCount, info)));
  BatchRef<const T, STRIDED> Ac;
  Ac.base = A.base;
  Ac.stride = A.stride;
  check_ret_error((potrs_batch_core<T, STRIDED>(h, side, uplo, m, n, Ac, lda, B, ldb, batchCount)));
  return KBLAS_Success;
}

static int posv_ws_check(KBlasHandle *h, bool strided, char side, int m, int n, int batchCount) {
  KBlasWorkspaceState need;
  posv_batch_wsquery_core(strided, m, n, side, batchCount, &need);  // reference Xposv_batch.cu:50-56
  return need.isSufficient(&h->work_space.allocated_ws_state) ? KBLAS_Success : KBLAS_InsufficientWorkspace;
}

template <typename T>
static int posv_batch_strided(KBlasHandle *h, char side, char uplo, int m, int n, T *A, int lda, long strideA, T *B,
                              int ldb, long strideB, int batchCount, int *info) {
  if (posv_ws_check(h, true, side, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<T, true> a = {A, strideA};
  BatchRef<T, true> b = {B, strideB};
  return posv_batch_core<T, true>(h, side, uplo, m, n, a, lda, b, ldb, batchCount, info);
}

template <typename T>
static int posv_batch_ptrs(KBlasHandle *h, char side, char uplo, int m, int n, T **A, long a_off, int lda, T **B,
                           long b_off, int ldb, int batchCount, int *info) {
  if (posv_ws_check(h, false, side, m, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<T, false> a = {A, a_off};
  BatchRef<T, false> b = {B, b_off};
  return posv_batch_core<T, false>(h, side, uplo, m, n, a, lda, b, ldb, batchCount, info);
}

}  // namespace kblasx

// ---- public API (reference Xposv_batch.cu:69-107 pointer array, 140-178 strided)
#define KX_POSV_API(P, T)                                                                                      \
  int kblas_posv_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n, T **A, int lda,   \
                       T **B, int ldb, int batchCount, int *info_array) {                                      \
    return kblasx::posv_batch_ptrs<T>(handle, side, uplo, m, n, A, 0, lda, B, 0, ldb, batchCount, info_array);       \
  }                                                                                                            \
  int kblas_posv_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n, T *A, int lda,    \
                       long strideA, T *B, int ldb, long strideB, int batchCount, int *info_array) {           \
    return kblasx::posv_batch_strided<T>(handle, side, uplo, m, n, A, lda, strideA, B, ldb, strideB,           \
                                         batchCount, info_array);                                              \
  }                                                                                                            \
  extern "C" int kblas##P##posv_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,    \
                                      T **A, int lda, T **B, int ldb, int batchCount, int *info_array) {       \
    return kblasx::posv_batch_ptrs<T>(handle, side, uplo, m, n, A, 0, lda, B, 0, ldb, batchCount, info_array);       \
  }                                                                                                            \
  extern "C" int kblas##P##posv_batch_strided(kblasHandle_t handle, char side, char uplo, const int m,         \
                                              const int n, T *A, int lda, long strideA, T *B, int ldb,         \
                                              long strideB, int batchCount, int *info_array) {                 \
    return kblasx::posv_batch_strided<T>(handle, side, uplo, m, n, A, lda, strideA, B, ldb, strideB,           \
                                         batchCount, info_array);                                              \
  }
// internal C++ entry points with sub-matrix offsets (reference Xposv_batch.cu:42-67, 112-138)
#define KX_POSV_OFFSET_API(T)                                                                                   \
  int Xposv_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n, T **A,           \
                         int A_row_off, int A_col_off, int lda, T **B, int B_row_off, int B_col_off, int ldb,   \
                         int batchCount, int *info_array) {                                                     \
    return kblasx::posv_batch_ptrs<T>(handle, side, uplo, m, n, A, A_row_off + (long)A_col_off * lda, lda, B,   \
                                      B_row_off + (long)B_col_off * ldb, ldb, batchCount, info_array);          \
  }                                                                                                             \
  int Xposv_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n, T *A,            \
                         int A_row_off, int A_col_off, int lda, long strideA, T *B, int B_row_off,              \
                         int B_col_off, int ldb, long strideB, int batchCount, int *info_array) {               \
    return kblasx::posv_batch_strided<T>(handle, side, uplo, m, n, A + A_row_off + (long)A_col_off * lda, lda,  \
                                         strideA, B + B_row_off + (long)B_col_off * ldb, ldb, strideB,          \
                                         batchCount, info_array);                                               \
  }
KX_POSV_OFFSET_API(float)
KX_POSV_OFFSET_API(double)

KX_POSV_API(S, float)
KX_POSV_API(D, double)
