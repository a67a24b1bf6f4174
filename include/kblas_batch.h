/*
 * kblas_batch.h -- batched potrf / trsm / potrs / posv (+ the gemm / syrk update steps), uniform size.
 *
 * Drop-in for the hot-path subset of the reference's include/kblas_batch.h:
 *   trsm : kblas_batch.h:773-897 (C++), 948-1053 (C)   potrf: 1380-1452, 1486-1566
 *   potrs: kblas_batch.h:2077-2150, 2190-2278          posv : 2772-2854, 2896-2990
 *
 * Conventions (all inherited from the reference):
 *  - column-major; element (i,j) of matrix b is  A[b*strideA + i + j*lda]  (strided)
 *    or  A_array[b][i + j*lda]  (pointer array; the ARRAY lives in device memory);
 *  - every matrix / pointer-array / info pointer is a DEVICE pointer, scalars by value;
 *  - work is enqueued on the handle's stream, asynchronously;
 *  - return value: KBLAS_Success (1) or a KBLAS_* error (<= 0), see kblas_defs.h;
 *  - the reference implements only uplo = 'L' (potrf/potrs/posv/trsm), diag = 'N' (trsm) and
 *    side = 'R' (potrs/posv) and returns KBLAS_NotImplemented for anything else
 *    (Xpotrf_batch_drivers.cuh:38-41, Xtrsm_batch_drivers.cuh:64-67,
 *     Xpotrs_batch_drivers.cuh:40-43, Xposv_batch_drivers.cuh:41-44).  Those are the fast paths
 *    here; uplo = 'U', diag = 'U' and side = 'L' are implemented as extensions (correct, tested, the
 *    Upper / Unit forms not tuned); the inverse family, syrk and the packed routines stay Lower only;
 *  - info_array is NOT written by default: the reference never stores a non-SPD
 *    code (Xpotrf_batch_kernels.cuh:121-129), a non-SPD input yields NaN/Inf in the
 *    factor.  Setting env KBLAS_B200_INFO_MODE=lapack before kblasCreate() opts in
 *    to LAPACK semantics (info[b] = j+1 of the first non-positive pivot, else 0).
 */
#ifndef KBLAS_B200_BATCH_H
#define KBLAS_B200_BATCH_H

#include "kblas_defs.h"

struct KBlasHandle;
typedef struct KBlasHandle *kblasHandle_t;

#ifdef __cplusplus
/* =====================================================================================
 * C++ API (mangled symbols, same overload set as the reference for float / double)
 * ===================================================================================== */

/* ---- workspace queries: accumulate (max) the bytes the corresponding call needs into
 *      handle->work_space.requested_ws_state; follow with kblasAllocateWorkspace().
 *      (reference src/workspace_queries.cu:257-266, 313-319, 340-346, 367-373) */
void kblas_trsm_batch_wsquery        (kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_trsm_batch_strided_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_potrf_batch_wsquery        (kblasHandle_t handle, const int n, int batchCount);
void kblas_potrf_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_potrs_batch_wsquery        (kblasHandle_t handle, const int m, const int n, int batchCount);
void kblas_potrs_batch_strided_wsquery(kblasHandle_t handle, const int m, const int n, int batchCount);
void kblas_posv_batch_wsquery        (kblasHandle_t handle, char side, const int m, const int n, int batchCount);
void kblas_posv_batch_strided_wsquery(kblasHandle_t handle, char side, const int m, const int n, int batchCount);

void kblas_gemm_batch_strided_wsquery(kblasHandle_t handle, int batchCount);      /* reference kblas_batch.h:35 */
void kblas_gemm_batch_nonuniform_wsquery(kblasHandle_t handle);
void kblas_syrk_batch_wsquery(kblasHandle_t handle, const int m, int batchCount); /* reference kblas_batch.h:485 */
void kblas_syrk_batch_nonuniform_wsquery(kblasHandle_t handle);
void kblas_trsm_batch_nonuniform_wsquery(kblasHandle_t handle);

/* the consumers of the factor (reference kblas_batch.h:1611-1622, 1840-1851, 2321-2332, 2547-2558) */
void kblas_trtri_batch_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_trtri_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_lauum_batch_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_lauum_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_potri_batch_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_potri_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_poti_batch_wsquery(kblasHandle_t handle, const int n, int batchCount);
void kblas_poti_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount);

#define KBLAS_B200_DECL_CPP(T)                                                              \
  /* op(A) X = alpha B (side L) or X op(A) = alpha B (side R); X overwrites B */            \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag,   \
                       const int m, const int n, const T alpha,                             \
                       const T **A, int lda, T **B, int ldb, int batchCount);               \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag,   \
                       const int m, const int n, const T alpha,                             \
                       const T *A, int lda, long strideA,                                   \
                       T *B, int ldb, long strideB, int batchCount);                        \
  /* A = L L^T in place, lower */                                                           \
  int kblas_potrf_batch(kblasHandle_t handle, char uplo, const int n,                       \
                        T **A, int lda, int batchCount, int *info_array);                   \
  int kblas_potrf_batch(kblasHandle_t handle, char uplo, const int n,                       \
                        T *A, int lda, long strideA, int batchCount, int *info_array);      \
  /* X (L L^T) = B, B is m x n, A is the n x n factor (side R) */                           \
  int kblas_potrs_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,\
                        const T **A, int lda, T **B, int ldb, int batchCount);              \
  int kblas_potrs_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,\
                        const T *A, int lda, long strideA,                                  \
                        T *B, int ldb, long strideB, int batchCount);                       \
  /* C = alpha op(A) op(B) + beta C  (reference kblas_batch.h:264-436: a cuBLAS wrapper there) */ \
  int kblas_gemm_batch(kblasHandle_t handle, char transA, char transB, const int m,         \
                       const int n, const int k, const T alpha, const T **A, int lda,       \
                       const T **B, int ldb, const T beta, T **C, int ldc, int batchCount); \
  int kblas_gemm_batch(kblasHandle_t handle, char transA, char transB, const int m,         \
                       const int n, const int k, const T alpha, const T *A, int lda,        \
                       long strideA, const T *B, int ldb, long strideB, const T beta,       \
                       T *C, int ldc, long strideC, int batchCount);                        \
  /* B(m x m, lower) = alpha op(A) op(A)^T + beta B, A m x n (trans N) or n x m (trans T)     \
     (reference kblas_batch.h:493-755) */                                                   \
  int kblas_syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m,            \
                       const int n, const T alpha, const T **A, int lda, const T beta,      \
                       T **B, int ldb, int batchCount);                                     \
  int kblas_syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m,            \
                       const int n, const T alpha, const T *A, int lda, long strideA,       \
                       const T beta, T *B, int ldb, long strideB, int batchCount);          \
  /* in place on the lower triangle: A := A^-1 (trtri), A := A^T A (lauum), A(=L) := (L L^T)^-1 (potri),   \
     A := A^-1 for SPD A (poti = potrf + potri); reference kblas_batch.h:1627-1797, 1856-2032, 2337-2506, 2563-2729 */ \
  int kblas_trtri_batch(kblasHandle_t handle, char uplo, char diag, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas_trtri_batch(kblasHandle_t handle, char uplo, char diag, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas_lauum_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas_lauum_batch(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas_potri_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas_potri_batch(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas_poti_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas_poti_batch(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  /* potrf(A) then potrs(A, B) */                                                           \
  int kblas_posv_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,\
                       T **A, int lda, T **B, int ldb, int batchCount, int *info_array);    \
  int kblas_posv_batch(kblasHandle_t handle, char side, char uplo, const int m, const int n,\
                       T *A, int lda, long strideA, T *B, int ldb, long strideB,            \
                       int batchCount, int *info_array);

KBLAS_B200_DECL_CPP(float)
KBLAS_B200_DECL_CPP(double)
#undef KBLAS_B200_DECL_CPP

extern "C" {
#endif /* __cplusplus */

/* =====================================================================================
 * C API (unmangled): kblas{S,D}<op>_batch[_strided]
 * ===================================================================================== */
#define KBLAS_B200_DECL_C(P, T)                                                             \
  int kblas##P##trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag,\
                           const int m, const int n, const T alpha,                         \
                           const T **A, int lda, T **B, int ldb, int batchCount);           \
  int kblas##P##trsm_batch_strided(kblasHandle_t handle, char side, char uplo, char trans,  \
                           char diag, const int m, const int n, const T alpha,              \
                           const T *A, int lda, long strideA,                               \
                           T *B, int ldb, long strideB, int batchCount);                    \
  int kblas##P##potrf_batch(kblasHandle_t handle, char uplo, const int n,                   \
                           T **A, int lda, int batchCount, int *info_array);                \
  int kblas##P##potrf_batch_strided(kblasHandle_t handle, char uplo, const int n,           \
                           T *A, int lda, long strideA, int batchCount, int *info_array);   \
  int kblas##P##potrs_batch(kblasHandle_t handle, char side, char uplo,                     \
                           const int m, const int n,                                        \
                           const T **A, int lda, T **B, int ldb, int batchCount);           \
  int kblas##P##potrs_batch_strided(kblasHandle_t handle, char side, char uplo,             \
                           const int m, const int n,                                        \
                           const T *A, int lda, long strideA,                               \
                           T *B, int ldb, long strideB, int batchCount);                    \
  int kblas##P##gemm_batch(kblasHandle_t handle, char transA, char transB, const int m,     \
                           const int n, const int k, const T alpha, const T **A, int lda,   \
                           const T **B, int ldb, const T beta, T **C, int ldc,              \
                           int batchCount);                                                 \
  int kblas##P##gemm_batch_strided(kblasHandle_t handle, char transA, char transB,          \
                           const int m, const int n, const int k, const T alpha,            \
                           const T *A, int lda, long strideA, const T *B, int ldb,          \
                           long strideB, const T beta, T *C, int ldc, long strideC,         \
                           int batchCount);                                                 \
  int kblas##P##syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m,        \
                           const int n, const T alpha, const T **A, int lda, const T beta,  \
                           T **B, int ldb, int batchCount);                                 \
  int kblas##P##syrk_batch_strided(kblasHandle_t handle, char uplo, char trans,             \
                           const int m, const int n, const T alpha, const T *A, int lda,    \
                           long strideA, const T beta, T *B, int ldb, long strideB,         \
                           int batchCount);                                                 \
  int kblas##P##trtri_batch(kblasHandle_t handle, char uplo, char diag, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas##P##trtri_batch_strided(kblasHandle_t handle, char uplo, char diag, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas##P##lauum_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas##P##lauum_batch_strided(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas##P##potri_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas##P##potri_batch_strided(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas##P##poti_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount, int *info_array); \
  int kblas##P##poti_batch_strided(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA, int batchCount, int *info_array); \
  int kblas##P##posv_batch(kblasHandle_t handle, char side, char uplo,                      \
                           const int m, const int n,                                        \
                           T **A, int lda, T **B, int ldb, int batchCount, int *info_array);\
  int kblas##P##posv_batch_strided(kblasHandle_t handle, char side, char uplo,              \
                           const int m, const int n,                                        \
                           T *A, int lda, long strideA, T *B, int ldb, long strideB,        \
                           int batchCount, int *info_array);

KBLAS_B200_DECL_C(S, float)
KBLAS_B200_DECL_C(D, double)
#undef KBLAS_B200_DECL_C

/** ceil(x / y) * y   (reference src/kblas_common.cu:257-260) */
int kblas_roundup(int x, int y);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* KBLAS_B200_BATCH_H */
