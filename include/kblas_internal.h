/*
 * kblas_internal.h -- C++-linkage names that are NOT in the reference's public headers but that its own
 * test programs and sibling routines link against (they put src/ on their include path, testing/Makefile:12).
 * libkblas-gpu.so exports every one of them with the reference's mangled name, so objects compiled against
 * the reference's src/*.ch headers link unchanged.  Device pointers throughout; work goes to the given stream
 * (helpers) or to handle->stream (routines).
 *
 *   Xset_pointer_{1,2,3}      src/Xhelper_funcs.ch:48-60, src/batch_triangular/Xhelper_funcs.cu:74-105
 *   iset_value_{1,2,4,5}      src/kblas_common.h:35-36,  src/kblas_common.cu:344-386
 *   X{potrf,potrs,posv}_batch_offset, Xtrsm_batch (uniform)
 *                             src/Xblas_core.ch:194-277, src/batch_triangular/Xpotrf_batch.cu:44-63,107-127,
 *                             Xpotrs_batch.cu:42-58,104-127, Xposv_batch.cu:42-67,112-138, Xtrsm_batch.cu:42-58,189-216
 *   REG_SIZE / CLOSEST_REG_SIZE / kblas_roundup_{l,s}
 *                             src/kblas_common.cu:241-268
 * The (row_off, col_off) arguments select the sub-matrix that starts at element (row_off, col_off) of every
 * batch entry; in pointer-array mode the offset is applied inside the kernels (the reference launches pointer
 * fix-up kernels into its d_ptrs workspace instead).  Workspace protocol and return codes as for the public calls.
 */
#ifndef KBLAS_B200_INTERNAL_H
#define KBLAS_B200_INTERNAL_H

#ifndef __cplusplus
#error "C++ only: plain-C users bind include/kblas_ffi.h"
#endif

#include <cstddef>
#include <cuda_runtime_api.h>
#include "kblas.h"

bool REG_SIZE(int n);
int CLOSEST_REG_SIZE(int n);
long kblas_roundup_l(long x, long y);
size_t kblas_roundup_s(size_t x, size_t y);

int iset_value_1(int *output_array, int input, long batchCount, cudaStream_t cuda_stream);
int iset_value_2(int *output_array1, int input1, int *output_array2, int input2, long batchCount,
                 cudaStream_t cuda_stream);
int iset_value_4(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, long batchCount, cudaStream_t cuda_stream);
int iset_value_5(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, int *output_array5, int input5, long batchCount,
                 cudaStream_t cuda_stream);

#define KBLAS_B200_DECL_INTERNAL(T)                                                                            \
  /* output_array[i] = input + i * batch_offset */                                                             \
  int Xset_pointer_1(T **output_array, const T *input, int lda, long batch_offset, long batchCount,            \
                     cudaStream_t cuda_stream);                                                                \
  int Xset_pointer_2(T **output_array1, const T *input1, int ldinput1, long batch_offset1,                     \
                     T **output_array2, const T *input2, int ldinput2, long batch_offset2,                     \
                     long batchCount, cudaStream_t cuda_stream);                                               \
  int Xset_pointer_3(T **output_array1, const T *input1, int ldinput1, long batch_offset1,                     \
                     T **output_array2, const T *input2, int ldinput2, long batch_offset2,                     \
                     T **output_array3, const T *input3, int ldinput3, long batch_offset3,                     \
                     long batchCount, cudaStream_t cuda_stream);                                               \
  /* pointer-array inputs: output_array[i] = input[i] + offset_r + offset_c * lda  (lda[i] in the first form) */ \
  int Xset_pointer_1(T **output_array, T **input, int offset_r, int offset_c, int *lda, long batchCount,       \
                     cudaStream_t cuda_stream);                                                                \
  int Xset_pointer_2(T **output_array1, const T **input1, int offset_r1, int offset_c1, int lda1,              \
                     T **output_array2, const T **input2, int offset_r2, int offset_c2, int lda2,              \
                     long batchCount, cudaStream_t cuda_stream);                                               \
  int Xset_pointer_3(T **output_array1, const T **input1, int offset_r1, int offset_c1, int lda1,              \
                     T **output_array2, const T **input2, int offset_r2, int offset_c2, int lda2,              \
                     T **output_array3, const T **input3, int offset_r3, int offset_c3, int lda3,              \
                     long batchCount, cudaStream_t cuda_stream);                                               \
  int Xpotrf_batch_offset(kblasHandle_t handle, char uplo, const int n,                                        \
                          T **A, int A_row_off, int A_col_off, int lda, int batchCount, int *info_array);      \
  int Xpotrf_batch_offset(kblasHandle_t handle, char uplo, const int n,                                        \
                          T *A, int A_row_off, int A_col_off, int lda, long strideA,                           \
                          int batchCount, int *info_array);                                                    \
  int Xpotrs_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n,                \
                          const T **A, int A_row_off, int A_col_off, int lda,                                  \
                          T **B, int B_row_off, int B_col_off, int ldb, int batchCount);                       \
  int Xpotrs_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n,                \
                          const T *A, int A_row_off, int A_col_off, int lda, long strideA,                     \
                          T *B, int B_row_off, int B_col_off, int ldb, long strideB, int batchCount);          \
  int Xposv_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n,                 \
                         T **A, int A_row_off, int A_col_off, int lda,                                         \
                         T **B, int B_row_off, int B_col_off, int ldb, int batchCount, int *info_array);       \
  int Xposv_batch_offset(kblasHandle_t handle, char side, char uplo, const int m, const int n,                 \
                         T *A, int A_row_off, int A_col_off, int lda, long strideA,                            \
                         T *B, int B_row_off, int B_col_off, int ldb, long strideB,                            \
                         int batchCount, int *info_array);                                                     \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int m, int n, T alpha,    \
                  T **A, int A_row_off, int A_col_off, int lda, long strideA,                                  \
                  T **B, int B_row_off, int B_col_off, int ldb, long strideB, int batchCount);                 \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int m, int n, T alpha,    \
                  T *A, int A_row_off, int A_col_off, int lda, long strideA,                                   \
                  T *B, int B_row_off, int B_col_off, int ldb, long strideB, int batchCount);                  \
  /* non-uniform batch: m, n, lda, ldb are DEVICE arrays of batchCount entries (the reference: MAGMA builds only,      \
     Xtrsm_batch_drivers.cuh:277-367; native here, max_m / max_n / strides accepted and ignored) */                      \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n, int max_m,   \
                  int max_n, T alpha, T **A, int A_row_off, int A_col_off, int *lda, long strideA,                \
                  T **B, int B_row_off, int B_col_off, int *ldb, long strideB, int batchCount);                   \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n, T alpha,     \
                  T **A, int A_row_off, int A_col_off, int *lda, long strideA,                                    \
                  T **B, int B_row_off, int B_col_off, int *ldb, long strideB, int batchCount);                   \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n,         \
                       int max_m, int max_n, T alpha, T **A, int *lda, T **B, int *ldb, int batchCount);

KBLAS_B200_DECL_INTERNAL(float)
KBLAS_B200_DECL_INTERNAL(double)
#undef KBLAS_B200_DECL_INTERNAL

#endif /* KBLAS_B200_INTERNAL_H */
