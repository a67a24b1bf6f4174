/*
 * kblas_ffi.h -- the complete C ABI of libkblas-gpu.so (plain C, no CUDA headers needed).
 *
 * This is the boundary a foreign-function binding (ctypes, cgo, JNI, N-API ...) binds.
 * It consists of
 *   (1) the kblas{S,D}<op>_batch[_strided] entry points, which already have C linkage in
 *       the reference (include/kblas_batch.h:948-1053, 1486-1566, 2190-2278, 2896-2990);
 *   (2) C-linkage twins of the management / workspace calls that the reference only
 *       offers with C++ linkage (include/kblas.h:54-108; kblas_batch.h:773,786,1380,1391,
 *       2077,2089,2772,2785; src/Xhelper_funcs.ch:48-55; src/kblas_common.h:35-36).
 *       Same names, same arguments; cudaStream_t / cublasHandle_t appear as void*.
 *       The mangled C++ symbols are exported as well (include/kblas.h), so code
 *       compiled against the reference's headers links unchanged;
 *   (3) a few kblasx_* introspection calls (no reference counterpart) used by the parity
 *       tests and the bench.
 *
 * Do not include this header together with kblas.h in one C++ translation unit: the two
 * declare the same names with different linkage on purpose.
 */
#ifndef KBLAS_B200_FFI_H
#define KBLAS_B200_FFI_H

#ifdef __cplusplus
#error "kblas_ffi.h is the plain-C FFI view; C++ code includes kblas.h"
#endif

#include <stddef.h>
#include "kblas_batch.h"   /* (1): kblas{S,D}{trsm,potrf,potrs,posv}_batch[_strided], kblas_roundup */

/* (2) management -- reference include/kblas.h:54-108, src/kblas_common.cu:35-202 */
int         kblasCreate(kblasHandle_t *handle);
int         kblasDestroy(kblasHandle_t *handle);
void        kblasTimerTic(kblasHandle_t handle);
void        kblasTimerRecordEnd(kblasHandle_t handle);
double      kblasTimerToc(kblasHandle_t handle);
int         kblasCreateStreams(kblasHandle_t handle, int nStreams);
void       *kblasGetStream(kblasHandle_t handle);                 /* cudaStream_t */
void        kblasSetStream(kblasHandle_t handle, void *stream);   /* cudaStream_t */
void       *kblasGetCublasHandle(kblasHandle_t handle);           /* cublasHandle_t */
int         kblasEnableMagma(kblasHandle_t handle);
const char *kblasGetErrorString(int error);
int         kblasAllocateWorkspace(kblasHandle_t handle);
int         kblasFreeWorkspace(kblasHandle_t handle);

/* (2) workspace queries -- reference src/workspace_queries.cu:257-266,313-319,340-346,367-373 */
void kblas_trsm_batch_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_trsm_batch_strided_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_potrf_batch_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_potrf_batch_strided_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_potrs_batch_wsquery(kblasHandle_t handle, int m, int n, int batchCount);
void kblas_potrs_batch_strided_wsquery(kblasHandle_t handle, int m, int n, int batchCount);
void kblas_posv_batch_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_posv_batch_strided_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount);
void kblas_gemm_batch_strided_wsquery(kblasHandle_t handle, int batchCount);       /* workspace_queries.cu:216-219 */
void kblas_syrk_batch_wsquery(kblasHandle_t handle, int m, int batchCount);        /* workspace_queries.cu:239-242 */
void kblas_trtri_batch_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_trtri_batch_strided_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_lauum_batch_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_lauum_batch_strided_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_potri_batch_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_potri_batch_strided_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_poti_batch_wsquery(kblasHandle_t handle, int n, int batchCount);
void kblas_poti_batch_strided_wsquery(kblasHandle_t handle, int n, int batchCount);

/* (2) pointer-array / value helpers the reference's own test binaries call
 *     (src/Xhelper_funcs.ch:48-55 -> S/D suffix replaces the C++ overload;
 *      src/kblas_common.h:35-36).  output[i] = input + i*batch_offset. */
int kblasSset_pointer_1(float **output_array, const float *input, int lda, long batch_offset,
                        long batchCount, void *stream);
int kblasDset_pointer_1(double **output_array, const double *input, int lda, long batch_offset,
                        long batchCount, void *stream);
int kblasSset_pointer_2(float **output_array1, const float *input1, int ld1, long batch_offset1,
                        float **output_array2, const float *input2, int ld2, long batch_offset2,
                        long batchCount, void *stream);
int kblasDset_pointer_2(double **output_array1, const double *input1, int ld1, long batch_offset1,
                        double **output_array2, const double *input2, int ld2, long batch_offset2,
                        long batchCount, void *stream);
int kblasSset_pointer_3(float **output_array1, const float *input1, int ld1, long batch_offset1,
                        float **output_array2, const float *input2, int ld2, long batch_offset2,
                        float **output_array3, const float *input3, int ld3, long batch_offset3,
                        long batchCount, void *stream);
int kblasDset_pointer_3(double **output_array1, const double *input1, int ld1, long batch_offset1,
                        double **output_array2, const double *input2, int ld2, long batch_offset2,
                        double **output_array3, const double *input3, int ld3, long batch_offset3,
                        long batchCount, void *stream);
int kblas_iset_value_1(int *output_array, int input, long batchCount, void *stream);
/* reference src/kblas_common.cu:351-386: several int arrays filled in one call */
int kblas_iset_value_2(int *output_array1, int input1, int *output_array2, int input2, long batchCount, void *stream);
int kblas_iset_value_4(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                       int *output_array4, int input4, long batchCount, void *stream);
int kblas_iset_value_5(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                       int *output_array4, int input4, int *output_array5, int input5, long batchCount, void *stream);

/* (3) introspection: no reference counterpart */
/** workspace bytes recorded in the handle: which = 0 requested, 1 allocated, 2 consumed;
 *  out[4] = {h_data, h_ptrs, d_data, d_ptrs} (KBlasWorkspaceState, src/kblas_struct.h:43-91). */
int         kblasx_workspace_state(kblasHandle_t handle, int which, size_t out[4]);
/** pure host arithmetic of the *_wsquery cores, usable without a GPU or a handle.
 *  op: 0 trsm, 1 potrf, 2 potrs, 3 posv.  out[4] as above. */
int         kblasx_wsquery_bytes(int op, int strided, char side, int m, int n, int batchCount,
                                 size_t out[4]);
/** shared-memory slot plan of the fp64 Cholesky for 32 < n <= 256 (csrc/kernels/potrf_smem.cuh): out[I*8+K] = 8 KiB slot
 *  of the 32 x 32 block (I, K), K <= I < nblk; returns the number of slots.  Host logic, exposed for the tests. */
int         kblasx_potrf_smem_plan(int nblk, unsigned char out[64]);
/** number of kernels this handle has launched since creation (the bench's gpu_launches). */
long        kblasx_launch_count(kblasHandle_t handle);
/** name of the kernel variant the most recent call on this handle dispatched to. */
const char *kblasx_last_kernel(kblasHandle_t handle);
/** library / build identification string ("kblas-b200 <ver> sm_100a ..."). */
const char *kblasx_version(void);
/** REG_SIZE / CLOSEST_REG_SIZE of the reference (src/kblas_common.cu:241-255). */
int         kblasx_reg_size(int n);
int         kblasx_closest_reg_size(int n);


/* (4) host-memory entry points: no reference counterpart (the reference takes device pointers only; its
 *     callers cudaMemcpy whole arrays around the call, testing/batch_triangular/test_Xpotrf_batch.cpp:170-206).
 *  Strided batch Cholesky of matrices that live in HOST memory (pinned for full speed): 256 MiB chunks
 *  through three device staging buffers on three streams, so H2D, the kernel and D2H overlap.
 *  Synchronous: the result is in A_out on return.  A_out == A_in is the in-place LAPACK form and is
 *  bit-identical to cudaMemcpy + kblas?potrf_batch_strided + cudaMemcpy.  info_host is written only in
 *  KBLAS_B200_INFO_MODE=lapack.  KBLAS_B200_HOSTCOPY=tri moves only the lower triangle (strided 3-D
 *  copies; measured slower on PCIe Gen5, see csrc/host_pipeline.cu); out of place it then leaves the
 *  elements of A_out above the diagonal 8 x 8 blocks unwritten. */
int kblasxSpotrf_batch_strided_host(kblasHandle_t handle, char uplo, int n, const float *A_in, float *A_out,
                                    int lda, long strideA, int batchCount, int *info_host);
int kblasxDpotrf_batch_strided_host(kblasHandle_t handle, char uplo, int n, const double *A_in, double *A_out,
                                    int lda, long strideA, int batchCount, int *info_host);


/* (5) packed lower-triangular batch layout: no reference counterpart (the reference's nearest routine
 *     is the pivoted batch_pstrf on full storage, include/batch_pstrf.h).  Matrix b is stored as LAPACK ?pptrf takes it
 *     for uplo = 'L':  AP_b[ j*n - j(j-1)/2 + (i-j) ] = A_b(i,j), i >= j;  n(n+1)/2 elements, matrix b at AP + b*strideAP
 *     (strideAP >= n(n+1)/2) or AP_array[b].  Physical bytes == algorithmic bytes: a 32 x 32 fp64 matrix moves 8448 B
 *     instead of the 12288 B of DRAM lines the column-major layout costs.  n <= 32; same contract as
 *     kblas?potrf_batch otherwise (Lower only, info untouched unless KBLAS_B200_INFO_MODE=lapack, async on the handle's
 *     stream, empty batch -> KBLAS_UnknownError).  For n % 8 == 0 the factor is bit-identical to kblas?potrf_batch's.
 *     The TMA-staged fast path needs 16-byte aligned matrices (AP and strideAP*sizeof(T) multiples of 16). */
int kblasxSpptrf_batch_strided(kblasHandle_t handle, char uplo, int n, float *AP, long strideAP, int batchCount,
                               int *info_array);
int kblasxDpptrf_batch_strided(kblasHandle_t handle, char uplo, int n, double *AP, long strideAP, int batchCount,
                               int *info_array);
int kblasxSpptrf_batch(kblasHandle_t handle, char uplo, int n, float **AP_array, int batchCount, int *info_array);
int kblasxDpptrf_batch(kblasHandle_t handle, char uplo, int n, double **AP_array, int batchCount, int *info_array);
/* lower triangle of column-major A (lda, strideA) -> packed AP, and back (unpack writes ONLY the lower triangle of A) */
int kblasxStri_pack_batch_strided(kblasHandle_t handle, char uplo, int n, const float *A, int lda, long strideA,
                                  float *AP, long strideAP, int batchCount);
int kblasxDtri_pack_batch_strided(kblasHandle_t handle, char uplo, int n, const double *A, int lda, long strideA,
                                  double *AP, long strideAP, int batchCount);
int kblasxStri_unpack_batch_strided(kblasHandle_t handle, char uplo, int n, const float *AP, long strideAP,
                                    float *A, int lda, long strideA, int batchCount);
int kblasxDtri_unpack_batch_strided(kblasHandle_t handle, char uplo, int n, const double *AP, long strideAP,
                                    double *A, int lda, long strideA, int batchCount);
/* packed matrices in HOST memory: the chunked H2D / pptrf / D2H pipeline of (4) on packed storage -- n(n+1)/2 elements per
 * matrix cross PCIe each way instead of n*n */
int kblasxSpptrf_batch_strided_host(kblasHandle_t handle, char uplo, int n, const float *AP_in, float *AP_out,
                                    long strideAP, int batchCount, int *info_host);
int kblasxDpptrf_batch_strided_host(kblasHandle_t handle, char uplo, int n, const double *AP_in, double *AP_out,
                                    long strideAP, int batchCount, int *info_host);

/* (6) non-uniform TRSM: per-matrix sizes in DEVICE arrays m[b], n[b], lda[b], ldb[b]; A_array / B_array device pointer
 *     arrays.  C twin of the reference's C++-only kblas_trsm_batch(handle, side, uplo, trans, diag, int *m, int *n, max_m,
 *     max_n, alpha, T **A, int *lda, T **B, int *ldb, batchCount) (MAGMA builds only there, native here; all side / uplo /
 *     trans / diag variants; matrices with a non-positive dimension are skipped) */
int kblasxStrsm_batch_nonuniform(kblasHandle_t handle, char side, char uplo, char trans, char diag, const int *m,
                                 const int *n, float alpha, const float *const *A_array, const int *lda,
                                 float *const *B_array, const int *ldb, int batchCount);
int kblasxDtrsm_batch_nonuniform(kblasHandle_t handle, char side, char uplo, char trans, char diag, const int *m,
                                 const int *n, double alpha, const double *const *A_array, const int *lda,
                                 double *const *B_array, const int *ldb, int batchCount);

#endif /* KBLAS_B200_FFI_H */
