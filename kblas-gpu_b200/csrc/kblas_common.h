// kblas_common.h -- internal helpers shared by every translation unit.
// Counterpart of reference src/kblas_common.h + src/kblas_error.h.
#pragma once

#include <cstdio>
#include <cuda_runtime.h>
#include "kblas_struct.h"

// reference src/kblas_common.cu:241-255 (C++ linkage there as well)
bool REG_SIZE(int n);          // n is a power of two (> 0)
int CLOSEST_REG_SIZE(int n);   // largest power of two strictly below n (n > 0), else 0
long kblas_roundup_l(long x, long y);
size_t kblas_roundup_s(size_t x, size_t y);

// reference src/kblas_common.h:35-36, src/kblas_common.cu:344-386
int iset_value_1(int *output_array, int input, long batchCount, cudaStream_t cuda_stream);
int iset_value_2(int *output_array1, int input1, int *output_array2, int input2, long batchCount,
                 cudaStream_t cuda_stream);
int iset_value_4(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, long batchCount, cudaStream_t cuda_stream);
int iset_value_5(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, int *output_array5, int input5, long batchCount,
                 cudaStream_t cuda_stream);

// reference src/kblas_common.cu:208-238: print to stderr, return 1 on success / 0 on error
int _kblas_error(cudaError_t err, const char *func, const char *file, int line);
int _kblas_error(int err, const char *func, const char *file, int line);

// reference src/kblas_error.h:31-40
#define check_error(err_) _kblas_error((err_), __func__, __FILE__, __LINE__)
#define check_error_ret(err_, ret_)                                       \
  do {                                                                    \
    if (!_kblas_error((err_), __func__, __FILE__, __LINE__)) return (ret_); \
  } while (0)
#define check_ret_error(call_)                                              \
  do {                                                                      \
    int _st = (call_);                                                      \
    if (!_kblas_error(_st, __func__, __FILE__, __LINE__)) return _st;        \
  } while (0)

// ---- launch helpers: per-handle (per-device) caches instead of function-local statics -------------
// resident CTAs per SM of `kern` at this block size / dynamic shared memory on the handle's device
template <typename K>
inline int kx_ctas_per_sm(KBlasHandle *h, K kern, int threads, size_t smem, int fallback) {
  KBlasHandle::KernelNote *k = h->kernel_note((const void *)kern);
  if (k->ctas_per_sm == 0) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    k->ctas_per_sm = occ > 0 ? occ : fallback;
  }
  return k->ctas_per_sm;
}
// raise the dynamic shared-memory limit of `kern` on the CURRENT device, once per handle, to the device's opt-in
// maximum: the attribute is per function and PROCESS-wide, so a per-size value set through one handle could be lowered
// again through another one (two handles, different n) -- the maximum is safe for every launch and costs nothing
// (occupancy follows the bytes actually requested at launch)
template <typename K>
inline cudaError_t kx_allow_smem(KBlasHandle *h, K kern, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  KBlasHandle::KernelNote *k = h->kernel_note((const void *)kern);
  if (k->smem_limit > 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin_max);
  if (e == cudaSuccess) k->smem_limit = h->smem_optin_max;
  return e;
}

// ---- implementation functions behind both the C++-mangled API (kblas_common.cu,
//      workspace_queries.cu) and its C-linkage twins (ffi.cu) ---------------------------
namespace kblasx {

enum WsOp { WS_TRSM = 0, WS_POTRF = 1, WS_POTRS = 2, WS_POSV = 3 };

// pure host arithmetic, reference src/workspace_queries.ch:60-74,111-122,157-171,194-201
void trsm_batch_wsquery_core(bool strided, int batchCount, char side, int m, int n, KBlasWorkspaceState *ws);
void potrf_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws);
void potrs_batch_wsquery_core(bool strided, int m, int n, int batchCount, KBlasWorkspaceState *ws);
void posv_batch_wsquery_core(bool strided, int m, int n, char side, int batchCount, KBlasWorkspaceState *ws);
void gemm_batch_offset_wsquery_core(int batchCount, bool offseted, KBlasWorkspaceState *ws);
void gemm_batch_strided_wsquery_core(int batchCount, KBlasWorkspaceState *ws);
void syrk_batch_wsquery_core(int m, int batchCount, KBlasWorkspaceState *ws);
void trmm_batch_wsquery_core(bool strided, int batchCount, char side, int m, int n, KBlasWorkspaceState *ws);
void lauum_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws);
void trtri_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws);
void potri_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws);
void poti_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws);

int create(KBlasHandle **handle);
int destroy(KBlasHandle **handle);
void set_stream(KBlasHandle *handle, cudaStream_t stream);
cublasHandle_t get_cublas(KBlasHandle *handle);
const char *error_string(int error);

// pointer-array builders, reference src/batch_triangular/Xhelper_funcs.cu:74-105
template <typename T>
int set_pointer_1(T **out, const T *in, int lda, long batch_offset, long batchCount, cudaStream_t s);
template <typename T>
int set_pointer_2(T **out1, const T *in1, int ld1, long off1, T **out2, const T *in2, int ld2, long off2,
                  long batchCount, cudaStream_t s);
template <typename T>
int set_pointer_3(T **out1, const T *in1, int ld1, long off1, T **out2, const T *in2, int ld2, long off2,
                  T **out3, const T *in3, int ld3, long off3, long batchCount, cudaStream_t s);

}  // namespace kblasx
