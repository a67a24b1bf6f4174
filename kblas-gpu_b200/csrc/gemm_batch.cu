// gemm_batch.cu -- kblas{S,D}gemm_batch[_strided] and kblas{S,D}syrk_batch[_strided] (uniform sizes).
//
// Counterparts of reference src/batch_triangular/Xgemm_batch.cu:126-169, 298-363 (a cuBLAS batched-GEMM wrapper,
// Xgemm_batch_core.cuh:170-313, 492-634) and Xsyrk_batch.cu:64-189 (register kernels on 8 / 16-wide diagonal blocks + one
// pointer-array GEMM per recursion level, Xsyrk_batch_drivers.cuh:32-433).  SURVEY.md §8(f)1: the trailing-update step of
// the Cholesky path as public entry points on this library's own fragment engine (kernels/gemm_tile.cuh) -- one launch,
// no cuBLAS, no workspace.  Contract kept: column-major, device pointers, asynchronous on handle->stream,
// KBLAS_Success == 1; SYRK Lower only (Upper -> KBLAS_NotImplemented with the reference's message,
// Xsyrk_batch_drivers.cuh:133-136); GEMM batchCount < 1 -> KBLAS_Error_WrongInput (Xgemm_batch_core.cuh:181-182);
// the workspace protocol (kblas_gemm_batch_strided_wsquery / kblas_syrk_batch_wsquery -> kblasAllocateWorkspace) is honoured:
// KBLAS_InsufficientWorkspace when the reference would answer that (Xsyrk_batch_drivers.cuh:141-149).
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/gemm_tile.cuh"

namespace kblasx {

template <typename T, bool STRIDED, bool SYRK, bool TA, bool TB>
static int launch_gemm_tile(KBlasHandle *h, const char *name, int m, int n, int k, T alpha, BatchRef<const T, STRIDED> A, int lda,
                            BatchRef<const T, STRIDED> B, int ldb, T beta, BatchRef<T, STRIDED> C, int ldc, int batchCount) {
  constexpr int WARPS = 4;
  const long tr = (m + 31) / 32, tc = (n + 31) / 32;
  const long tiles = SYRK ? tr * (tr + 1) / 2 : tr * tc;
  const long ntask = tiles * batchCount;
  if (ntask <= 0) return KBLAS_Success;
  long grid = (ntask + WARPS - 1) / WARPS;
  const long cap = (long)h->sm_count * 64;
  if (grid > cap) grid = cap;
  gemm_tile_kernel<T, STRIDED, SYRK, TA, TB, WARPS><<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, beta,
                                                                                                  C, ldc, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

static bool is_trans(char t) { return t == KBLAS_Trans || t == 't' || t == 'C' || t == 'c'; }

template <typename T, bool STRIDED>
static int gemm_batch_core(KBlasHandle *h, char transA, char transB, int m, int n, int k, T alpha, BatchRef<const T, STRIDED> A, int lda,
                           BatchRef<const T, STRIDED> B, int ldb, T beta, BatchRef<T, STRIDED> C, int ldc, int batchCount) {
  if (batchCount < 1) return KBLAS_Error_WrongInput;  // reference Xgemm_batch_core.cuh:181-182
  const bool ta = is_trans(transA), tb = is_trans(transB);
#define KX_G(TA_, TB_) launch_gemm_tile<T, STRIDED, false, TA_, TB_>(h, sizeof(T) == 8 ? "gemm_tile_dmma" : "gemm_tile_tf32x3", m, n, k, \
                                                                     alpha, A, lda, B, ldb, beta, C, ldc, batchCount)
  if (!ta && !tb) return KX_G(false, false);
  if (!ta && tb) return KX_G(false, true);
  if (ta && !tb) return KX_G(true, false);
  return KX_G(true, true);
#undef KX_G
}

template <typename T, bool STRIDED>
static int syrk_batch_core(KBlasHandle *h, char uplo, char trans, int m, int n, T alpha, BatchRef<const T, STRIDED> A, int lda, T beta,
                           BatchRef<T, STRIDED> B, int ldb, int batchCount) {
  if (uplo == KBLAS_Upper) {
    printf("Upper SYRK_BATCH is not implemented yet\n");  // reference Xsyrk_batch_drivers.cuh:133-136
    return KBLAS_NotImplemented;
  }
  if (m > 16) {  // where the reference needs (and checks) its pointer workspace, Xsyrk_batch_drivers.cuh:141-149
    KBlasWorkspaceState need;
    syrk_batch_wsquery_core(m, batchCount, &need);
    if (!need.isSufficient(&h->work_space.allocated_ws_state)) return KBLAS_InsufficientWorkspace;
  }
  if (batchCount <= 0) {
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);  // empty grid in the reference
  }
  BatchRef<const T, STRIDED> none = A;
  if (is_trans(trans))
    return launch_gemm_tile<T, STRIDED, true, true, false>(h, sizeof(T) == 8 ? "syrk_tile_dmma" : "syrk_tile_tf32x3", m, m, n, alpha, A, lda,
                                                           none, lda, beta, B, ldb, batchCount);
  return launch_gemm_tile<T, STRIDED, true, false, false>(h, sizeof(T) == 8 ? "syrk_tile_dmma" : "syrk_tile_tf32x3", m, m, n, alpha, A, lda,
                                                          none, lda, beta, B, ldb, batchCount);
}

static int gemm_strided_ws_check(KBlasHandle *h, int batchCount) {
  KBlasWorkspaceState need;
  gemm_batch_strided_wsquery_core(batchCount, &need);  // 0 bytes without MAGMA (reference workspace_queries.cu:203-208)
  return need.isSufficient(&h->work_space.allocated_ws_state) ? KBLAS_Success : KBLAS_InsufficientWorkspace;
}

}  // namespace kblasx

using kblasx::BatchRef;

// ---- public API: C++ overloads + C names (reference Xgemm_batch.cu:126-169, 298-363; Xsyrk_batch.cu:64-189) ----------
#define KX_GEMM_API(P, T)                                                                                             \
  int kblas_gemm_batch(kblasHandle_t handle, char transA, char transB, const int m, const int n, const int k,         \
                       const T alpha, const T **A, int lda, const T **B, int ldb, const T beta, T **C, int ldc,       \
                       int batchCount) {                                                                              \
    BatchRef<const T, false> a = {A, 0}, b = {B, 0};                                                                  \
    BatchRef<T, false> c = {C, 0};                                                                                    \
    return kblasx::gemm_batch_core<T, false>(handle, transA, transB, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc,    \
                                             batchCount);                                                             \
  }                                                                                                                   \
  int kblas_gemm_batch(kblasHandle_t handle, char transA, char transB, const int m, const int n, const int k,         \
                       const T alpha, const T *A, int lda, long strideA, const T *B, int ldb, long strideB,           \
                       const T beta, T *C, int ldc, long strideC, int batchCount) {                                   \
    if (batchCount >= 1 && kblasx::gemm_strided_ws_check(handle, batchCount) != KBLAS_Success)                        \
      return KBLAS_InsufficientWorkspace;                                                                             \
    BatchRef<const T, true> a = {A, strideA}, b = {B, strideB};                                                       \
    BatchRef<T, true> c = {C, strideC};                                                                               \
    return kblasx::gemm_batch_core<T, true>(handle, transA, transB, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc,     \
                                            batchCount);                                                              \
  }                                                                                                                   \
  extern "C" int kblas##P##gemm_batch(kblasHandle_t handle, char transA, char transB, const int m, const int n,       \
                                      const int k, const T alpha, const T **A, int lda, const T **B, int ldb,         \
                                      const T beta, T **C, int ldc, int batchCount) {                                 \
    return kblas_gemm_batch(handle, transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, batchCount);        \
  }                                                                                                                   \
  extern "C" int kblas##P##gemm_batch_strided(kblasHandle_t handle, char transA, char transB, const int m,            \
                                              const int n, const int k, const T alpha, const T *A, int lda,           \
                                              long strideA, const T *B, int ldb, long strideB, const T beta, T *C,    \
                                              int ldc, long strideC, int batchCount) {                                \
    return kblas_gemm_batch(handle, transA, transB, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,   \
                            strideC, batchCount);                                                                     \
  }                                                                                                                   \
  int kblas_syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m, const int n, const T alpha,          \
                       const T **A, int lda, const T beta, T **B, int ldb, int batchCount) {                          \
    BatchRef<const T, false> a = {A, 0};                                                                              \
    BatchRef<T, false> b = {B, 0};                                                                                    \
    return kblasx::syrk_batch_core<T, false>(handle, uplo, trans, m, n, alpha, a, lda, beta, b, ldb, batchCount);     \
  }                                                                                                                   \
  int kblas_syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m, const int n, const T alpha,          \
                       const T *A, int lda, long strideA, const T beta, T *B, int ldb, long strideB,                  \
                       int batchCount) {                                                                              \
    BatchRef<const T, true> a = {A, strideA};                                                                         \
    BatchRef<T, true> b = {B, strideB};                                                                               \
    return kblasx::syrk_batch_core<T, true>(handle, uplo, trans, m, n, alpha, a, lda, beta, b, ldb, batchCount);      \
  }                                                                                                                   \
  extern "C" int kblas##P##syrk_batch(kblasHandle_t handle, char uplo, char trans, const int m, const int n,          \
                                      const T alpha, const T **A, int lda, const T beta, T **B, int ldb,              \
                                      int batchCount) {                                                               \
    return kblas_syrk_batch(handle, uplo, trans, m, n, alpha, A, lda, beta, B, ldb, batchCount);                      \
  }                                                                                                                   \
  extern "C" int kblas##P##syrk_batch_strided(kblasHandle_t handle, char uplo, char trans, const int m, const int n,  \
                                              const T alpha, const T *A, int lda, long strideA, const T beta, T *B,   \
                                              int ldb, long strideB, int batchCount) {                                \
    return kblas_syrk_batch(handle, uplo, trans, m, n, alpha, A, lda, strideA, beta, B, ldb, strideB, batchCount);    \
  }
KX_GEMM_API(S, float)
KX_GEMM_API(D, double)

// workspace queries (reference src/workspace_queries.cu:196-224, 227-254)
void kblas_gemm_batch_strided_wsquery(kblasHandle_t handle, int batchCount) {
  kblasx::gemm_batch_strided_wsquery_core(batchCount, &handle->work_space.requested_ws_state);
}
void kblas_gemm_batch_offset_wsquery(kblasHandle_t handle, int batchCount, bool offseted) {
  kblasx::gemm_batch_offset_wsquery_core(batchCount, offseted, &handle->work_space.requested_ws_state);
}
void kblas_gemm_batch_nonuniform_wsquery(kblasHandle_t) {}  // MAGMA-only in the reference: records nothing without it
void kblas_syrk_batch_wsquery(kblasHandle_t handle, const int m, int batchCount) {
  kblasx::syrk_batch_wsquery_core(m, batchCount, &handle->work_space.requested_ws_state);
}
void kblas_syrk_batch_nonuniform_wsquery(kblasHandle_t) {}
