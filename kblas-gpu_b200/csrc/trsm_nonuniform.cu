// trsm_nonuniform.cu -- kblas_trsm_batch with per-matrix sizes (device arrays m[b], n[b], lda[b], ldb[b]).
//
// The reference implements the non-uniform batch only through MAGMA (Xtrsm_batch_nonuniform_core,
// Xtrsm_batch_drivers.cuh:277-367: it needs max(m), max(n) -- reduced on the device into workspace when the caller does not
// pass them -- and a MAGMA build; otherwise KBLAS_WrongConfig).  Here it is native: one warp per matrix
// walks ITS OWN 32-vector slabs and 32 x 32 factor blocks with the runtime-sized blocked substitution of
// kernels/trsm_blocked.cuh, so neither the maxima nor workspace are needed; all side / uplo / trans / diag variants.
// Matrices with m[b] <= 0 or n[b] <= 0 are skipped.
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/trsm_blocked.cuh"

namespace kblasx {

template <typename T, bool LEFT, bool FORWARD>
__global__ void __launch_bounds__(128)
tri_solve_nonuniform_kernel(const int *__restrict__ m, const int *__restrict__ n, const T alpha, const T *const *__restrict__ A_array,
                            const int A_row_off, const int A_col_off, const int *__restrict__ lda, T *const *__restrict__ B_array,
                            const int B_row_off, const int B_col_off, const int *__restrict__ ldb, const int batchCount,
                            const int flags) {
  constexpr int WARPS = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *smem_w = reinterpret_cast<T *>(smem_raw) + warp * TriBlockedSmem<T, 32>::per_warp;
  const long mat = (long)blockIdx.x * WARPS + warp;
  if (mat >= (long)batchCount) return;  // warp-uniform
  const int mm = m[mat], nn = n[mat];
  const int k = LEFT ? mm : nn, vec = LEFT ? nn : mm;
  if (k <= 0 || vec <= 0) return;
  const int la = lda[mat], lb = ldb[mat];
  const T *Aq[1] = {A_array[mat] + A_row_off + (long)A_col_off * la};
  T *B = B_array[mat] + B_row_off + (long)B_col_off * lb;
  for (int v0 = 0; v0 < vec; v0 += 32) {
    const int my = v0 + lane;
    tri_blocked_pass<T, LEFT, FORWARD, 32>(k, alpha, Aq, la, B, lb, my, my < vec, smem_w, smem_w, lane, flags);
    __syncwarp();
  }
}

template <typename T>
static int trsm_nonuniform(KBlasHandle *h, char side, char uplo, char trans, char diag, const int *m, const int *n, T alpha,
                           const T *const *A, int A_row_off, int A_col_off, const int *lda, T *const *B, int B_row_off,
                           int B_col_off, const int *ldb, int batchCount) {
  const bool left = (side == KBLAS_Left);
  if (!left && side != KBLAS_Right) return KBLAS_NotImplemented;
  if (batchCount <= 0) return KBLAS_Success;
  const bool upper = (uplo == KBLAS_Upper), unit = (diag == KBLAS_Unit);
  const bool notrans = (trans == KBLAS_NoTrans) != upper;  // an upper factor is staged as L = U^T
  const bool forward = (left == notrans);                  // forward: (R, T) and (L, N)
  const int flags = (upper ? TRI_FLAG_UPPER : 0) | (unit ? TRI_FLAG_UNIT : 0);
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * TriBlockedSmem<T, 32>::per_warp * sizeof(T);
  const unsigned grid = (unsigned)((batchCount + WARPS - 1) / WARPS);
#define KX_GO(L_, F_)                                                                                                      \
  do {                                                                                                                     \
    auto kern = tri_solve_nonuniform_kernel<T, L_, F_>;                                                                    \
    check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);                                                       \
    kern<<<grid, WARPS * 32, smem, h->stream>>>(m, n, alpha, A, A_row_off, A_col_off, lda, B, B_row_off, B_col_off, ldb,   \
                                                batchCount, flags);                                                        \
  } while (0)
  if (left && forward) KX_GO(true, true);
  else if (left) KX_GO(true, false);
  else if (forward) KX_GO(false, true);
  else KX_GO(false, false);
#undef KX_GO
  h->note_launch("tri_nonuniform");
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

}  // namespace kblasx

// ---- the reference's three non-uniform entry points (Xtrsm_batch.cu, kblas_batch.h) + C twins for foreign-function callers
#define KX_TRSM_NONUNIFORM_API(P, T)                                                                                          \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n, int /*max_m*/,           \
                  int /*max_n*/, T alpha, T **A, int A_row_off, int A_col_off, int *lda, long /*strideA*/, T **B,             \
                  int B_row_off, int B_col_off, int *ldb, long /*strideB*/, int batchCount) {                                 \
    return kblasx::trsm_nonuniform<T>(handle, side, uplo, trans, diag, m, n, alpha, A, A_row_off, A_col_off, lda, B,          \
                                      B_row_off, B_col_off, ldb, batchCount);                                                 \
  }                                                                                                                           \
  int Xtrsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n, T alpha, T **A,          \
                  int A_row_off, int A_col_off, int *lda, long /*strideA*/, T **B, int B_row_off, int B_col_off, int *ldb,    \
                  long /*strideB*/, int batchCount) {                                                                         \
    return kblasx::trsm_nonuniform<T>(handle, side, uplo, trans, diag, m, n, alpha, A, A_row_off, A_col_off, lda, B,          \
                                      B_row_off, B_col_off, ldb, batchCount);                                                 \
  }                                                                                                                           \
  int kblas_trsm_batch(kblasHandle_t handle, char side, char uplo, char trans, char diag, int *m, int *n, int /*max_m*/,      \
                       int /*max_n*/, T alpha, T **A, int *lda, T **B, int *ldb, int batchCount) {                            \
    return kblasx::trsm_nonuniform<T>(handle, side, uplo, trans, diag, m, n, alpha, A, 0, 0, lda, B, 0, 0, ldb, batchCount);  \
  }                                                                                                                           \
  extern "C" int kblasx##P##trsm_batch_nonuniform(kblasHandle_t handle, char side, char uplo, char trans, char diag,          \
                                                  const int *m, const int *n, T alpha, const T *const *A, const int *lda,     \
                                                  T *const *B, const int *ldb, int batchCount) {                              \
    return kblasx::trsm_nonuniform<T>(handle, side, uplo, trans, diag, m, n, alpha, A, 0, 0, lda, B, 0, 0, ldb, batchCount);  \
  }
KX_TRSM_NONUNIFORM_API(S, float)
KX_TRSM_NONUNIFORM_API(D, double)
