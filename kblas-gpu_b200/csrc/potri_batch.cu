// potri_batch.cu -- kblas{S,D}{trtri,lauum,potri,poti}_batch[_strided]: the consumers of the Cholesky factor.
//
// Counterparts of reference src/batch_triangular/X{trtri,lauum,potri,poti}_batch.cu and their drivers
// (Xtrtri_batch_drivers.cuh:31-125, Xlauum_batch_drivers.cuh:31-, Xpotri_batch_drivers.cuh:31-, Xpoti_batch_drivers.cuh:31-89).
// n <= 32: one launch of kernels/tri_inv.cuh (poti: potrf launch + potri launch).  trtri for n > 32 follows the reference's
// own recursion over TRSM (Xtrtri_batch_drivers.cuh:31-84) on this library's one-launch TRSM; lauum for n > 32 is one
// launch of the blocked in-place kernel (no TRMM / SYRK launches, no workspace); potri = trtri + lauum, poti = potrf + potri.
// Contract as the reference: Lower only, NonUnit only ("(Upper | DIAG) TRTRI_BATCH is not implemented yet",
// Xtrtri_batch_drivers.cuh:96-99), in place, info_array not written, workspace protocol honoured.
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/tri_inv.cuh"
#include "potrf_batch.h"
#include "tri_batch.h"

namespace kblasx {

template <typename T, int NP, int OP, bool STRIDED>
static int launch_tri_inv(KBlasHandle *h, const char *name, int n, BatchRef<T, STRIDED> A, int lda, int batchCount) {
  constexpr int WARPS = 4, MPW = 32 / NP;
  const long wtasks = ((long)batchCount + MPW - 1) / MPW;
  long grid = (wtasks + WARPS - 1) / WARPS;
  const long cap = (long)h->sm_count * 32;
  if (grid > cap) grid = cap;
  tri_inv_kernel<T, NP, OP, WARPS, STRIDED><<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(n, A, lda, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, int OP, bool STRIDED>
static int tri_inv_small(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount) {
  const char *nm = OP == TI_TRTRI ? "tri_inv<trtri>" : OP == TI_LAUUM ? "tri_inv<lauum>" : "tri_inv<potri>";
  if (n <= 8) return launch_tri_inv<T, 8, OP, STRIDED>(h, nm, n, A, lda, batchCount);
  if (n <= 16) return launch_tri_inv<T, 16, OP, STRIDED>(h, nm, n, A, lda, batchCount);
  return launch_tri_inv<T, 32, OP, STRIDED>(h, nm, n, A, lda, batchCount);
}

// sub-matrix (r, c) of every batch entry
template <typename T>
static BatchRef<T, true> sub(BatchRef<T, true> A, int r, int c, int lda) {
  BatchRef<T, true> s = {A.base + r + (long)c * lda, A.stride};
  return s;
}
template <typename T>
static BatchRef<T, false> sub(BatchRef<T, false> A, int r, int c, int lda) {
  BatchRef<T, false> s = {A.base, A.stride + r + (long)c * lda};
  return s;
}
template <typename T, bool STRIDED>
static BatchRef<const T, STRIDED> as_const(BatchRef<T, STRIDED> A) {
  BatchRef<const T, STRIDED> c;
  c.base = A.base;
  c.stride = A.stride;
  return c;
}

// inverse of the lower triangle, any n: [A11 0; A21 A22]^-1 = [A11^-1 0; -A22^-1 A21 A11^-1  A22^-1]
// (the reference's Xtrtri_trsm_rec, Xtrtri_batch_drivers.cuh:31-84, with 32 instead of 16 as the leaf size)
template <typename T, bool STRIDED>
static int trtri_rec(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount) {
  if (n <= 32) return tri_inv_small<T, TI_TRTRI, STRIDED>(h, n, A, lda, batchCount);
  const int n1 = REG_SIZE(n) ? n / 2 : CLOSEST_REG_SIZE(n), n2 = n - n1;
  // A21 := -A21 A11^-1 ;  A21 := A22^-1 A21
  check_ret_error((tri_solve_core<T, STRIDED>(h, /*left=*/false, TRI_BACKWARD, n2, n1, T(-1), as_const(A), lda, sub(A, n1, 0, lda), lda, batchCount)));
  check_ret_error((tri_solve_core<T, STRIDED>(h, /*left=*/true, TRI_FORWARD, n2, n1, T(1), as_const(sub(A, n1, n1, lda)), lda, sub(A, n1, 0, lda), lda,
                                              batchCount)));
  check_ret_error((trtri_rec<T, STRIDED>(h, n1, A, lda, batchCount)));
  return trtri_rec<T, STRIDED>(h, n2, sub(A, n1, n1, lda), lda, batchCount);
}

// A := L^T L for n > 32 (kernels/tri_inv.cuh, lauum_blocked_kernel)
template <typename T, bool STRIDED>
static int lauum_blocked(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * LauumBlockedSmem<T>::per_warp * sizeof(T);
  auto kern = lauum_blocked_kernel<T, WARPS, STRIDED>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)((batchCount + WARPS - 1) / WARPS), WARPS * 32, smem, h->stream>>>(n, A, lda, batchCount);
  h->note_launch("lauum_blocked");
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

enum InvRoutine { R_TRTRI, R_LAUUM, R_POTRI, R_POTI };

template <typename T, bool STRIDED>
static int inv_family(KBlasHandle *h, int routine, char uplo, char diag, int n, BatchRef<T, STRIDED> A, int lda, int batchCount, int *info) {
  if (uplo == KBLAS_Upper || diag == KBLAS_Unit) {
    const char *nm = routine == R_TRTRI ? "(Upper | DIAG) TRTRI" : routine == R_LAUUM ? "Upper LAUUM" : routine == R_POTRI ? "Upper POTRI" : "Upper POTI";
    printf("%s_BATCH is not implemented yet\n", nm);
    return KBLAS_NotImplemented;
  }
  if (batchCount <= 0) {
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);  // empty grid in the reference
  }
  if (n <= 0) return KBLAS_Success;
  // workspace protocol (reference X{trtri,lauum,potri,poti}_batch.cu: *_wsquery_core -> isSufficient)
  KBlasWorkspaceState need;
  if (routine == R_TRTRI) trtri_batch_wsquery_core(STRIDED, n, batchCount, &need);
  else if (routine == R_LAUUM) lauum_batch_wsquery_core(STRIDED, n, batchCount, &need);
  else if (routine == R_POTRI) potri_batch_wsquery_core(STRIDED, n, batchCount, &need);
  else poti_batch_wsquery_core(STRIDED, n, batchCount, &need);
  if (!need.isSufficient(&h->work_space.allocated_ws_state)) return KBLAS_InsufficientWorkspace;
  if (routine == R_TRTRI) return trtri_rec<T, STRIDED>(h, n, A, lda, batchCount);
  if (routine == R_POTI) check_ret_error((potrf_batch_core<T, STRIDED>(h, uplo, n, A, lda, batchCount, info)));
  if (n <= 32) {
    if (routine == R_LAUUM) return tri_inv_small<T, TI_LAUUM, STRIDED>(h, n, A, lda, batchCount);
    return tri_inv_small<T, TI_POTRI, STRIDED>(h, n, A, lda, batchCount);
  }
  // n > 32: potri = trtri (TRSM recursion) + lauum (one blocked launch), as Xpotri_batch_drivers.cuh composes it
  if (routine != R_LAUUM) check_ret_error((trtri_rec<T, STRIDED>(h, n, A, lda, batchCount)));
  return lauum_blocked<T, STRIDED>(h, n, A, lda, batchCount);
}

}  // namespace kblasx

using kblasx::BatchRef;

#define KX_INV_API1(P, T, NAME, ROUTINE, DIAGDECL, DIAGARG)                                                          \
  int kblas_##NAME##_batch(kblasHandle_t handle, char uplo DIAGDECL, const int n, T **A, int lda, int batchCount,     \
                           int *info_array) {                                                                         \
    BatchRef<T, false> a = {A, 0};                                                                                    \
    return kblasx::inv_family<T, false>(handle, kblasx::ROUTINE, uplo, DIAGARG, n, a, lda, batchCount, info_array);   \
  }                                                                                                                   \
  int kblas_##NAME##_batch(kblasHandle_t handle, char uplo DIAGDECL, const int n, T *A, int lda, long strideA,        \
                           int batchCount, int *info_array) {                                                         \
    BatchRef<T, true> a = {A, strideA};                                                                               \
    return kblasx::inv_family<T, true>(handle, kblasx::ROUTINE, uplo, DIAGARG, n, a, lda, batchCount, info_array);    \
  }                                                                                                                   \
  extern "C" int kblas##P##NAME##_batch(kblasHandle_t handle, char uplo DIAGDECL, const int n, T **A, int lda,        \
                                        int batchCount, int *info_array) {                                            \
    BatchRef<T, false> a = {A, 0};                                                                                    \
    return kblasx::inv_family<T, false>(handle, kblasx::ROUTINE, uplo, DIAGARG, n, a, lda, batchCount, info_array);   \
  }                                                                                                                   \
  extern "C" int kblas##P##NAME##_batch_strided(kblasHandle_t handle, char uplo DIAGDECL, const int n, T *A, int lda, \
                                                long strideA, int batchCount, int *info_array) {                      \
    BatchRef<T, true> a = {A, strideA};                                                                               \
    return kblasx::inv_family<T, true>(handle, kblasx::ROUTINE, uplo, DIAGARG, n, a, lda, batchCount, info_array);    \
  }
#define KX_COMMA_DIAG , char diag
#define KX_INV_API(P, T)                                        \
  KX_INV_API1(P, T, trtri, R_TRTRI, KX_COMMA_DIAG, diag)        \
  KX_INV_API1(P, T, lauum, R_LAUUM, , KBLAS_NonUnit)            \
  KX_INV_API1(P, T, potri, R_POTRI, , KBLAS_NonUnit)            \
  KX_INV_API1(P, T, poti, R_POTI, , KBLAS_NonUnit)
KX_INV_API(S, float)
KX_INV_API(D, double)

// workspace queries (reference src/workspace_queries.cu: kblas_{trtri,lauum,potri,poti}_batch[_strided]_wsquery)
#define REQ(h) (&((h)->work_space.requested_ws_state))
#define KX_INV_WS(NAME)                                                                                      \
  void kblas_##NAME##_batch_wsquery(kblasHandle_t handle, const int n, int batchCount) {                     \
    kblasx::NAME##_batch_wsquery_core(false, n, batchCount, REQ(handle));                                    \
  }                                                                                                          \
  void kblas_##NAME##_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount) {             \
    kblasx::NAME##_batch_wsquery_core(true, n, batchCount, REQ(handle));                                     \
  }
KX_INV_WS(trtri)
KX_INV_WS(lauum)
KX_INV_WS(potri)
KX_INV_WS(poti)
