// workspace_queries.cu -- kblas_*_batch[_strided]_wsquery entry points and their cores.
//
// The sm_100a kernels of this library need no scratch memory.  The queries are kept
// byte-for-byte compatible with the reference anyway, because callers (and the reference's
// own test binaries) follow the protocol  wsquery -> kblasAllocateWorkspace -> call  and a
// call made without it returns KBLAS_InsufficientWorkspace in the reference
// (src/batch_triangular/Xpotrf_batch.cu:54-56).  Arithmetic restated from
// src/workspace_queries.ch:60-74,111-122,157-171,194-201 and
// src/workspace_queries.cu:188-208,227-238 (built without MAGMA, CUDA >= 8).
#include "kblas.h"
#include "kblas_common.h"

namespace kblasx {

// reference src/workspace_queries.cu:188-194: 3 pointer arrays for an offset (sub-matrix) GEMM
void gemm_batch_offset_wsquery_core(int batchCount, bool offseted, KBlasWorkspaceState *ws) {
  if (offseted)
    ws->d_ptrs_bytes = kblasx_max((size_t)(batchCount > 1) * size_t(batchCount) * 3 * sizeof(void *), ws->d_ptrs_bytes);
}

// reference src/workspace_queries.cu:203-208: nothing with CUDA >= 8 and no MAGMA
void gemm_batch_strided_wsquery_core(int /*batchCount*/, KBlasWorkspaceState * /*ws*/) {}

// reference src/workspace_queries.cu:227-238: pointer triples for the flattened SYRK recursion
void syrk_batch_wsquery_core(int m, int batchCount, KBlasWorkspaceState *ws) {
  if (m > 16) {
    int depth = 0, s = 16;
    while (s < m) {
      s <<= 1;
      depth++;
    }
    ws->d_ptrs_bytes = kblasx_max(size_t(1 << (depth - 1)) * batchCount * 3 * sizeof(void *), ws->d_ptrs_bytes);
  }
}

// reference src/workspace_queries.ch:60-74
void trsm_batch_wsquery_core(bool strided, int batchCount, char side, int m, int n, KBlasWorkspaceState *ws) {
  if (((side == KBLAS_Right) && (n > 16)) || ((side == KBLAS_Left) && (m > 16))) {
    if (strided)
      gemm_batch_strided_wsquery_core(batchCount, ws);
    else
      gemm_batch_offset_wsquery_core(batchCount, true, ws);
  }
}

// reference src/workspace_queries.ch:111-122
void potrf_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws) {
  int n1 = CLOSEST_REG_SIZE(n);
  trsm_batch_wsquery_core(strided, batchCount, KBLAS_Right, n - n1, n1, ws);
  syrk_batch_wsquery_core(n - n1, batchCount, ws);
}

// reference src/workspace_queries.ch:157-171
void potrs_batch_wsquery_core(bool strided, int m, int n, int batchCount, KBlasWorkspaceState *ws) {
  int n1 = CLOSEST_REG_SIZE(n);
  trsm_batch_wsquery_core(strided, batchCount, KBLAS_Right, m, n1, ws);
  if (strided)
    gemm_batch_strided_wsquery_core(batchCount, ws);
  else
    gemm_batch_offset_wsquery_core(batchCount, true, ws);
}

// reference src/workspace_queries.ch:80-97 (no MAGMA: only the GEMM part of the TRMM recursion)
void trmm_batch_wsquery_core(bool strided, int batchCount, char side, int m, int n, KBlasWorkspaceState *ws) {
  if (((side == KBLAS_Right) && (n > 16)) || ((side == KBLAS_Left) && (m > 16))) {
    if (strided)
      gemm_batch_strided_wsquery_core(batchCount, ws);
    else
      gemm_batch_offset_wsquery_core(batchCount, true, ws);
  }
}

// reference src/workspace_queries.ch:127-136
void lauum_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws) {
  int n1 = CLOSEST_REG_SIZE(n);
  trmm_batch_wsquery_core(strided, batchCount, KBLAS_Left, n - n1, n1, ws);
  syrk_batch_wsquery_core(n1, batchCount, ws);
}

// reference src/workspace_queries.ch:141-154
void trtri_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws) {
  if (n > 16) {
    int n1 = CLOSEST_REG_SIZE(n);
    trsm_batch_wsquery_core(strided, batchCount, KBLAS_Left, n - n1, n1, ws);
    trsm_batch_wsquery_core(strided, batchCount, KBLAS_Right, n - n1, n1, ws);
  }
}

// reference src/workspace_queries.ch:176-191
void potri_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws) {
  trtri_batch_wsquery_core(strided, n, batchCount, ws);
  lauum_batch_wsquery_core(strided, n, batchCount, ws);
}
void poti_batch_wsquery_core(bool strided, int n, int batchCount, KBlasWorkspaceState *ws) {
  potrf_batch_wsquery_core(strided, n, batchCount, ws);
  potri_batch_wsquery_core(strided, n, batchCount, ws);
}

// reference src/workspace_queries.ch:194-201
void posv_batch_wsquery_core(bool strided, int m, int n, char side, int batchCount, KBlasWorkspaceState *ws) {
  potrf_batch_wsquery_core(strided, (side == KBLAS_Right) ? n : m, batchCount, ws);
  potrs_batch_wsquery_core(strided, m, n, batchCount, ws);
}

}  // namespace kblasx

// ---- public C++-linkage entry points (reference src/workspace_queries.cu:257-266,313-319,340-346,367-373)
#define REQ(h) (&((h)->work_space.requested_ws_state))

void kblas_trsm_batch_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount) {
  kblasx::trsm_batch_wsquery_core(false, batchCount, side, m, n, REQ(handle));
}
void kblas_trsm_batch_strided_wsquery(kblasHandle_t handle, char side, int m, int n, int batchCount) {
  kblasx::trsm_batch_wsquery_core(true, batchCount, side, m, n, REQ(handle));
}
// non-uniform TRSM exists in the reference only through MAGMA (Xtrsm_batch_drivers.cuh:277-367); without
// USE_MAGMA its query records nothing (src/workspace_queries.cu:268-277) -- same here
void kblas_trsm_batch_nonuniform_wsquery(kblasHandle_t /*handle*/) {}
void kblas_potrf_batch_wsquery(kblasHandle_t handle, const int n, int batchCount) {
  kblasx::potrf_batch_wsquery_core(false, n, batchCount, REQ(handle));
}
void kblas_potrf_batch_strided_wsquery(kblasHandle_t handle, const int n, int batchCount) {
  kblasx::potrf_batch_wsquery_core(true, n, batchCount, REQ(handle));
}
void kblas_potrs_batch_wsquery(kblasHandle_t handle, const int m, const int n, int batchCount) {
  kblasx::potrs_batch_wsquery_core(false, m, n, batchCount, REQ(handle));
}
void kblas_potrs_batch_strided_wsquery(kblasHandle_t handle, const int m, const int n, int batchCount) {
  kblasx::potrs_batch_wsquery_core(true, m, n, batchCount, REQ(handle));
}
void kblas_posv_batch_wsquery(kblasHandle_t handle, char side, const int m, const int n, int batchCount) {
  kblasx::posv_batch_wsquery_core(false, m, n, side, batchCount, REQ(handle));
}
void kblas_posv_batch_strided_wsquery(kblasHandle_t handle, char side, const int m, const int n, int batchCount) {
  kblasx::posv_batch_wsquery_core(true, m, n, side, batchCount, REQ(handle));
}
