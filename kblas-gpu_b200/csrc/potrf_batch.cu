// potrf_batch.cu -- kblas_potrf_batch: entry points + dispatch.
//
// Counterpart of reference src/batch_triangular/Xpotrf_batch.cu:44-160 (entry points,
// workspace check) and Xpotrf_batch_drivers.cuh:30-137 (driver).  The reference's driver
// recursion (n/2 split -> potrf, trsm, syrk, potrf launches) is gone: one kernel per call.
#include <cstdlib>
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/potrf_small.cuh"
#include "kernels/potrf_panel_mma.cuh"
#include "kernels/potrf_smem.cuh"
#include "kernels/tri_swap.cuh"
#include "potrf_batch.h"

namespace kblasx {

template <typename T, int NP, int G, int WARPS, int MINB, bool STRIDED, bool EXACT, bool LOCKSTEP>
static int launch_potrf_reg(KBlasHandle *h, const char *name, int n, BatchRef<T, STRIDED> A, int lda, int batchCount,
                            int *info) {
  constexpr int MPW = 32 / G;
  const long per_cta = (long)WARPS * MPW;
  auto kern = potrf_reg_kernel<T, NP, G, WARPS, MINB, STRIDED, EXACT, LOCKSTEP>;
  // persistent CTAs: one resident wave, each CTA strides over its share of the batch
  const long need = (batchCount + per_cta - 1) / per_cta;
  const long wave = (long)h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, 0, MINB);
  const long grid = need < wave ? need : wave;
  kern<<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(n, A, lda, batchCount, info, h->info_mode, 0u);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

#define KX_STR2(x) #x
#define KX_STR(x) KX_STR2(x)
#define KX_LAUNCH_REG(NP, G, W, MB, EX, LS)                                                              \
  launch_potrf_reg<T, NP, G, W, MB, STRIDED, EX, LS>(                                                    \
      h, "potrf_reg<NP=" KX_STR(NP) ",G=" KX_STR(G) ",W=" KX_STR(W) ",MB=" KX_STR(MB) ",EX=" KX_STR(EX) ",LS=" KX_STR(LS) ">", \
      n, A, lda, batchCount, info)

// n <= 32: register-resident kernel, padded size NP = roundup(n, 8).  The EXACT instantiation
// (n == NP, info untouched) carries no bounds predicates; everything else takes the generic one.
// KBLAS_B200_ELEMENT_EXACT_STORES=1 (handle->exact_stores) also selects the generic one: it never writes a
// strict-upper element, for callers that update the upper triangle concurrently.
template <typename T, bool STRIDED>
static int potrf_small_dispatch(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount, int *info) {
  const bool exact = (n % 8 == 0) && (h->info_mode == KBLASX_INFO_COMPAT) && !h->exact_stores;
  constexpr bool F32 = sizeof(T) == 4;
  if (n <= 8) return exact ? KX_LAUNCH_REG(8, 8, 4, 4, true, false) : KX_LAUNCH_REG(8, 8, 4, 4, false, false);
  if (n <= 16) {
    if (!exact) return KX_LAUNCH_REG(16, 8, 4, 4, false, true);
    // measured (B200, 2^20 matrices, ms): fp64 lockstep 3 / 4 / 6 CTAs per SM 0.572 / 0.587 / 0.672, free-running 0.600;
    // fp32 lockstep 0.439, free-running 0.419
    return F32 ? KX_LAUNCH_REG(16, 8, 4, 4, true, false) : KX_LAUNCH_REG(16, 8, 4, 3, true, true);
  }
  if (n <= 24) {
    if (!exact) return KX_LAUNCH_REG(24, 8, 4, 3, false, true);
    // measured (B200, batch 2^20): fp64 one 8-warp lockstep CTA per SM 0.59 vs 3 x 4 warps 0.52;
    // fp32 the other way round (0.45 vs 0.49)
    // (round 2, 16-byte broadcast vectors: fp32 3 x 4 warps free-running 0.788 against 0.818 in lockstep; fp64 unchanged)
    return F32 ? KX_LAUNCH_REG(24, 8, 4, 3, true, false) : KX_LAUNCH_REG(24, 8, 8, 1, true, true);
  }
  if constexpr (F32) {
    // measured (B200, 2^20 matrices, 16-byte broadcast vectors, ms best / mean): 2 x 8 warps (128 registers, 96 B of spills)
    // 1.42 / 1.44, 1 x 8 warps (255) 1.35 / 1.37, 3 x 4 warps (168, 32 B) in lockstep 1.31 / 1.32, free-running 1.26 / 1.37
    return exact ? KX_LAUNCH_REG(32, 8, 4, 3, true, true) : KX_LAUNCH_REG(32, 8, 4, 4, false, true);
  } else {
    // fp64: 160 registers of matrix data per lane -> one 8-warp CTA per SM, warps in lockstep
    return exact ? KX_LAUNCH_REG(32, 8, 8, 1, true, true) : KX_LAUNCH_REG(32, 8, 4, 2, false, true);
  }
}

// n > 32: left-looking 32-column panels, one warp per matrix, update on the mma.sync tensor path -- DMMA for
// fp64, 3 x TF32 for fp32 (kernels/potrf_panel_mma.cuh)
template <typename T, int THREADS, bool STRIDED>
static int launch_potrf_panel_mma(KBlasHandle *h, const char *name, int n, BatchRef<T, STRIDED> A, int lda, int batchCount,
                                  int *info) {
  auto kern = potrf_panel_mma_kernel<T, THREADS, STRIDED>;
  const size_t smem = PanelMmaSmem<T, THREADS>::bytes;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)batchCount, THREADS, smem, h->stream>>>(n, A, lda, batchCount, info, h->info_mode);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// 32 < n <= 256, fp64: one CTA per matrix, factor resident in shared memory (kernels/potrf_smem.cuh)
template <int WARPS, int MINB, bool STRIDED>
static int launch_potrf_smem(KBlasHandle *h, const char *name, int n, BatchRef<double, STRIDED> A, int lda, int batchCount,
                             int *info) {
  auto kern = potrf_smem_kernel<WARPS, MINB, STRIDED>;
  const SmemPotrfPlan plan = plan_potrf_slots((n + 31) / 32);
  const size_t smem = PotrfSmemGeom::bytes(plan.nslots);
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)batchCount, WARPS * 32, smem, h->stream>>>(n, A, lda, batchCount, info, h->info_mode, plan);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, bool STRIDED>
static int potrf_panel_dispatch(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount, int *info) {
  if constexpr (sizeof(T) == 8) {
    // KBLAS_B200_VARIANT = 31 / 34 (2 warps), 32 / 35 (4 warps), 33 (8 warps per matrix): the shared-memory resident kernel
    // (kernels/potrf_smem.cuh).  Measured on B200 (batch 64K, pointer array, profiles/r02_large_n_variants.txt) it reads
    // every element from DRAM once (1.1x compulsory against 4.6x), but is SLOWER than the one-warp-per-matrix kernel below
    // -- n = 64 / 128 / 256: 1.48 / 5.24 / 31.4 ms against 1.05 / 4.80 / 25.5 ms -- because one matrix per CTA has nothing to
    // overlap the 32-step dependent chains of the diagonal factorisation (7-10K cycles) and of the row solves (4-6K cycles
    // per block) with (profiles/r02_smem_phase_trace.txt), which eight independent matrices per SM overlap for free.  Opt-in.
    const int v = h->variant_override;
    if (n <= 256 && v >= 31 && v <= 36) {
      if (v == 36) return launch_potrf_smem<1, 8, STRIDED>(h, "potrf_smem<W=1>", n, A, lda, batchCount, info);
      if (v == 31) return launch_potrf_smem<2, 8, STRIDED>(h, "potrf_smem<W=2,MB=8>", n, A, lda, batchCount, info);
      if (v == 34) return launch_potrf_smem<2, 4, STRIDED>(h, "potrf_smem<W=2,MB=4>", n, A, lda, batchCount, info);
      if (v == 32) return launch_potrf_smem<4, 4, STRIDED>(h, "potrf_smem<W=4,MB=4>", n, A, lda, batchCount, info);
      if (v == 35) return launch_potrf_smem<4, 2, STRIDED>(h, "potrf_smem<W=4,MB=2>", n, A, lda, batchCount, info);
      return launch_potrf_smem<8, 1, STRIDED>(h, "potrf_smem<W=8>", n, A, lda, batchCount, info);
    }
  }
  // warps per matrix: few warps -> more matrices in flight per SM, which is what hides the serial
  // pivot chain of the diagonal blocks.
  // measured (B200, fp64, batch 64K, n = 64 / 128 / 256, 128-register cap): 1 warp 4.8 / 8.5 / 10.8 TFLOP/s,
  // 2 warps 2.9 / 6.1 / 9.7, 4 warps 1.5 / 4.5 / 8.9, 8 warps 0.8 / 2.5 / 6.6.  One warp with the
  // cap lifted to 255 registers (8 resident warps per SM, no spills): 5.5 / 9.5 / 14.6; 2 / 4 warps per matrix at 255
  // registers (round 2): 3.3 / 6.9 / 12.6 and 1.7 / 4.5 / 10.7 -- one warp per matrix stays.  Forming the update of TWO 32-row
  // slabs per pass (the panel's own fragments fetched once per 64 rows, 25 % fewer loads per DMMA; 242 registers, no spills)
  // changed nothing either: n = 128 / 256 4.95 / 13.29 ms against 4.89 / 13.08 -- the update is not bound by its load count
  return launch_potrf_panel_mma<T, 32, STRIDED>(h, sizeof(T) == 8 ? "potrf_panel_dmma<T=32>" : "potrf_panel_tf32x3<T=32>", n, A, lda,
                                                batchCount, info);
}

// A := A^T on the n x n leading blocks (kernels/tri_swap.cuh)
template <typename T, bool STRIDED>
static int transpose_inplace(KBlasHandle *h, int n, BatchRef<T, STRIDED> A, int lda, int batchCount) {
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * TriSwapSmem<T>::per_warp * sizeof(T);
  auto kern = transpose_inplace_kernel<T, WARPS, STRIDED>;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  kern<<<(unsigned)((batchCount + WARPS - 1) / WARPS), WARPS * 32, smem, h->stream>>>(n, A, lda, batchCount);
  h->note_launch("transpose_inplace");
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// Xpotrf_batch_core of the reference (Xpotrf_batch_drivers.cuh:30-137)
template <typename T, bool STRIDED>
int potrf_batch_core(KBlasHandle *h, char uplo, int n, BatchRef<T, STRIDED> A, int lda, int batchCount, int *info) {
  if (uplo == KBLAS_Upper) {
    // KBLAS_NotImplemented in the reference (:38-41).  Extension: A = U^T U through the lower kernels on
    // the transposed matrix -- transpose, factor, transpose back; the strictly lower triangle keeps its contents.
    if (batchCount <= 0) {
      check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);
    }
    if (n <= 0) return KBLAS_Success;
    check_ret_error((transpose_inplace<T, STRIDED>(h, n, A, lda, batchCount)));
    const int rc = potrf_batch_core<T, STRIDED>(h, KBLAS_Lower, n, A, lda, batchCount, info);
    const char *factor_kernel = h->last_kernel;
    const int rc2 = transpose_inplace<T, STRIDED>(h, n, A, lda, batchCount);  // also after a failed launch: A goes back to its layout
    h->last_kernel = factor_kernel;
    return rc != KBLAS_Success ? rc : rc2;
  }
  if (batchCount <= 0) {
    // reference: grid.x == 0 -> launch error -> KBLAS_UnknownError (drivers.cuh:82-88)
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);
  }
  if (n <= 0) return KBLAS_Success;  // reference launches a kernel that does nothing
  if (n <= 32) return potrf_small_dispatch<T, STRIDED>(h, n, A, lda, batchCount, info);
  return potrf_panel_dispatch<T, STRIDED>(h, n, A, lda, batchCount, info);
}

template int potrf_batch_core<float, true>(KBlasHandle *, char, int, BatchRef<float, true>, int, int, int *);
template int potrf_batch_core<float, false>(KBlasHandle *, char, int, BatchRef<float, false>, int, int, int *);
template int potrf_batch_core<double, true>(KBlasHandle *, char, int, BatchRef<double, true>, int, int, int *);
template int potrf_batch_core<double, false>(KBlasHandle *, char, int, BatchRef<double, false>, int, int, int *);

// workspace check of Xpotrf_batch_offset (Xpotrf_batch.cu:50-56, 113-119)
static int potrf_ws_check(KBlasHandle *h, bool strided, int n, int batchCount) {
  KBlasWorkspaceState need;
  potrf_batch_wsquery_core(strided, n, batchCount, &need);
  return need.isSufficient(&h->work_space.allocated_ws_state) ? KBLAS_Success : KBLAS_InsufficientWorkspace;
}

template <typename T>
int potrf_batch_strided(KBlasHandle *h, char uplo, int n, T *A, int lda, long strideA, int batchCount, int *info) {
  if (potrf_ws_check(h, true, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<T, true> ref = {A, strideA};
  return potrf_batch_core<T, true>(h, uplo, n, ref, lda, batchCount, info);
}

template <typename T>
int potrf_batch_ptrs(KBlasHandle *h, char uplo, int n, T **A, long elem_off, int lda, int batchCount, int *info) {
  if (potrf_ws_check(h, false, n, batchCount) != KBLAS_Success) return KBLAS_InsufficientWorkspace;
  BatchRef<T, false> ref = {A, elem_off};
  return potrf_batch_core<T, false>(h, uplo, n, ref, lda, batchCount, info);
}

}  // namespace kblasx

// ---- public API: C++ overloads (reference Xpotrf_batch.cu:67-80,130-143) and C names (:86-98,149-160)
#define KX_POTRF_API(P, T)                                                                                    \
  int kblas_potrf_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda, int batchCount,         \
                        int *info_array) {                                                                    \
    return kblasx::potrf_batch_ptrs<T>(handle, uplo, n, A, 0, lda, batchCount, info_array);                      \
  }                                                                                                           \
  int kblas_potrf_batch(kblasHandle_t handle, char uplo, const int n, T *A, int lda, long strideA,            \
                        int batchCount, int *info_array) {                                                    \
    return kblasx::potrf_batch_strided<T>(handle, uplo, n, A, lda, strideA, batchCount, info_array);          \
  }                                                                                                           \
  extern "C" int kblas##P##potrf_batch(kblasHandle_t handle, char uplo, const int n, T **A, int lda,          \
                                       int batchCount, int *info_array) {                                     \
    return kblasx::potrf_batch_ptrs<T>(handle, uplo, n, A, 0, lda, batchCount, info_array);                      \
  }                                                                                                           \
  extern "C" int kblas##P##potrf_batch_strided(kblasHandle_t handle, char uplo, const int n, T *A, int lda,   \
                                               long strideA, int batchCount, int *info_array) {               \
    return kblasx::potrf_batch_strided<T>(handle, uplo, n, A, lda, strideA, batchCount, info_array);          \
  }
// internal C++ entry points with sub-matrix offsets that sibling routines and tests of the reference link against
// (reference Xpotrf_batch.cu:44-63 pointer array, 107-127 strided; declared in src/Xblas_core.ch:249-262)
#define KX_POTRF_OFFSET_API(T)                                                                                  \
  int Xpotrf_batch_offset(kblasHandle_t handle, char uplo, const int n, T **A, int A_row_off, int A_col_off,    \
                          int lda, int batchCount, int *info_array) {                                           \
    return kblasx::potrf_batch_ptrs<T>(handle, uplo, n, A, A_row_off + (long)A_col_off * lda, lda, batchCount,  \
                                       info_array);                                                             \
  }                                                                                                             \
  int Xpotrf_batch_offset(kblasHandle_t handle, char uplo, const int n, T *A, int A_row_off, int A_col_off,     \
                          int lda, long strideA, int batchCount, int *info_array) {                             \
    return kblasx::potrf_batch_strided<T>(handle, uplo, n, A + A_row_off + (long)A_col_off * lda, lda, strideA, \
                                          batchCount, info_array);                                              \
  }
KX_POTRF_OFFSET_API(float)
KX_POTRF_OFFSET_API(double)

KX_POTRF_API(S, float)
KX_POTRF_API(D, double)
