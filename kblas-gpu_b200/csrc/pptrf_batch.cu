// pptrf_batch.cu -- kblasx{S,D}pptrf_batch[_strided]: batched Cholesky on LAPACK packed lower storage, plus the
// pack / unpack kernels between the drop-in column-major layout and the packed one.
//
// No reference counterpart (SURVEY §8(f)4 "interleaved / packed batch layout"; the reference's nearest neighbour is
// batch_pstrf, include/batch_pstrf.h, src/batch_svd/batch_pstrf.cu:226-246 -- pivoted Cholesky on full storage).
// Why it exists: with full column-major storage a B200 moves 12288 bytes of DRAM lines for the 8448 algorithmic
// bytes of a 32 x 32 fp64 lower triangle; in packed storage physical == algorithmic (kernels/potrf_packed.cuh).
// Semantics mirror kblas_potrf_batch (Xpotrf_batch.cu:44-160): Lower only (Upper -> KBLAS_NotImplemented, same
// message), info untouched unless KBLAS_B200_INFO_MODE=lapack, empty batch -> KBLAS_UnknownError, asynchronous on
// handle->stream, n <= 32 (larger -> KBLAS_NotImplemented: unpack and call kblas_potrf_batch).
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/potrf_packed.cuh"
#include "potrf_batch.h"

namespace kblasx {

template <typename T, int NP, int WARPS, int MINB, bool STRIDED, bool IN_BULK, bool OUT_BULK, bool LOCKSTEP>
static int launch_packed(KBlasHandle *h, const char *name, BatchRef<T, STRIDED> AP, int batchCount) {
  auto kern = potrf_packed_kernel<T, NP, WARPS, MINB, STRIDED, IN_BULK, OUT_BULK, LOCKSTEP>;
  const size_t smem = PackedSmem<T, NP, WARPS, IN_BULK, OUT_BULK>::bytes;
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  const long per_cta = (long)WARPS * 4;
  const long need = (batchCount + per_cta - 1) / per_cta;
  const long wave = (long)h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, smem, MINB);
  const long grid = need < wave ? need : wave;
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(AP, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, int NP, int WARPS, int MINB, bool STRIDED>
static int launch_packed_generic(KBlasHandle *h, const char *name, int n, BatchRef<T, STRIDED> AP, int batchCount, int *info) {
  auto kern = potrf_packed_generic_kernel<T, NP, WARPS, MINB, STRIDED>;
  const long per_cta = (long)WARPS * 4;
  const long need = (batchCount + per_cta - 1) / per_cta;
  const long wave = (long)h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, 0, MINB);
  const long grid = need < wave ? need : wave;
  kern<<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(n, AP, batchCount, info, h->info_mode);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// n == N, contiguous batch (strideAP == N(N+1)/2): one lane per matrix (kernels/potrf_packed.cuh)
template <typename T, int N, int WARPS, int MINB>
static int launch_packed_lane(KBlasHandle *h, const char *name, T *AP, int batchCount) {
  auto kern = potrf_packed_lane_kernel<T, N, WARPS, MINB>;
  const size_t smem = (size_t)WARPS * 32 * (packed_size(N) + 1) * sizeof(T);
  check_error_ret(kx_allow_smem(h, kern, smem), KBLAS_CUDA_Error);
  const long per_cta = (long)WARPS * 32;
  const long need = (batchCount + per_cta - 1) / per_cta;
  const long wave = (long)h->sm_count * kx_ctas_per_sm(h, kern, WARPS * 32, smem, MINB);
  const long grid = need < wave ? need : wave;
  kern<<<(unsigned)grid, WARPS * 32, smem, h->stream>>>(AP, batchCount);
  h->note_launch(name);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

// variant_override (env KBLAS_B200_VARIANT, A/B runs): 20 = plain loads + plain stores, 21 = TMA bulk loads + plain
// stores, 22 = TMA in + TMA out with the per-batch CTA barrier of potrf_reg_kernel; default = TMA in + TMA out, warps
// free-running.  Measured on B200, 2^20 matrices, fraction of the measured HBM copy peak (profiles/r02_packed_variants.txt):
//   fp64 n=32: 0.66 / 0.75 / 0.79 / 0.87     n=24: 0.70 / 0.81 / 0.87 / 0.93     n=16: 0.85 (default)
//   fp32 n=32: 0.51 / 0.60 / 0.68 / 0.74     (kblas?potrf_batch_strided on full storage: fp64 0.68 / 0.59 / 0.61, fp32 0.46)
#define KX_STR2(x) #x
#define KX_STR(x) KX_STR2(x)
#define KX_PK(NP, W, MB, IB, OB, LS, NAME) \
  launch_packed<T, NP, W, MB, STRIDED, IB, OB, LS>(h, NAME, AP, batchCount)

template <typename T, int NP, int W, int MB, bool STRIDED>
static int packed_exact(KBlasHandle *h, BatchRef<T, STRIDED> AP, int batchCount, bool aligned) {
  const int v = h->variant_override;
  if (!aligned || v == 20) return KX_PK(NP, W, MB, false, false, true, "potrf_packed<ldg,stg>");
  if (v == 21) return KX_PK(NP, W, MB, true, false, true, "potrf_packed<tma-in,stg>");
  if (v == 22) return KX_PK(NP, W, MB, true, true, true, "potrf_packed<tma-in,tma-out,lockstep>");
  return KX_PK(NP, W, MB, true, true, false, "potrf_packed<tma-in,tma-out>");
}

template <typename T, bool STRIDED>
int pptrf_batch_core(KBlasHandle *h, char uplo, int n, BatchRef<T, STRIDED> AP, int batchCount, int *info, bool aligned) {
  if (uplo == KBLAS_Upper) {
    printf("Upper POTRF_BATCH is not implemented yet\n");
    return KBLAS_NotImplemented;
  }
  if (batchCount <= 0) {
    check_error_ret(cudaErrorInvalidConfiguration, KBLAS_UnknownError);  // kblas_potrf_batch's empty-grid behaviour
  }
  if (n <= 0) return KBLAS_Success;
  if (n > 32) return KBLAS_NotImplemented;
  constexpr bool F32 = sizeof(T) == 4;
  const bool exact = (n % 8 == 0) && h->info_mode == KBLASX_INFO_COMPAT;
  if (exact) {
    if constexpr (STRIDED) {
      // tiny matrices, contiguous batch: one lane per matrix (any A/B override 20..24 selects the 8-lane kernels)
      const bool ab = h->variant_override >= 20 && h->variant_override <= 24;
      if (!ab && AP.stride == (long)packed_size(n)) {
        if (n == 8) return launch_packed_lane<T, 8, 4, 4>(h, "potrf_packed_lane<N=8>", AP.base, batchCount);
        if constexpr (F32) {
          if (n == 16) return launch_packed_lane<T, 16, 4, 3>(h, "potrf_packed_lane<N=16>", AP.base, batchCount);
        }
      }
    }
    if (n == 8) return packed_exact<T, 8, 4, 8, STRIDED>(h, AP, batchCount, aligned);
    if (n == 16) return packed_exact<T, 16, 4, 4, STRIDED>(h, AP, batchCount, aligned);
    if (n == 24) return packed_exact<T, 24, 4, 3, STRIDED>(h, AP, batchCount, aligned);
    if constexpr (F32) return packed_exact<T, 32, 8, 2, STRIDED>(h, AP, batchCount, aligned);
    else return packed_exact<T, 32, 8, 1, STRIDED>(h, AP, batchCount, aligned);
  }
  if (n <= 8) return launch_packed_generic<T, 8, 4, 4, STRIDED>(h, "potrf_packed_generic<NP=8>", n, AP, batchCount, info);
  if (n <= 16) return launch_packed_generic<T, 16, 4, 4, STRIDED>(h, "potrf_packed_generic<NP=16>", n, AP, batchCount, info);
  if (n <= 24) return launch_packed_generic<T, 24, 4, 3, STRIDED>(h, "potrf_packed_generic<NP=24>", n, AP, batchCount, info);
  return launch_packed_generic<T, 32, 4, 2, STRIDED>(h, "potrf_packed_generic<NP=32>", n, AP, batchCount, info);
}

template int pptrf_batch_core<float, true>(KBlasHandle *, char, int, BatchRef<float, true>, int, int *, bool);
template int pptrf_batch_core<double, true>(KBlasHandle *, char, int, BatchRef<double, true>, int, int *, bool);
template int pptrf_batch_core<float, false>(KBlasHandle *, char, int, BatchRef<float, false>, int, int *, bool);
template int pptrf_batch_core<double, false>(KBlasHandle *, char, int, BatchRef<double, false>, int, int *, bool);

template <typename T, bool STRIDED, bool UNPACK>
static int tri_pack(KBlasHandle *h, char uplo, int n, BatchRef<T, STRIDED> A, int lda, BatchRef<T, STRIDED> AP, int batchCount) {
  if (uplo == KBLAS_Upper) return KBLAS_NotImplemented;
  if (batchCount <= 0 || n <= 0) return KBLAS_Success;
  if (lda < n) return KBLAS_Error_WrongInput;
  long blocks = ((long)batchCount + 7) / 8;
  const long cap = (long)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  tri_pack_kernel<T, STRIDED, UNPACK><<<(unsigned)blocks, 256, 0, h->stream>>>(n, A, lda, AP, batchCount);
  h->note_launch(UNPACK ? "tri_unpack" : "tri_pack");
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

}  // namespace kblasx

// ---- C ABI (include/kblas_ffi.h) ---------------------------------------------------------------------------------
#define KX_PPTRF_API(P, T)                                                                                         \
  extern "C" int kblasx##P##pptrf_batch_strided(kblasHandle_t handle, char uplo, int n, T *AP, long strideAP,      \
                                                int batchCount, int *info_array) {                                 \
    if (n > 0 && strideAP < (long)kblasx::packed_size(n)) return KBLAS_Error_WrongInput;                           \
    const bool aligned = (reinterpret_cast<size_t>(AP) % 16 == 0) && (((size_t)strideAP * sizeof(T)) % 16 == 0);   \
    kblasx::BatchRef<T, true> ref = {AP, strideAP};                                                                \
    return kblasx::pptrf_batch_core<T, true>(handle, uplo, n, ref, batchCount, info_array, aligned);               \
  }                                                                                                                \
  /* pointer array: the entries' alignment is unknown on the host -> plain loads / stores */                       \
  extern "C" int kblasx##P##pptrf_batch(kblasHandle_t handle, char uplo, int n, T **AP_array, int batchCount,      \
                                        int *info_array) {                                                         \
    kblasx::BatchRef<T, false> ref = {AP_array, 0};                                                                \
    return kblasx::pptrf_batch_core<T, false>(handle, uplo, n, ref, batchCount, info_array, false);                \
  }                                                                                                                \
  extern "C" int kblasx##P##tri_pack_batch_strided(kblasHandle_t handle, char uplo, int n, const T *A, int lda,    \
                                                   long strideA, T *AP, long strideAP, int batchCount) {           \
    kblasx::BatchRef<T, true> a = {const_cast<T *>(A), strideA}, ap = {AP, strideAP};                              \
    return kblasx::tri_pack<T, true, false>(handle, uplo, n, a, lda, ap, batchCount);                              \
  }                                                                                                                \
  extern "C" int kblasx##P##tri_unpack_batch_strided(kblasHandle_t handle, char uplo, int n, const T *AP,          \
                                                     long strideAP, T *A, int lda, long strideA, int batchCount) { \
    kblasx::BatchRef<T, true> a = {A, strideA}, ap = {const_cast<T *>(AP), strideAP};                              \
    return kblasx::tri_pack<T, true, true>(handle, uplo, n, a, lda, ap, batchCount);                               \
  }
KX_PPTRF_API(S, float)
KX_PPTRF_API(D, double)
