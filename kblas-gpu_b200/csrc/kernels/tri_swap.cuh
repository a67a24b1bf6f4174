// kernels/tri_swap.cuh -- in-place transpose of the n x n leading block of every matrix of a batch (sm_100a).
//
// Used for uplo = Upper in potrf (KBLAS_NotImplemented in the reference, Xpotrf_batch_drivers.cuh:38-41; extension, SURVEY.md
// §8(f)3): transpose, factor the lower triangle with the kernels of this library, transpose back.  The upper triangle then
// holds U = L^T and the strictly lower triangle its original contents ("not referenced"), at the price of two extra passes
// over the matrices -- the Upper form is correct, not fast.
// One warp per matrix, 32 x 32 tiles through padded shared memory: both the reads and the writes are coalesced.
#pragma once

#include "common.cuh"

namespace kblasx {

template <typename T>
struct TriSwapSmem {
  static constexpr int NB = 32, P = 33;
  static constexpr int per_warp = 2 * NB * P;
};

template <typename T, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
transpose_inplace_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount) {
  constexpr int NB = 32, P = TriSwapSmem<T>::P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *SX = reinterpret_cast<T *>(smem_raw) + warp * TriSwapSmem<T>::per_warp;
  T *SY = SX + NB * P;
  const long mat = (long)blockIdx.x * WARPS + warp;
  if (mat >= (long)batchCount) return;  // warp-uniform
  T *A = Aref.at(mat);
  const int nblk = (n + NB - 1) / NB;
  for (int J = 0; J < nblk; ++J) {
    const int j0 = J * NB, jb = (n - j0 < NB) ? (n - j0) : NB;
    for (int I = 0; I <= J; ++I) {
      const int i0 = I * NB, ib = (n - i0 < NB) ? (n - i0) : NB;
      // X = block (I, J), Y = block (J, I); lane = row of the block, stored as S[c * P + r]
#pragma unroll 8
      for (int c = 0; c < NB; ++c) {
        if (lane < ib && c < jb) SX[c * P + lane] = A[(i0 + lane) + (long)(j0 + c) * lda];
        if (I < J && lane < jb && c < ib) SY[c * P + lane] = A[(j0 + lane) + (long)(i0 + c) * lda];
      }
      __syncwarp();
      // block (J, I) := X^T, block (I, J) := Y^T
#pragma unroll 8
      for (int c = 0; c < NB; ++c) {
        if (lane < jb && c < ib) A[(j0 + lane) + (long)(i0 + c) * lda] = SX[lane * P + c];
        if (I < J && lane < ib && c < jb) A[(i0 + lane) + (long)(j0 + c) * lda] = SY[lane * P + c];
      }
      __syncwarp();
    }
  }
}

}  // namespace kblasx
