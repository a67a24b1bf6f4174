// kernels/gemm_tile.cuh -- batched small GEMM / lower SYRK on the mma.sync fragment engine (sm_100a).
//
// the trailing-update step of the path exposed as public entry points, so that the library is
// self-contained without cuBLAS.  The reference's kblas_gemm_batch is a wrapper around cublas?gemmBatched /
// ?gemmStridedBatched (Xgemm_batch_core.cuh:170-313, 492-634) and its kblas_syrk_batch runs 8 / 16-wide register
// kernels on the diagonal blocks plus one batched GEMM per recursion level for the rest, through pointer fix-up
// kernels (Xsyrk_batch_drivers.cuh:32-433).  Here: ONE launch; one warp owns a 32 x 32 tile of C of one matrix and walks
// the inner dimension with
//   fp64: DMMA  m8n8k4  (4 x 4 tiles, 16 accumulator pairs),
//   fp32: 3 x TF32 m16n8k8 on hi/lo split operands (fp32-grade accuracy, profiles/r01_accuracy_fp32_tf32x3.txt)
// -- the same fragment code as the Cholesky panel update (kernels/potrf_panel_mma.cuh).  Operands are fetched in
// fragment order straight from global memory (every load instruction covers 8 rows x 4 columns = whole sectors for
// the non-transposed operand), out-of-range elements read as zero, so every m, n, k, transpose combination and
// leading dimension works.  SYRK computes only the tiles on or below the diagonal and stores only i >= j.
#pragma once

#include "common.cuh"
#include "potrf_panel_mma.cuh"

namespace kblasx {

// element (i, kk) of op(A), A stored column-major with leading dimension ld; 0 outside [0, rows) x [0, depth)
template <typename T, bool TRANS>
__device__ __forceinline__ T op_elem(const T *__restrict__ A, int ld, int i, int kk, int rows, int depth) {
  T v = T(0);
  const bool ok = i < rows && kk < depth;
  const long idx = TRANS ? ((long)kk + (long)i * ld) : ((long)i + (long)kk * ld);
  ldg_stream_if(v, A + (ok ? idx : 0), ok);
  return v;
}

// acc(32 x 32 tile at (r0, c0)) = sum_kk opA(r0 + i, kk) * opB'(c0 + j, kk),  opB' = "row j of the B side":
// GEMM: opB'(j, kk) = op(B)(kk, j);  SYRK: opB' = opA.
template <bool TA, bool TBP>
__device__ __forceinline__ void tile_mma(double (&acc)[4][4][2], const double *__restrict__ A, int lda, const double *__restrict__ Bp,
                                         int ldb, int r0, int c0, int rows, int cols, int depth, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
  for (int k0 = 0; k0 < depth; k0 += 4) {
    double af[4], bf[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      af[b] = op_elem<double, TA>(A, lda, r0 + 8 * b + fr, k0 + fk, rows, depth);
      bf[b] = op_elem<double, TBP>(Bp, ldb, c0 + 8 * b + fr, k0 + fk, cols, depth);
    }
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af[rb], bf[cb]);
  }
}

template <bool TA, bool TBP>
__device__ __forceinline__ void tile_mma(float (&acc)[2][4][4], const float *__restrict__ A, int lda, const float *__restrict__ Bp,
                                         int ldb, int r0, int c0, int rows, int cols, int depth, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  for (int k0 = 0; k0 < depth; k0 += 8) {
    unsigned ah[4][2], al[4][2], bh[4][2], bl[4][2];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        split_tf32(op_elem<float, TA>(A, lda, r0 + 8 * b + fr, k0 + fk + 4 * h, rows, depth), ah[b][h], al[b][h]);
        split_tf32(op_elem<float, TBP>(Bp, ldb, c0 + 8 * b + fr, k0 + fk + 4 * h, cols, depth), bh[b][h], bl[b][h]);
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        mma_tf32_m16n8k8(acc[mt][nt], al[2 * mt][0], al[2 * mt + 1][0], al[2 * mt][1], al[2 * mt + 1][1], bh[nt][0], bh[nt][1]);
        mma_tf32_m16n8k8(acc[mt][nt], ah[2 * mt][0], ah[2 * mt + 1][0], ah[2 * mt][1], ah[2 * mt + 1][1], bl[nt][0], bl[nt][1]);
        mma_tf32_m16n8k8(acc[mt][nt], ah[2 * mt][0], ah[2 * mt + 1][0], ah[2 * mt][1], ah[2 * mt + 1][1], bh[nt][0], bh[nt][1]);
      }
  }
}

// C(i, j) = alpha * acc + beta * C(i, j) for one accumulator element (beta == 0: C is not read, BLAS semantics)
template <typename T>
__device__ __forceinline__ void tile_store(T *__restrict__ C, int ldc, int i, int j, int rows, int cols, bool lower_only, T alpha, T beta,
                                           T v) {
  if (i < rows && j < cols && (!lower_only || i >= j)) {
    T *p = C + i + (long)j * ldc;
    *p = (beta == T(0)) ? alpha * v : fma_t(beta, *p, alpha * v);
  }
}

__device__ __forceinline__ void tile_epilogue(const double (&acc)[4][4][2], double *C, int ldc, int r0, int c0, int rows, int cols,
                                              bool lower_only, double alpha, double beta, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        tile_store<double>(C, ldc, r0 + 8 * rb + fr, c0 + 8 * cb + 2 * fk + e, rows, cols, lower_only, alpha, beta, acc[rb][cb][e]);
}
__device__ __forceinline__ void tile_epilogue(const float (&acc)[2][4][4], float *C, int ldc, int r0, int c0, int rows, int cols,
                                              bool lower_only, float alpha, float beta, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        tile_store<float>(C, ldc, r0 + 16 * mt + fr + 8 * (e >> 1), c0 + 8 * nt + 2 * fk + (e & 1), rows, cols, lower_only, alpha, beta,
                          acc[mt][nt][e]);
}

template <typename T> struct TileAcc;
template <> struct TileAcc<double> { typedef double type[4][4][2]; };
template <> struct TileAcc<float> { typedef float type[2][4][4]; };

// GEMM: C(m x n) = alpha op(A) op(B) + beta C.   SYRK (lower): C(m x m) = alpha op(A) op(A)^T + beta C, inner dimension kdim.
// task = (matrix, tile row, tile column); SYRK enumerates only tiles with tile row >= tile column.
template <typename T, bool STRIDED, bool SYRK, bool TA, bool TB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gemm_tile_kernel(const int m, const int n, const int kdim, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda,
                 BatchRef<const T, STRIDED> Bref, const int ldb, const T beta, BatchRef<T, STRIDED> Cref, const int ldc,
                 const int batchCount) {
  const int lane = threadIdx.x & 31;
  const int tr = (m + 31) / 32, tc = (n + 31) / 32;
  const long tiles = SYRK ? ((long)tr * (tr + 1)) / 2 : (long)tr * tc;
  const long ntask = tiles * batchCount;
  for (long task = (long)blockIdx.x * WARPS + (threadIdx.x >> 5); task < ntask; task += (long)gridDim.x * WARPS) {
    const long b = task / tiles;
    int t = (int)(task % tiles), ti, tj;
    if (SYRK) {  // t = ti (ti + 1) / 2 + tj, tj <= ti
      ti = 0;
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      tj = t - ti * (ti + 1) / 2;
    } else {
      ti = t % tr;
      tj = t / tr;
    }
    const T *A = Aref.at(b);
    const T *Bp = SYRK ? A : Bref.at(b);
    typename TileAcc<T>::type acc;
    // the "B side" is indexed (column of C, kk): op(B)(kk, j) -> for TB == false that is B[kk + j*ldb], i.e. the
    // TRANSPOSED access pattern of op_elem; SYRK uses op(A) itself
    tile_mma<TA, SYRK ? TA : !TB>(acc, A, lda, Bp, SYRK ? lda : ldb, 32 * ti, 32 * tj, m, SYRK ? m : n, kdim, lane);
    tile_epilogue(acc, Cref.at(b), ldc, 32 * ti, 32 * tj, m, SYRK ? m : n, SYRK, alpha, beta, lane);
  }
}

}  // namespace kblasx
