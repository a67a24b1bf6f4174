// kernels/potrf_panel_mma.cuh -- batched Cholesky for n > 32 with the left-looking trailing update on
// the tensor path that mma.sync still reaches on sm_100a, one CTA (by default one WARP) per matrix:
//   fp64: DMMA  (mma.sync m8n8k4 f64);
//   fp32: 3 x TF32 (mma.sync m16n8k8 tf32, fp32 accumulate) on operands split x = hi + lo with
//         hi = tf32(x), lo = tf32(x - hi):  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi.  The dropped
//         a_lo*b_lo term and the rounding of lo are O(2^-22) relative -- fp32-grade results (parity
//         tests: same tolerances as the FFMA kernel) at 1/3 of the TF32 tensor rate, which is still
//         several times the FFMA pipe and needs no shared-memory broadcast traffic.
//
// For the panel of columns j0 .. j0+31 every warp owns a 32-row x 32-column block of the panel.
//   1. update: acc(32x32) = L[rows, 0:j0] * L[j0:j0+32, 0:j0]^T, operands fetched straight from
//      global/L2 in fragment order: both the A fragment (rows of this warp) and the B fragment
//      (rows j0.. of the panel) are "element (base + lane/4, k + lane%4)" of the same column-major
//      matrix -- 8 consecutive rows x 4 columns per load instruction, whole 32/64-byte sectors.
//      16 (fp64) / 8 x 4 (fp32) independent accumulator registers hide the MMA latency.
//      No shared-memory broadcast traffic at all: the FMA version needs 16 LDS per 64 FMAs and is
//      MIO-bound.
//   2. the accumulators go through (padded) shared memory into a row-per-thread layout,
//      p = A[row, j0:j0+32] - acc;
//   3. warp 0 factors the diagonal block (row per lane), 4. the other rows solve against it,
//   5. rows are stored.
// Measured on B200: DMMA peaks at 63.5 FMA/clk/SM, the same as the DFMA pipe
// (profiles/r01_microbench_pipes.txt); it wins by freeing issue slots and the MIO pipe.
#pragma once

#include <cstdint>

#include "common.cuh"
#include "trsm_small.cuh"

namespace kblasx {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// D(16x8) += A(16x8, row) * B(8x8, col), TF32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                                 unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x = hi + lo, both representable in TF32 (round to nearest)
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);  // exact
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

template <typename T, int THREADS>
struct PanelMmaSmem {
  static constexpr int NB = 32;
  static constexpr int LD = 33;  // odd row stride: conflict-free row-per-thread reads
  static constexpr int warps = THREADS / 32;
  static constexpr size_t bytes = sizeof(T) * (NB * NB + NB + (size_t)warps * NB * LD);
};

// acc_w(32 x LD, row-major) = L[wrow0 .. wrow0+31, 0:j0] * L[j0 .. j0+31, 0:j0]^T   (one warp)
__device__ __forceinline__ void panel_update_mma(double *acc_w, const int LD, const double *__restrict__ A, const int lda,
                                                 const int n, const int wrow0, const int j0, const int lane) {
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane
  double acc[4][4][2];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
#ifndef KX_PANEL_NO_PAIRED_LOADS
  if (((reinterpret_cast<unsigned long long>(A) | ((unsigned long long)(lda & 1) << 3)) & 15) == 0 && ((wrow0 | j0) & 1) == 0) {
    // 16-byte aligned columns: a lane fetches TWO consecutive rows with one 16-byte load and feeds them to two different
    // fragment tiles -- tile (p, h) holds rows 16 p + 2 fr + h instead of 8 (2 p + h) + fr.  Any bijection between fragment
    // slots and rows is a valid m8n8k4 operand as long as the accumulator is written back with the same one; this one makes
    // every load instruction cover 128 contiguous bytes per column (8 lanes x 16 B) instead of 64, with half as many loads.
    int ar[2], br[2];
    // pairs that start beyond the last row read the last pair instead; a pair may END one row past it when n is odd: that
    // element is the head of the next column (k < j0 <= n - 1, so it is inside the matrix) and only reaches accumulator
    // rows / columns >= n, which nobody uses
    const int last2 = (n - 1) & ~1;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      ar[p] = wrow0 + 16 * p + 2 * fr;
      ar[p] = ar[p] < n ? ar[p] : last2;
      br[p] = j0 + 16 * p + 2 * fr;
      br[p] = br[p] < n ? br[p] : last2;
    }
    const double *col = A + (long)fk * lda;
#pragma unroll 4
    for (int k = 0; k < j0; k += 4) {
      double2 a2[2], b2[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        a2[p] = *reinterpret_cast<const double2 *>(col + ar[p]);
        b2[p] = *reinterpret_cast<const double2 *>(col + br[p]);
      }
      const double af[4] = {a2[0].x, a2[0].y, a2[1].x, a2[1].y}, bf[4] = {b2[0].x, b2[0].y, b2[1].x, b2[1].y};
#pragma unroll
      for (int rb = 0; rb < 4; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af[rb], bf[cb]);
      col += 4 * (long)lda;
    }
    // C fragment of tile (rb = 2 p + h, cb = 2 q + g): row slot fr -> row 16 p + 2 fr + h, column slots 2 fk, 2 fk + 1 ->
    // columns 16 q + 4 fk + g and 16 q + 4 fk + 2 + g
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int row = 16 * (rb >> 1) + 2 * fr + (rb & 1), c0 = 16 * (cb >> 1) + 4 * fk + (cb & 1);
        acc_w[row * LD + c0] = acc[rb][cb][0];
        acc_w[row * LD + c0 + 2] = acc[rb][cb][1];
      }
    return;
  }
#endif
  // rows beyond n read row n-1 instead (finite data, results discarded)
  int arow[4], brow[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    arow[b] = wrow0 + 8 * b + fr;
    arow[b] = arow[b] < n ? arow[b] : n - 1;
    brow[b] = j0 + 8 * b + fr;
    brow[b] = brow[b] < n ? brow[b] : n - 1;
  }
  const double *col = A + (long)fk * lda;
#pragma unroll 4
  for (int k = 0; k < j0; k += 4) {
    double af[4], bf[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      af[b] = col[arow[b]];
      bf[b] = col[brow[b]];
    }
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af[rb], bf[cb]);
    col += 4 * (long)lda;
  }
  // C fragment: lane holds (row fr, cols 2*fk, 2*fk+1) of each 8x8 tile
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk] = acc[rb][cb][0];
      acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk + 1] = acc[rb][cb][1];
    }
}

__device__ __forceinline__ void panel_update_mma(float *acc_w, const int LD, const float *__restrict__ A, const int lda,
                                                 const int n, const int wrow0, const int j0, const int lane) {
  const int fr = lane >> 2, fk = lane & 3;
  float acc[2][4][4];  // [16-row tile][8-column tile][C fragment]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  int arow[4], brow[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    arow[b] = wrow0 + 8 * b + fr;
    arow[b] = arow[b] < n ? arow[b] : n - 1;
    brow[b] = j0 + 8 * b + fr;
    brow[b] = brow[b] < n ? brow[b] : n - 1;
  }
  // A fragment of tile mt: (row 16mt + fr [+8], k + fk [+4]);  B fragment of tile nt: (k + fk [+4], row j0 + 8nt + fr)
  const float *col0 = A + (long)fk * lda;
  const float *col1 = A + (long)(fk + 4) * lda;
#pragma unroll 2
  for (int k = 0; k < j0; k += 8) {
    float af[4][2], bf[4][2];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      af[b][0] = col0[arow[b]];
      af[b][1] = col1[arow[b]];
      bf[b][0] = col0[brow[b]];
      bf[b][1] = col1[brow[b]];
    }
    unsigned ah[4][2], al[4][2], bh[4][2], bl[4][2];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        split_tf32(af[b][h], ah[b][h], al[b][h]);
        split_tf32(bf[b][h], bh[b][h], bl[b][h]);
      }
    // small terms first
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        mma_tf32_m16n8k8(acc[mt][nt], al[2 * mt][0], al[2 * mt + 1][0], al[2 * mt][1], al[2 * mt + 1][1], bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        mma_tf32_m16n8k8(acc[mt][nt], ah[2 * mt][0], ah[2 * mt + 1][0], ah[2 * mt][1], ah[2 * mt + 1][1], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        mma_tf32_m16n8k8(acc[mt][nt], ah[2 * mt][0], ah[2 * mt + 1][0], ah[2 * mt][1], ah[2 * mt + 1][1], bh[nt][0], bh[nt][1]);
    col0 += 8 * (long)lda;
    col1 += 8 * (long)lda;
  }
  // C fragment: (row fr, cols 2fk, 2fk+1), (row fr+8, same cols) of each 16x8 tile
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float *q = acc_w + (16 * mt + fr) * LD + 8 * nt + 2 * fk;
      q[0] = acc[mt][nt][0];
      q[1] = acc[mt][nt][1];
      q[8 * LD] = acc[mt][nt][2];
      q[8 * LD + 1] = acc[mt][nt][3];
    }
}

// resident warps per SM the register allocation is sized for: fp64 keeps 16 x 2 accumulators + a
// prefetch window live and needs the full 255 registers (8 warps); measured n = 64 / 128 / 256:
// 16 warps (128 regs, spills) 4.8 / 8.5 / 10.8 TFLOP/s, 12 warps 5.4 / 9.6 / 12.4, 8 warps 5.5 / 9.5 / 14.6
#ifndef KX_PANEL_WARPS_PER_SM
#define KX_PANEL_WARPS_PER_SM 8
#endif
#ifndef KX_PANEL_WARPS_PER_SM_F32
#define KX_PANEL_WARPS_PER_SM_F32 16
#endif
template <typename T>
struct PanelMmaOcc {
  static constexpr int warps_per_sm = sizeof(T) == 8 ? KX_PANEL_WARPS_PER_SM : KX_PANEL_WARPS_PER_SM_F32;
};

// the whole factorisation of ONE matrix by the calling CTA (THREADS threads); returns the LAPACK-style info value
// (0, or 1 + the index of the first non-positive pivot when info_mode is set).  Also the first half of the fused POSV kernel
// (kernels/posv_fused.cuh).
template <typename T, int THREADS>
__device__ __forceinline__ int potrf_panel_mma_body(const int n, T *__restrict__ A, const int lda, const int info_mode,
                                                    unsigned char *smem_raw) {
  constexpr int NB = 32;
  constexpr int LD = PanelMmaSmem<T, THREADS>::LD;
  T *Lkk = reinterpret_cast<T *>(smem_raw);  // factored diagonal block, column-major, identity padded
  T *invd = Lkk + NB * NB;                    // 1 / diag(L_JJ)
  T *accs = invd + NB;                        // per warp: 32 x LD transpose buffer

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  T *acc_w = accs + warp * NB * LD;
  int bad = 0;

  for (int j0 = 0; j0 < n; j0 += NB) {
    const int jb = (n - j0 < NB) ? (n - j0) : NB;
    const int m = n - j0;  // rows of this panel
    for (int r0 = 0; r0 < m; r0 += THREADS) {
      const int wrow0 = j0 + r0 + warp * 32;  // first row of this warp's block
      const bool warp_has_rows = wrow0 < n;   // warp-uniform
      const int row = wrow0 + lane;
      const bool valid = row < n;
      T p[NB];

      // ---- 1. acc = L[rows, 0:j0] * L[j0:j0+32, 0:j0]^T on the tensor path --------------------
      // (the rows of the NEXT slab are pulled into L2 meanwhile: with ~1200 matrices in flight per GPU the
      //  factored columns do not survive in L2 between panels -- ncu: 30 % L2 hit rate, 40 % of the stalls on
      //  the scoreboard -- so without the hint every slab starts with a DRAM round trip)
#ifndef KX_PANEL_NO_PREFETCH
      if (j0 > 0) {
        const int nrow0 = wrow0 + THREADS;  // first row of this warp's next slab
        if (nrow0 < n) {
          constexpr int LPC = (32 * (int)sizeof(T)) / 128 > 0 ? (32 * (int)sizeof(T)) / 128 : 1;  // 128-byte lines per 32-row column segment
          for (int l = lane; l < LPC * j0; l += 32) {
            const int c = l / LPC, hh = l % LPC;
            int r = nrow0 + hh * (128 / (int)sizeof(T));
            r = r < n ? r : n - 1;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A + r + (long)c * lda));
          }
        }
      }
#endif
      if (warp_has_rows && j0 > 0) {
        panel_update_mma(acc_w, LD, A, lda, n, wrow0, j0, lane);
      }
      __syncwarp();

      // ---- 2. my row of the panel: p = A[row, j0 : j0+32] - acc -------------------------------
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        p[c] = T(0);
        ldg_stream_if(p[c], A + row + (long)(j0 + c) * lda, valid && c < jb);
      }
      if (j0 > 0) {
#pragma unroll
        for (int c = 0; c < NB; ++c) p[c] -= acc_w[lane * LD + c];
      }
      __syncwarp();

      // ---- 3. diagonal block: rows j0 .. j0+31 are warp 0's rows in the first slab -------------
      if (r0 == 0) {
        if (warp == 0) {
          if (lane >= jb) {  // identity padding of a ragged last panel
#pragma unroll
            for (int c = 0; c < NB; ++c) p[c] = (c == lane) ? T(1) : T(0);
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const T d = shfl_seg<32>(p[j], j);
            if (info_mode && bad == 0 && j < jb && !(d > T(0))) bad = j0 + j + 1;
            const T r = rsqrt_t(d);
            p[j] *= r;
            Lkk[lane + j * NB] = p[j];
            if (lane == j) invd[j] = r;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NB; ++k) p[k] = fma_t(-p[j], lds_one(Lkk + k + j * NB), p[k]);
          }
        }
        __syncthreads();  // L_JJ and invd are published
      }

      // ---- 4. forward substitution of the rows below the diagonal block -----------------------
      if (!(r0 == 0 && warp == 0)) tri_forward<T, NB>(p, Lkk, invd);

      // ---- 5. store --------------------------------------------------------------------------
#pragma unroll
      for (int c = 0; c < NB; ++c)
        stg_stream_if(A + row + (long)(j0 + c) * lda, p[c], valid && c < jb && row >= j0 + c);
    }
    // the factored panel must be visible to the whole CTA before panel J+1 reads it from global
    __threadfence_block();
    __syncthreads();
  }
  return bad;
}

template <typename T, int THREADS, bool STRIDED>
__global__ void __launch_bounds__(THREADS, (32 * PanelMmaOcc<T>::warps_per_sm) / THREADS)
potrf_panel_mma_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount, int *__restrict__ info,
                       const int info_mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int bad = potrf_panel_mma_body<T, THREADS>(n, Aref.at(blockIdx.x), lda, info_mode, smem_raw);
  if (info_mode && threadIdx.x == 0) info[blockIdx.x] = bad;
}

}  // namespace kblasx
