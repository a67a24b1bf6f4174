// kernels/trsm_mma.cuh -- side-R triangular solves and the fused POTRS with a k x k lower fp64 factor, k > 32: blocked
// substitution whose off-diagonal part runs on DMMA (mma.sync m8n8k4 f64), one launch per call (sm_100a).
//
// Why: the FMA version (kernels/trsm_blocked.cuh) feeds every two FMAs with one shared-memory load; for config 4 (dposv,
// 16 right-hand-side rows, n = 64 / 128 / 256, batch 64K) its solve half took 1.33 / 4.5 / 16.3 ms where HBM + FP64 allow
// about 0.5 / 1.2 / 5 ms, with the LSU pipe > 60 % busy and 255 registers plus spills
// (profiles/r02_ncu_dposv64_solve_tri_blocked.json).  The blocked substitution is, for every 32-column block J,
//   forward  (X L^T = aB):  X_J = ( a B_J - X[:, 0:j0]      L[j0:j0+32, 0:j0]^T   ) L[J][J]^-T
//   backward (X L   = aB):  X_J = ( a B_J - X[:, j0+32:k]   L[j0+32:k, j0:j0+32]  ) L[J][J]^-1
// and the products are plain (rows x K)(K x 32) contractions: they are done here exactly like the panel update of the
// Cholesky kernel (kernels/potrf_panel_mma.cuh) -- both operands fetched from global/L2 in fragment order (8 rows x 4
// columns = whole sectors per load), 8 or 16 independent accumulator pairs, no shared-memory traffic -- and only the
// 32 x 32 diagonal solve uses the row-per-lane FMA scheme of kernels/trsm_small.cuh.
// A warp owns either two matrices with <= 16 right-hand-side rows each (GP = 16: the 16-row products are two 8-row
// fragment tiles; the diagonal solves of the two matrices share the warp, one half-warp each) or one 32-row slab of one
// matrix (GP = 32).  Replaces the reference's recursion TRSM -> GEMM (cuBLAS batched) -> TRSM
// (Xtrsm_batch_drivers.cuh:127-266) and the 4 x TRSM + 2 x GEMM POTRS (Xpotrs_batch_drivers.cuh:94-171).
#pragma once

#include "common.cuh"
#include "potrf_panel_mma.cuh"  // dmma_m8n8k4
#include "trsm_small.cuh"
#include "trsm_left_vec.cuh"  // tri_vec_substitute

// measured on B200 (dposv pointer array, 16 rhs rows, n = 128 / 256 at batch 64K / 32K, ms): unroll 2 without the L2 prefetch of
// the next panel 8.50 / 19.8, unroll 4 8.69 / 20.3, unroll 2 with the prefetch 8.86 / 20.8 (3500 matrices in flight: the
// prefetched panels evict each other), the FMA kernel 9.34 / 21.1
#ifndef KX_TMMA_UNROLL
#define KX_TMMA_UNROLL 2
#endif

namespace kblasx {

struct TriMmaSmem {
  static constexpr int NB = 32;
  static constexpr int LD = 33;  // odd row stride of the accumulator tile: conflict-free row-per-lane reads
  // per warp: two regions of NB*NB + 2*NB doubles; each holds the accumulator tile of one matrix (NB x LD) and later its
  // diagonal block + reciprocal diagonal (NB*NB + NB) -- never live at the same time
  static constexpr int region = NB * NB + 2 * NB;
  static constexpr int per_warp = 2 * region;
};

// acc_w(8*MT x LD, row-major) = X[xr0 .. xr0+8*MT-1, kbeg:kend] * op(L), one warp.
//   FORWARD:  op(L)[kk][c] = L[j0 + c][kk]   (kbeg = 0, kend = j0: multiples of 32)
//   backward: op(L)[kk][c] = L[kk][j0 + c]   (kbeg = j0 + 32, kend = k: may be ragged)
// Rows of X beyond nrows and rows / columns of L beyond k are redirected to valid addresses (finite data; those results
// are never used), columns kk >= kend contribute zero.
template <int MT, bool FORWARD>
__device__ __forceinline__ void solve_update_mma(double *acc_w, const double *X, const int ldx, const int xr0, const int nrows,
                                                 const double *L, const int lda, const int k, const int j0, const int kbeg,
                                                 const int kend, const int lane) {
  constexpr int LD = TriMmaSmem::LD;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[MT][4][2];
#pragma unroll
  for (int rb = 0; rb < MT; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
  int xrow[MT], lidx[4];
#pragma unroll
  for (int b = 0; b < MT; ++b) {
    xrow[b] = xr0 + 8 * b + fr;
    xrow[b] = xrow[b] < nrows ? xrow[b] : nrows - 1;
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    lidx[b] = j0 + 8 * b + fr;  // row (forward) or column (backward) of L
    lidx[b] = lidx[b] < k ? lidx[b] : k - 1;
  }
  constexpr int UNR = KX_TMMA_UNROLL;
  if (!FORWARD && ((reinterpret_cast<unsigned long long>(L) | (unsigned long long)(lda & 1) << 3) & 15) == 0) {
    // backward, 16-byte aligned columns: the contraction index runs down the columns of L, so a lane fetches TWO consecutive
    // k per column with one 16-byte load (fragment slot fk of the first DMMA = k0 + 2 fk, of the second = k0 + 2 fk + 1; X is
    // fetched with the same permutation).  64-byte instead of 32-byte segments per column: half the load instructions and
    // L2 requests of the plain form below, which made the backward pass 1.5x slower than the forward one.
#pragma unroll UNR
    for (int kk0 = kbeg; kk0 < kend; kk0 += 8) {
      const int ka = kk0 + 2 * fk;  // even
      const bool live0 = ka < kend, live1 = ka + 1 < kend;  // a ragged end: dead slots contribute exact zeros
      const int kc = live0 ? ka : kbeg;
      const int kx1 = live1 ? kc + 1 : kc;  // column kend of X does not exist
      double af0[MT], af1[MT];
      double2 bf[4];
#pragma unroll
      for (int b = 0; b < MT; ++b) {
        af0[b] = X[xrow[b] + (long)kc * ldx];
        af1[b] = X[xrow[b] + (long)kx1 * ldx];
        af0[b] = live0 ? af0[b] : 0.0;
        af1[b] = live1 ? af1[b] : 0.0;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        // row kc + 1 <= k of a column < k - 1: inside the matrix' storage even when it is past the last row
        bf[b] = *reinterpret_cast<const double2 *>(L + kc + (long)lidx[b] * lda);
        bf[b].x = live0 ? bf[b].x : 0.0;
        bf[b].y = live1 ? bf[b].y : 0.0;
      }
#pragma unroll
      for (int rb = 0; rb < MT; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af0[rb], bf[cb].x);
#pragma unroll
      for (int rb = 0; rb < MT; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af1[rb], bf[cb].y);
    }
  } else {
#pragma unroll UNR
    for (int kk0 = kbeg; kk0 < kend; kk0 += 4) {
      const int kk = kk0 + fk;
      const bool live = FORWARD || kk < kend;
      const int kc = live ? kk : kend - 1;
      double af[MT], bf[4];
#pragma unroll
      for (int b = 0; b < MT; ++b) {
        af[b] = X[xrow[b] + (long)kc * ldx];
        if (!FORWARD) af[b] = live ? af[b] : 0.0;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = FORWARD ? L[lidx[b] + (long)kc * lda] : L[kc + (long)lidx[b] * lda];
#pragma unroll
      for (int rb = 0; rb < MT; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[rb][cb][0], acc[rb][cb][1], af[rb], bf[cb]);
    }
  }
#pragma unroll
  for (int rb = 0; rb < MT; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk] = acc[rb][cb][0];
      acc_w[(8 * rb + fr) * LD + 8 * cb + 2 * fk + 1] = acc[rb][cb][1];
    }
}

// one pass (forward or backward) over all 32-column blocks.  GP = 16: matrices Aq[0], Aq[1] / Bq[0], Bq[1], lane group g owns
// rows lane % 16 of matrix g (MPW = 1: only matrix 0, lanes 16..31 hold no rows -- the fused POSV kernel); GP = 32: one
// matrix (Aq[0], Bq[0]), lane = row xr0 + lane.
template <bool FORWARD, int GP, int MPW = 32 / GP>
__device__ __forceinline__ void tri_mma_pass(const int k, const double alpha, const double *const (&Aq)[2], const int lda,
                                             double *const (&Bq)[2], const int ldb, const int xr0, const int nrows,
                                             const bool have, double *smem_w, const int lane) {
  constexpr int NB = 32, LD = TriMmaSmem::LD, RG = TriMmaSmem::region;
  const int g = (GP == 16 && MPW == 2) ? (lane >> 4) : 0;
  const int r = GP == 16 ? (lane & 15) : lane;  // my row inside the region of my matrix (MPW = 1, GP = 16: lanes 16..31 idle along)
  double *B = Bq[g];
  const int my = xr0 + r;
  const int nblk = (k + NB - 1) / NB;
  for (int bi = 0; bi < nblk; ++bi) {
    const int J = FORWARD ? bi : (nblk - 1 - bi);
    const int j0 = J * NB;
    const int jb = (k - j0 < NB) ? (k - j0) : NB;
    // my rows of block J are pulled into L2 now and loaded after the product (64 registers that the accumulators need)
    if (lane < jb) {
#pragma unroll
      for (int q = 0; q < MPW; ++q) {
        const double *p0 = Bq[q] + xr0 + (long)(j0 + lane) * ldb;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p0));
        if (GP == 32 && xr0 + 16 < nrows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + 16));
      }
    }
    if (bi > 0) {
      const int kbeg = FORWARD ? 0 : j0 + NB, kend = FORWARD ? j0 : k;
#pragma unroll
      for (int q = 0; q < MPW; ++q)
        solve_update_mma<GP == 16 ? 2 : 4, FORWARD>(smem_w + q * RG, Bq[q], ldb, xr0, nrows, Aq[q], lda, k, j0, kbeg, kend, lane);
      __syncwarp();
    }
    double x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      x[c] = 0.0;
      ldg_stream_if(x[c], B + my + (long)(j0 + c) * ldb, have && c < jb);
    }
    if (bi > 0) {
      const double *acc_w = smem_w + g * RG;
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = fma(alpha, x[c], -acc_w[r * LD + c]);
      __syncwarp();  // the accumulator tiles are dead: their regions take the diagonal blocks
    } else {
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] *= alpha;
    }
    // diagonal block(s): global -> shared with cp.async (no staging registers next to the 64 that hold x), zero-filled
    // outside jb x jb; the reciprocal diagonal is 1 there, so the padded entries of x stay zero
#pragma unroll
    for (int q = 0; q < MPW; ++q) {
      const double *D = Aq[q] + j0 + (long)j0 * lda;
      double *Ls = smem_w + q * RG;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const bool in = lane < jb && c < jb && lane >= c;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(Ls + lane + c * NB)),
                     "l"(in ? D + lane + (long)c * lda : D), "r"(in ? 8 : 0) : "memory");
      }
    }
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int q = 0; q < MPW; ++q) {
      double *Ls = smem_w + q * RG;
      Ls[NB * NB + lane] = lane < jb ? 1.0 / Ls[lane + lane * NB] : 1.0;
    }
    __syncwarp();
    const double *Lkk = smem_w + g * RG, *invd = Lkk + NB * NB;
    if (FORWARD) tri_vec_substitute<double, NB, TRI_FORWARD>(x, Lkk, invd);
    else tri_vec_substitute<double, NB, TRI_BACKWARD>(x, Lkk, invd);
#pragma unroll
    for (int c = 0; c < NB; ++c) stg_stream_if(B + my + (long)(j0 + c) * ldb, x[c], have && c < jb);
    // the rows just written are operands of the next block's product, fetched by other lanes of this warp
    __threadfence_block();
    __syncwarp();
  }
}

#ifndef KX_TRI_MMA_MINB
#define KX_TRI_MMA_MINB 3
#endif
template <int OP, int GP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, KX_TRI_MMA_MINB)
tri_solve_mma_kernel(const int k, const int vec, const double alpha, BatchRef<const double, STRIDED> Aref, const int lda,
                     BatchRef<double, STRIDED> Bref, const int ldb, const int batchCount, const int slabs) {
  constexpr int MPW = 32 / GP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double *smem_w = reinterpret_cast<double *>(smem_raw) + warp * TriMmaSmem::per_warp;

  const long task = (long)blockIdx.x * WARPS + warp;  // (warp-batch of MPW matrices, 32-row slab)
  const long wbatches = ((long)batchCount + MPW - 1) / MPW;
  if (task >= wbatches * slabs) return;  // warp-uniform
  const long mat0 = (task / slabs) * MPW;
  const long last = (long)batchCount - 1;
  const int xr0 = GP == 16 ? 0 : (int)(task % slabs) * 32;
  const int g = GP == 16 ? (lane >> 4) : 0;
  const double *Aq[2];
  double *Bq[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const long m = (q < MPW && mat0 + q < last) ? mat0 + q : (q < MPW ? last : mat0);
    Aq[q] = Aref.at(m);
    Bq[q] = Bref.at(m);
  }
  const int r = GP == 16 ? (lane & 15) : lane;
  const bool have = (mat0 + g <= last) && (xr0 + r < vec);

  if (OP == TRI_FORWARD || OP == TRI_BOTH)
    tri_mma_pass<true, GP>(k, alpha, Aq, lda, Bq, ldb, xr0, vec, have, smem_w, lane);
  if (OP == TRI_BACKWARD || OP == TRI_BOTH)
    tri_mma_pass<false, GP>(k, OP == TRI_BOTH ? 1.0 : alpha, Aq, lda, Bq, ldb, xr0, vec, have, smem_w, lane);
}

}  // namespace kblasx
