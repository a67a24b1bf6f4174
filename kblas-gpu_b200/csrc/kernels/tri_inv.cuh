// kernels/tri_inv.cuh -- batched triangular inverse family for n <= 32 (sm_100a): trtri, lauum, potri in ONE launch.
//
// SURVEY.md §8(f)2 ("next" row: the consumers of the Cholesky factor).  The reference builds these from its recursion
// again -- trtri: register kernels up to 16 + two TRSMs per level (Xtrtri_batch_drivers.cuh:31-125), lauum: register
// kernels + TRMM + SYRK (Xlauum_batch_drivers.cuh:31-), potri = trtri then lauum (Xpotri_batch_drivers.cuh:31-),
// poti = potrf then potri (Xpoti_batch_drivers.cuh:82-89).  Here a group of NP lanes owns one matrix; lane j owns COLUMN j
// of the result in registers, the operand sits in the group's shared-memory tile and is read as broadcasts:
//   trtri : X = L^-1.   column j of X is the forward substitution of e_j:  x_k *= 1/l_kk;  x_i -= l_ik x_k  (i > k)
//   lauum : R = L^T L.  r_ij = sum_{k >= i} l_ki l_kj  (i >= j): l_kj from the lane's own column, l_ki broadcast
//   potri : trtri into the tile, then lauum on it: (L L^T)^-1 = X^T X, lower triangle.
// Results go back through the tile so that loads and stores are both row-per-lane (coalesced along columns).
// Only the lower triangle of A is read or written; ragged n is padded with the identity.
#pragma once

#include "common.cuh"

namespace kblasx {

enum TriInvOp { TI_TRTRI = 0, TI_LAUUM = 1, TI_POTRI = 2 };

// NP in {8, 16, 32}: lanes per matrix = padded order; 32 / NP matrices per warp
template <typename T, int NP, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
tri_inv_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount) {
  constexpr int MPW = 32 / NP;
  constexpr int LD = NP + 1;  // odd stride: row-per-lane and column-per-lane accesses both conflict-free
  __shared__ T tile[WARPS * MPW * NP * LD];
  __shared__ T invd_s[WARPS * MPW * NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / NP, j = lane % NP;
  T *Ls = tile + (warp * MPW + g) * NP * LD;  // Ls[i + k*LD] = L(i, k)
  T *invd = invd_s + (warp * MPW + g) * NP;
  const long nwb = ((long)batchCount + MPW - 1) / MPW;
  const long last = (long)batchCount - 1;
  for (long wb = (long)blockIdx.x * WARPS + warp; wb < nwb; wb += (long)gridDim.x * WARPS) {
    const long mat = wb * MPW + g;
    const bool active = mat <= last;
    T *__restrict__ A = Aref.at(active ? mat : last);
    // ---- load the lower triangle, lane = row (identity padding) ------------------------------
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      T v = (j == k) ? T(1) : T(0);
      ldg_stream_if(v, A + j + (long)k * lda, j < n && k < n && j >= k);
      Ls[j + k * LD] = v;
    }
    __syncwarp();
    T x[NP];
    if (OP == TI_TRTRI || OP == TI_POTRI) {
      invd[j] = T(1) / Ls[j + j * LD];
      __syncwarp();
      // column j of L^-1: forward substitution of e_j
#pragma unroll
      for (int i = 0; i < NP; ++i) x[i] = (i == j) ? T(1) : T(0);
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        x[k] *= lds_one(invd + k);
#pragma unroll
        for (int i = k + 1; i < NP; ++i) x[i] = fma_t(-lds_one(Ls + i + k * LD), x[k], x[i]);
      }
      __syncwarp();  // everybody is done reading L
      if (OP == TI_POTRI) {
        // X replaces L in the tile (column j by lane j; zero above the diagonal by construction)
#pragma unroll
        for (int i = 0; i < NP; ++i) Ls[i + j * LD] = x[i];
        __syncwarp();
      }
    }
    if (OP == TI_LAUUM || OP == TI_POTRI) {
      T c[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) c[k] = (OP == TI_POTRI) ? x[k] : ((k >= j) ? lds_one(Ls + k + j * LD) : T(0));
      // r_i = sum_{k >= i} l_ki l_kj for i >= j (rows above the diagonal are never stored)
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        T s = T(0);
#pragma unroll
        for (int k = i; k < NP; ++k) s = fma_t(lds_one(Ls + k + i * LD), c[k], s);
        x[i] = s;
      }
      __syncwarp();
    }
    // ---- result column j -> tile -> global, lane = row ----------------------------------------
#pragma unroll
    for (int i = 0; i < NP; ++i) Ls[i + j * LD] = x[i];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NP; ++k) stg_stream_if(A + j + (long)k * lda, Ls[j + k * LD], active && j < n && k < n && j >= k);
    __syncwarp();
  }
}

}  // namespace kblasx

namespace kblasx {

// ---- LAUUM for n > 32: A := L^T L (lower triangle), in place, one warp per matrix ------------------------------------------
// Block row J of the result, J ascending, needs only block rows I >= J of L:
//   (L^T L)[J][K] = sum_{I >= J} L[I][J]^T L[I][K],  K <= J,
// so writing block row J (columns K < J first, the diagonal block last: every product of the row reads L[J][J]) never
// destroys an operand of a later product.  Each 32 x 32 product is staged through shared memory (coalesced loads; the
// triangular blocks L[J][J] are zero-filled above the diagonal) and accumulated in dot form, lane = result column.
// The reference composes the same result recursively from LAUUM + SYRK + TRMM launches (Xlauum_batch_drivers.cuh:31-);
// this is one launch and needs neither workspace nor the TRMM routine.
template <typename T>
struct LauumBlockedSmem {
  static constexpr int NB = 32, P = 33;
  static constexpr int per_warp = NB * NB + NB * P;  // L[I][J] in memory order, L[I][K] with an odd column pitch
};

template <typename T, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32)
lauum_blocked_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount) {
  constexpr int NB = 32, P = LauumBlockedSmem<T>::P;
  typedef typename Vec2T<T>::type V2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *SJ = reinterpret_cast<T *>(smem_raw) + warp * LauumBlockedSmem<T>::per_warp;
  T *SK = SJ + NB * NB;
  const long mat = (long)blockIdx.x * WARPS + warp;
  if (mat >= (long)batchCount) return;  // warp-uniform
  T *__restrict__ A = Aref.at(mat);
  const int nblk = (n + NB - 1) / NB;
  for (int J = 0; J < nblk; ++J) {
    const int j0 = J * NB, jb = (n - j0 < NB) ? (n - j0) : NB;
    for (int K = 0; K <= J; ++K) {
      const int k0 = K * NB, kb = (n - k0 < NB) ? (n - k0) : NB;
      T C[NB];
#pragma unroll
      for (int r = 0; r < NB; ++r) C[r] = T(0);
      for (int I = J; I < nblk; ++I) {
        const int i0 = I * NB, ib = (n - i0 < NB) ? (n - i0) : NB;
        __syncwarp();  // the previous product is done with the tiles
        {
          T vj[NB], vk[NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            vj[c] = T(0);
            vk[c] = T(0);
            ldg_stream_if(vj[c], A + (i0 + lane) + (long)(j0 + c) * lda, lane < ib && c < jb && !(I == J && lane < c));
            ldg_stream_if(vk[c], A + (i0 + lane) + (long)(k0 + c) * lda, lane < ib && c < kb && !(I == K && lane < c));
          }
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            SJ[c * NB + lane] = vj[c];
            SK[c * P + lane] = vk[c];
          }
        }
        __syncwarp();
        T v[NB];
#pragma unroll
        for (int t = 0; t < NB; ++t) v[t] = SK[lane * P + t];  // column `lane` of L[I][K]
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          T acc[2] = {T(0), T(0)};
#pragma unroll
          for (int t = 0; t < NB; t += 2) {
            const V2 s2 = lds_pair(SJ + r * NB + t);  // L[i0 + t][j0 + r], L[i0 + t + 1][j0 + r]
            acc[0] = fma_t(s2.x, v[t], acc[0]);
            acc[1] = fma_t(s2.y, v[t + 1], acc[1]);
          }
          C[r] += acc[0] + acc[1];
        }
      }
      // column k0 + lane of block (J, K); only the lower part of the diagonal block
#pragma unroll
      for (int r = 0; r < NB; ++r)
        stg_stream_if(A + (j0 + r) + (long)(k0 + lane) * lda, C[r], r < jb && lane < kb && (K < J || r >= lane));
    }
  }
}

}  // namespace kblasx
