// kernels/potrf_small.cuh -- register-resident batched Cholesky for n <= 32 (sm_100a).
//
// Replaces the reference's 1-4 launches per call (K1-K4 potrf kernels + K5 trsm + K13 syrk,
// Xpotrf_batch_kernels.cuh:37-488, driver recursion Xpotrf_batch_drivers.cuh:43-137) by ONE
// launch in which every matrix is read once and written once.
//
// Mapping.  A group of G lanes owns one matrix, 32/G matrices per warp.  Lane l of the group
// holds rows  G*s + l  (s = 0 .. NP/G-1, "slot") of the lower triangle: slot s keeps columns
// 0 .. G*(s+1)-1, so the per-lane register footprint is G*S(S+1)/2 values (80 for n=32, G=8)
// and the cyclic-by-slot row layout lets slot s drop out of the trailing update as soon as the
// factorisation has passed its rows -- 230 FMA warp-instructions per n=32 matrix instead of
// the 496 of a row-per-lane layout.
//
// Right-looking column step j (same recurrence as Xpotrf_batch_kernels.cuh:50-69):
//   d = A[j][j] (width-G shuffle), r = rsqrt(d), column j *= r, publish column j through a
//   double-buffered shared-memory line, every lane reads L[k][j] (k > j) back as broadcast
//   128-bit loads and applies  A[i][k] -= L[i][j] * L[k][j]  to the rows it owns.
// Shared memory replaces the reference's per-element __shfl broadcast (2 SHFL per fp64 value,
// include/kblas_operators.h:88-97): 1 LDS.128 delivers two values to four matrices at once.
//
// Memory.  Column-major matrix b at A.at(b); a lane's G-row segment of one column is
// contiguous (64 B for fp64, G=8), so every LDG/STG of a warp covers whole 32-byte sectors.
// Sectors that lie entirely in the strict upper triangle are never loaded, and only
// elements with row >= col are stored: the strict upper triangle is bit-preserved, as in
// the reference (stores guarded by tx >= i, Xpotrf_batch_kernels.cuh:73-77).
// Persistent CTAs walk the batch; while one warp-batch is being factored the next one is pulled
// into L2 (prefetch.global.L2), and finished block columns are stored as soon as they are final.
// The whole column loop is unrolled (register indexing), ~45 KB of code: the warps of a CTA are
// re-aligned once per warp-batch (LOCKSTEP) so that they share instruction-cache lines.
// n < NP is handled by padding with the identity in registers (EXACT = false).
#pragma once

#include <cstdint>
#include "common.cuh"

namespace kblasx {

// 16-byte shared-memory broadcast vector: 2 doubles / 4 floats
template <typename T> struct BcastVec;
template <> struct BcastVec<double> {
  typedef double2 type;
  static __device__ __forceinline__ double get(const double2 &v, int h) { return h ? v.y : v.x; }
};
template <> struct BcastVec<float> {
  typedef float4 type;
  static __device__ __forceinline__ float get(const float4 &v, int h) { return h == 0 ? v.x : h == 1 ? v.y : h == 2 ? v.z : v.w; }
};

// L2 prefetch of the lower triangle of one matrix: lane = column, the 128-byte lines spanning rows
// [col, n) of that column (DRAM reads are line granular on B200, profiles/r01_dram_granularity.md).
template <typename T>
__device__ __forceinline__ void prefetch_lower_l2(const T *A, int n, int lda, int lane) {
  if (lane < n) {
    const char *p0 = reinterpret_cast<const char *>(A + lane + (long)lane * lda);
    const char *p1 = reinterpret_cast<const char *>(A + (n - 1) + (long)lane * lda);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p0));
    if ((reinterpret_cast<uintptr_t>(p1) >> 7) != (reinterpret_cast<uintptr_t>(p0) >> 7))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p1));
    if (p1 - p0 > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + 128));
  }
}

// EXACT: n == NP (no bounds predicates, no info support); LOCKSTEP: __syncthreads per warp-batch.
// With EXACT the stores are SECTOR granular: a 32-byte sector that straddles the diagonal is written
// back whole, its strict-upper elements carrying the very bits that were loaded (the register
// updates are predicated so they never touch them).  Partially written sectors cost an L2
// read-modify-write fill each; measured on B200 (tools/microbench_pattern.cu) the pure access
// pattern runs at 3.19 ms / 2^20 matrices with element-exact stores and 2.23 ms with whole sectors.
template <typename T, int NP, int G, int WARPS, int MINB, bool STRIDED, bool EXACT, bool LOCKSTEP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
potrf_reg_kernel(const int n_arg, BatchRef<T, STRIDED> Aref, const int lda, const int batchCount,
                 int *__restrict__ info, const int info_mode_arg, const unsigned stagger_ns) {
  constexpr int S = NP / G;                    // row slots per lane
  constexpr int MPW = 32 / G;                  // matrices per warp
  constexpr int GH = (16 / G) > 0 ? (16 / G) : 1;  // lane groups per half-warp
  constexpr int SE = SectorElems<T>::value;
  // broadcast line of one buffer: [group / GH][row vector][group % GH][VB]  -- the 16 lanes of a half-warp write
  // contiguous bytes (conflict-free stores), a group reads 16-byte vectors: VB = 2 doubles / 4 floats per LDS.128
  // (fp32 with 8-byte pairs kept the shared-memory pipe 90 % busy on n = 32, profiles/r02_ncu_spotrf32_potrf_reg.json)
  constexpr int VB = 16 / (int)sizeof(T);
  constexpr int PAIR = GH * VB;
  constexpr int BUF_STRIDE = (NP / VB) * MPW * VB;  // elements per buffer
  static_assert(NP % VB == 0, "whole broadcast vectors");
  static_assert(NP % G == 0 && G % 2 == 0 && 32 % G == 0, "bad tiling");

  __shared__ __align__(16) T bc[WARPS * 2 * BUF_STRIDE];
  // EXACT: original bits of the strict-upper elements that share a sector with the diagonal,
  // NU per diagonal block and lane (element (row, c) with row < c <= row | (SE-1)), kept in shared
  // memory from load to store so that the register updates need no predication
  constexpr int NU = ((SE < G) ? SE : G) - 1;
  __shared__ T sv[EXACT ? WARPS * S * NU * 32 : 1];

  const int n = EXACT ? NP : n_arg;
  const int info_mode = EXACT ? 0 : info_mode_arg;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int l = lane % G;
  const int g = lane / G;
  T *const wbase = bc + warp * (2 * BUF_STRIDE) + (g / GH) * ((NP / VB) * PAIR) + (g % GH) * VB;

  // persistent CTAs: CTA-batch cb covers warp-batches [cb*WARPS, cb*WARPS + WARPS), a warp-batch
  // covers matrices [wb*MPW, wb*MPW + MPW)
  const long nwb = ((long)batchCount + MPW - 1) / MPW;
  const long ncb = (nwb + WARPS - 1) / WARPS;
  // MINB >= 2 lockstep CTAs per SM: start the second half of the grid half a warp-batch late so that
  // one CTA's load burst overlaps the other's arithmetic (CTAs b and b + #SM share an SM)
  if (LOCKSTEP && MINB >= 2 && stagger_ns > 0 && blockIdx.x >= gridDim.x / 2) __nanosleep(stagger_ns);
  for (long cb = blockIdx.x; cb < ncb; cb += gridDim.x) {
    if (LOCKSTEP) __syncthreads();
    const long wb = cb * WARPS + warp;
    const long mat = wb * MPW + g;
    const bool active = mat < (long)batchCount;
    // inactive groups factor a copy of the last matrix (never stored): keeps the loads unpredicated
    T *__restrict__ A = Aref.at(active ? mat : (long)batchCount - 1);

#define KX_IDX(s_, c_) (G * (((s_) * ((s_) + 1)) / 2) + (c_))
    T a[G * (S * (S + 1)) / 2];

    // ---- load the lower triangle (whole sectors only) -------------------------------------
    {
      const T *pc = A + l;
#pragma unroll
      for (int col = 0; col < NP; ++col) {
#pragma unroll
        for (int s = col / G; s < S; ++s) {
          const int row = G * s + l;
          const bool inside = EXACT || ((row < n) && (col < n));
          const bool need = ((row | (SE - 1)) >= col);  // sector holds at least one lower element
          T v = (inside || row != col) ? T(0) : T(1);   // identity padding; 0 for skipped sectors
          if (EXACT && s > col / G) v = ldg_stream(pc + G * s);      // always needed: plain load
          else ldg_stream_if(v, pc + G * s, inside && need);         // predicated, branch-free
          a[KX_IDX(s, col)] = v;
        }
        pc += lda;
      }
    }

    if (EXACT) {
      // column of slot i for this lane: (l & ~(SE-1)) + 1 + i inside the diagonal block
#pragma unroll
      for (int t = 0; t < S; ++t)
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          T v = a[KX_IDX(t, G * t + 1 + i)];
#pragma unroll
          for (int h = SE; h < G; h += SE)
            if (h + 1 + i < G) v = ((l & ~(SE - 1)) == h) ? a[KX_IDX(t, G * t + h + 1 + i)] : v;
          sv[((warp * S + t) * NU + i) * 32 + lane] = v;
        }
    }

    // ---- pull the NEXT warp-batch of this warp into L2 while this one is being factored ----
    {
      const long nb = wb + (long)gridDim.x * WARPS;
#pragma unroll
      for (int q = 0; q < MPW; ++q) {
        const long m2 = nb * MPW + q;
        if (m2 < (long)batchCount) prefetch_lower_l2<T>(Aref.at(m2), n, lda, lane);
      }
    }

    int bad = 0;
    T *pst = A + l;  // column pointer of the next block of columns to store

    // ---- right-looking factorisation -------------------------------------------------------
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int t = j / G, c = j % G;
      const T d = shfl_seg<G>(a[KX_IDX(t, j)], c);
      if (!EXACT && info_mode) {
        if (bad == 0 && j < n && !(d > T(0))) bad = j + 1;
      }
      const T r = rsqrt_t(d);
#pragma unroll
      for (int s = t; s < S; ++s) a[KX_IDX(s, j)] *= r;

      if (j + 1 < NP) {
        T *wbuf = wbase + (j & 1) * BUF_STRIDE;
        // publish column j (rows of every slot that still reaches below the diagonal)
#pragma unroll
        for (int s = t; s < S; ++s) {
          if (G * s + G - 1 > j) {
            const int k = G * s + l;
            wbuf[(k / VB) * PAIR + (k % VB)] = a[KX_IDX(s, j)];
          }
        }
        __syncwarp();
        // trailing update: A[i][k] -= L[i][j] * L[k][j] for every owned row i >= k > j.
        // Rows above the diagonal inside the diagonal slot pick up garbage that is never published
        // or read (like the reference's unguarded register updates); the EXACT kernel stores the
        // saved original bits for them, the generic one does not store them at all.
#pragma unroll
        for (int p = (j + 1) / VB; p < NP / VB; ++p) {
          const typename BcastVec<T>::type vv = *reinterpret_cast<const typename BcastVec<T>::type *>(wbuf + p * PAIR);
#pragma unroll
          for (int h = 0; h < VB; ++h) {
            const int k = VB * p + h;
            if (k > j) {
              const T v = BcastVec<T>::get(vv, h);
#pragma unroll
              for (int s = k / G; s < S; ++s) a[KX_IDX(s, k)] = fma_t(-a[KX_IDX(s, j)], v, a[KX_IDX(s, k)]);
            }
          }
        }
      }

      // ---- store a finished block of G columns as soon as it is final (spreads the stores) ----
      if (c == G - 1) {
#pragma unroll
        for (int col = j - (G - 1); col <= j; ++col) {
#pragma unroll
          for (int s = col / G; s < S; ++s) {
            const int row = G * s + l;
            const bool keep = EXACT ? ((row | (SE - 1)) >= col) : (row >= col && row < n);
            T val = a[KX_IDX(s, col)];
            if (EXACT && s == col / G && (col % G) % SE != 0)  // sector straddles the diagonal
              val = (l >= col % G) ? val : sv[((warp * S + s) * NU + ((col % G) % SE) - 1) * 32 + lane];
            stg_stream_if(pst + G * s, val, active && keep);
          }
          pst += lda;
        }
      }
    }
    if (!EXACT && info_mode && active && l == 0) info[mat] = bad;
    __syncwarp();  // the broadcast buffers are reused by the next warp-batch
#undef KX_IDX
  }
}

}  // namespace kblasx
