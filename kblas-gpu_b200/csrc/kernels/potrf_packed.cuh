// kernels/potrf_packed.cuh -- batched Cholesky on PACKED lower-triangular storage, n <= 32 (sm_100a).
//
// SURVEY §8(f)4 / VERDICT round 1 item 4: a batch layout whose physical bytes equal the algorithmic bytes.  The
// drop-in column-major layout of kblas_potrf_batch forces whole 128-byte DRAM lines to be fetched for a triangle
// (48 lines read + 48 written per 32 x 32 fp64 matrix = 12288 B of DRAM time for 8448 algorithmic bytes,
// profiles/r01_dram_granularity.md).  Here matrix b is stored the LAPACK way ("packed", uplo = 'L', as ?pptrf takes it):
//     AP_b[ j*n - j(j-1)/2 + (i - j) ] = A_b(i, j),  i >= j          n(n+1)/2 elements, stride >= n(n+1)/2
// so a matrix is ONE contiguous run of bytes, read once and written once.  The reference has no packed routine
// (its batch_pstrf, include/batch_pstrf.h, src/batch_svd/batch_pstrf.cu:226-246, is pivoted Cholesky on full
// storage); the arithmetic is the one of kernels/potrf_small.cuh (same mapping, same column loop, bit-identical
// factors for n % 8 == 0), which restates Xpotrf_batch_kernels.cuh:50-69.
//
// Data movement (n % 8 == 0, 16-byte aligned matrices):
//   in : one TMA bulk copy per matrix (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP), issued by
//        lanes 0..3 of the warp that owns the four matrices.  The staging buffer is re-armed for the warp's NEXT four
//        matrices as soon as the current ones sit in registers, so the load of batch i+1 flies during the whole
//        factorisation of batch i (no register staging, no L2 prefetch hints, nothing on the LSU pipe).
//   out: finished 8-column blocks (contiguous in packed storage) go registers -> shared -> global with bulk stores
//        (cp.async.bulk shared -> global, bulk_group completion), or straight from registers (OUT_BULK = false).
// Shared-memory strides are padded so that the four lane groups of a warp hit disjoint banks.
// Ragged n, unaligned pointer-array entries and LAPACK-info mode take the generic instantiation (plain loads / stores).
#pragma once

#include <cstdint>
#include "common.cuh"
#include "tma.cuh"
#include "potrf_small.cuh"

namespace kblasx {

// ---- packed geometry ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int packed_col_off(int n, int c) { return c * n - (c * (c - 1)) / 2; }  // offset of (c, c)
__host__ __device__ constexpr int packed_size(int n) { return (n * (n + 1)) / 2; }

template <typename T, int NP>
struct PackedGeom {
  static constexpr int G = 8;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int SZ = packed_size(NP);  // elements per matrix
  // per-matrix stride of a staging buffer: == G*ES (mod 128 bytes), so that the lane groups of a half-warp (fp64) /
  // warp (fp32) cover disjoint banks when they read the same (row slot, column) of their four matrices
  __host__ __device__ static constexpr int pad_bytes(int elems) { return (((G * ES) - (elems * ES) % 128) + 128) % 128; }
  static constexpr int STRIDE = SZ + pad_bytes(SZ) / ES;
  // finished 8-column block t = packed elements [col_off(8t), col_off(8t+8)): contiguous
  __host__ __device__ static constexpr int blk_off(int t) { return packed_col_off(NP, G * t); }
  __host__ __device__ static constexpr int blk_size(int t) { return packed_col_off(NP, G * t + G) - packed_col_off(NP, G * t); }
  static constexpr int OSZ = blk_size(0);  // the largest block
  static constexpr int OSTRIDE = OSZ + pad_bytes(OSZ) / ES;
};

template <typename T, int NP, int WARPS, bool IN_BULK, bool OUT_BULK>
struct PackedSmem {
  using Geo = PackedGeom<T, NP>;
  static constexpr int MPW = 4;
  static constexpr int BC = 2 * NP * MPW;                          // broadcast line, double buffered (elements)
  static constexpr int IN = IN_BULK ? MPW * Geo::STRIDE : 0;       // staged input, 4 matrices
  static constexpr int OUT = OUT_BULK ? MPW * Geo::OSTRIDE : 0;    // staged output block, 4 matrices
  static constexpr int per_warp_bytes = ((BC + IN + OUT) * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr size_t bytes = (size_t)WARPS * per_warp_bytes + 8 * WARPS + 128;
};

// n == NP, packed storage; IN_BULK requires 16-byte aligned matrices (checked by the launcher).
template <typename T, int NP, int WARPS, int MINB, bool STRIDED, bool IN_BULK, bool OUT_BULK, bool LOCKSTEP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
potrf_packed_kernel(BatchRef<T, STRIDED> APref, const int batchCount) {
  using Geo = PackedGeom<T, NP>;
  using Sm = PackedSmem<T, NP, WARPS, IN_BULK, OUT_BULK>;
  constexpr int G = 8, S = NP / G, MPW = 4, GH = 2;
  constexpr int PAIR = GH * 2;
  constexpr int BUF_STRIDE = NP * MPW;  // elements per broadcast buffer
  constexpr int ES = (int)sizeof(T);
  typedef typename Vec2T<T>::type V2;

  extern __shared__ __align__(128) unsigned char smem_pk[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int l = lane % G;
  const int g = lane / G;
  T *const wsm = reinterpret_cast<T *>(smem_pk + (size_t)warp * Sm::per_warp_bytes);
  T *const wbase = wsm + (g / GH) * ((NP / 2) * PAIR) + (g % GH) * 2;  // broadcast line, as in potrf_small.cuh
  T *const in_w = wsm + Sm::BC;
  T *const out_w = in_w + Sm::IN;
  const T *const in_g = in_w + g * Geo::STRIDE;
  T *const out_g = out_w + g * Geo::OSTRIDE;
  const uint32_t bar = smem_u32(smem_pk + (size_t)WARPS * Sm::per_warp_bytes) + 8 * warp;
  uint32_t parity = 0;

  const long nwb = ((long)batchCount + MPW - 1) / MPW;
  const long ncb = (nwb + WARPS - 1) / WARPS;
  const long last = (long)batchCount - 1;

  // four matrices of warp-batch wb into the staging buffer (inactive tails re-read the last matrix)
  auto issue_load = [&](long wb) {
    if (lane == 0) mbar_expect_tx(bar, MPW * Geo::SZ * ES);
    __syncwarp();
    if (lane < MPW) {
      long m = wb * MPW + lane;
      m = m < last ? m : last;
      bulk_g2s(smem_u32(in_w + lane * Geo::STRIDE), APref.at(m), Geo::SZ * ES, bar);
    }
  };

  if (IN_BULK) {
    if (lane == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long wb0 = (long)blockIdx.x * WARPS + warp;
    if (wb0 < nwb) issue_load(wb0);
  }

  for (long cb = blockIdx.x; cb < ncb; cb += gridDim.x) {
    if (LOCKSTEP) __syncthreads();
    const long wb = cb * WARPS + warp;
    if (wb >= nwb) continue;  // warp-uniform; only in the CTA's last round (no later __syncthreads is skipped by others: see below)
    const long mat = wb * MPW + g;
    const bool active = mat <= last;
    T *__restrict__ AP = APref.at(active ? mat : last);

#define KX_IDX(s_, c_) (G * (((s_) * ((s_) + 1)) / 2) + (c_))
    T a[G * (S * (S + 1)) / 2];

    // ---- lower triangle -> registers ---------------------------------------------------------
    if (IN_BULK) {
      mbar_wait(bar, parity);
      parity ^= 1;
#pragma unroll
      for (int col = 0; col < NP; ++col) {
#pragma unroll
        for (int s = col / G; s < S; ++s) {
          const int e = packed_col_off(NP, col) + (G * s - col);  // + l
          if (s > col / G) a[KX_IDX(s, col)] = lds_one(in_g + e + l);
          else {
            // diagonal slot: rows above the diagonal are not stored (the address belongs to the previous column)
            const T v = lds_one(in_g + e + l);   // e = col_off(col) - col % 8 >= 0
            a[KX_IDX(s, col)] = (l >= col % G) ? v : T(0);
          }
        }
      }
      // the buffer is free: start the load of this warp's next batch (the generic-proxy reads above are ordered
      // before the async-proxy writes by the fence + warp barrier)
      fence_async_smem();
      __syncwarp();
      const long nb = wb + (long)gridDim.x * WARPS;
      if (nb < nwb) issue_load(nb);
    } else {
      const T *pc = AP + l;
#pragma unroll
      for (int col = 0; col < NP; ++col) {
#pragma unroll
        for (int s = col / G; s < S; ++s) {
          const int e = packed_col_off(NP, col) + (G * s - col);
          T v = T(0);
          if (s > col / G) v = ldg_stream(pc + e);
          else ldg_stream_if(v, pc + e, l >= col % G);
          a[KX_IDX(s, col)] = v;
        }
      }
      // pull the NEXT batch of this warp into L2 meanwhile (contiguous 128-byte lines)
      const long nb = wb + (long)gridDim.x * WARPS;
#pragma unroll
      for (int q = 0; q < MPW; ++q) {
        const long m2 = nb * MPW + q;
        if (m2 <= last) {
          const char *p = reinterpret_cast<const char *>(APref.at(m2));
          for (int o = lane * 128; o < Geo::SZ * ES; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        }
      }
    }

    // ---- right-looking factorisation (identical to potrf_reg_kernel, EXACT) ------------------
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int t = j / G, c = j % G;
      const T d = shfl_seg<G>(a[KX_IDX(t, j)], c);
      const T r = rsqrt_t(d);
#pragma unroll
      for (int s = t; s < S; ++s) a[KX_IDX(s, j)] *= r;

      if (j + 1 < NP) {
        T *wbuf = wbase + (j & 1) * BUF_STRIDE;
#pragma unroll
        for (int s = t; s < S; ++s) {
          if (G * s + G - 1 > j) {
            const int k = G * s + l;
            wbuf[(k >> 1) * PAIR + (k & 1)] = a[KX_IDX(s, j)];
          }
        }
        __syncwarp();
#pragma unroll
        for (int p = (j + 1) / 2; p < NP / 2; ++p) {
          const V2 v2 = *reinterpret_cast<const V2 *>(wbuf + p * PAIR);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * p + h;
            if (k > j) {
              const T v = h ? v2.y : v2.x;
#pragma unroll
              for (int s = k / G; s < S; ++s) a[KX_IDX(s, k)] = fma_t(-a[KX_IDX(s, j)], v, a[KX_IDX(s, k)]);
            }
          }
        }
      }

      // ---- a finished block of 8 columns leaves as soon as it is final --------------------------
      if (c == G - 1) {
        if (OUT_BULK) {
          // the previous block's bulk store must have finished READING the staging buffer
          if (lane < MPW) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int col = j - (G - 1); col <= j; ++col) {
#pragma unroll
            for (int s = t; s < S; ++s) {
              const int e = packed_col_off(NP, col) - Geo::blk_off(t) + (G * s - col);
              if (s > t || l >= col % G) out_g[e + l] = a[KX_IDX(s, col)];
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane < MPW && wb * MPW + lane <= last) {
            bulk_s2g(APref.at(wb * MPW + lane) + Geo::blk_off(t), smem_u32(out_w + lane * Geo::OSTRIDE), Geo::blk_size(t) * ES);
            bulk_commit();
          }
        } else {
          T *pst = AP + l;
#pragma unroll
          for (int col = j - (G - 1); col <= j; ++col) {
#pragma unroll
            for (int s = t; s < S; ++s) {
              const int e = packed_col_off(NP, col) + (G * s - col);
              stg_stream_if(pst + e, a[KX_IDX(s, col)], active && (s > t || l >= col % G));
            }
          }
        }
      }
    }
    __syncwarp();  // the broadcast buffers are reused by the next warp-batch
#undef KX_IDX
  }
  if (OUT_BULK) {
    if (lane < MPW) bulk_wait0();  // bulk stores are complete before the CTA's shared memory goes away
  }
}

// Generic packed kernel: any n <= NP (identity padding in registers), any alignment, optional LAPACK info; plain
// predicated loads / stores with packed addressing.  Same arithmetic.
template <typename T, int NP, int WARPS, int MINB, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, MINB)
potrf_packed_generic_kernel(const int n, BatchRef<T, STRIDED> APref, const int batchCount, int *__restrict__ info,
                            const int info_mode) {
  constexpr int G = 8, S = NP / G, MPW = 4, GH = 2;
  constexpr int PAIR = GH * 2;
  constexpr int BUF_STRIDE = NP * MPW;
  typedef typename Vec2T<T>::type V2;
  __shared__ __align__(16) T bc[WARPS * 2 * BUF_STRIDE];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int l = lane % G;
  const int g = lane / G;
  T *const wbase = bc + warp * (2 * BUF_STRIDE) + (g / GH) * ((NP / 2) * PAIR) + (g % GH) * 2;
  const long nwb = ((long)batchCount + MPW - 1) / MPW;
  const long last = (long)batchCount - 1;
  for (long wb = (long)blockIdx.x * WARPS + warp; wb < nwb; wb += (long)gridDim.x * WARPS) {
    const long mat = wb * MPW + g;
    const bool active = mat <= last;
    T *__restrict__ AP = APref.at(active ? mat : last);
#define KX_IDX(s_, c_) (G * (((s_) * ((s_) + 1)) / 2) + (c_))
    T a[G * (S * (S + 1)) / 2];
#pragma unroll
    for (int col = 0; col < NP; ++col) {
      const int co = col * n - (col * (col - 1)) / 2 - col;  // + row
#pragma unroll
      for (int s = col / G; s < S; ++s) {
        const int row = G * s + l;
        T v = (row == col) ? T(1) : T(0);  // identity padding
        ldg_stream_if(v, AP + co + row, row < n && col < n && row >= col);
        a[KX_IDX(s, col)] = v;
      }
    }
    int bad = 0;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int t = j / G, c = j % G;
      const T d = shfl_seg<G>(a[KX_IDX(t, j)], c);
      if (info_mode && bad == 0 && j < n && !(d > T(0))) bad = j + 1;
      const T r = rsqrt_t(d);
#pragma unroll
      for (int s = t; s < S; ++s) a[KX_IDX(s, j)] *= r;
      if (j + 1 < NP) {
        T *wbuf = wbase + (j & 1) * BUF_STRIDE;
#pragma unroll
        for (int s = t; s < S; ++s) {
          if (G * s + G - 1 > j) {
            const int k = G * s + l;
            wbuf[(k >> 1) * PAIR + (k & 1)] = a[KX_IDX(s, j)];
          }
        }
        __syncwarp();
#pragma unroll
        for (int p = (j + 1) / 2; p < NP / 2; ++p) {
          const V2 v2 = *reinterpret_cast<const V2 *>(wbuf + p * PAIR);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * p + h;
            if (k > j) {
              const T v = h ? v2.y : v2.x;
#pragma unroll
              for (int s = k / G; s < S; ++s) a[KX_IDX(s, k)] = fma_t(-a[KX_IDX(s, j)], v, a[KX_IDX(s, k)]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int col = 0; col < NP; ++col) {
      const int co = col * n - (col * (col - 1)) / 2 - col;
#pragma unroll
      for (int s = col / G; s < S; ++s) {
        const int row = G * s + l;
        stg_stream_if(AP + co + row, a[KX_IDX(s, col)], active && row < n && col < n && row >= col);
      }
    }
    if (info_mode && active && l == 0) info[mat] = bad;
    __syncwarp();
#undef KX_IDX
  }
}

// ---- n = 8 (and fp32 n = 16): ONE LANE PER MATRIX ------------------------------------------------------------------
// The 8-lanes-per-matrix mapping above spends its time on cross-lane traffic when a matrix is only 36 values
// (ncu, round 1: shared-memory data pipe 87 % busy at n = 8).  Here a warp owns 32 consecutive packed matrices --
// 32*SZ contiguous elements when strideAP == SZ -- and
//   1. reads them with fully coalesced loads (flat element f = k*32 + lane, k = 0..SZ-1),
//   2. transposes through shared memory with one element of padding per matrix (stride SZ+1: odd word stride,
//      conflict-free both ways) so that lane m ends up with all SZ values of matrix m in registers,
//   3. factors with no cross-lane traffic at all (fully unrolled, same recurrence: d = a_jj, r = rsqrt(d),
//      column *= r, a_ik -= l_ij l_kj),
//   4. transposes back and stores coalesced.
// Requires strideAP == SZ (contiguous batch); other strides take the generic kernel.
template <typename T, int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
potrf_packed_lane_kernel(T *__restrict__ AP, const int batchCount) {
  constexpr int SZ = packed_size(N);
  constexpr int LD = SZ + 1;  // padded per-matrix stride in shared memory
  extern __shared__ __align__(128) unsigned char smem_pk[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  T *const sm = reinterpret_cast<T *>(smem_pk) + (size_t)warp * 32 * LD;
  const long nwb = ((long)batchCount + 31) / 32;
  for (long wb = (long)blockIdx.x * WARPS + warp; wb < nwb; wb += (long)gridDim.x * WARPS) {
    T *base = AP + wb * 32 * SZ;
    const long rem = (long)batchCount - wb * 32;
    const int cnt = rem < 32 ? (int)rem : 32;      // matrices in this warp-batch
    const int total = cnt * SZ;                    // valid flat elements
    T v[SZ];
#pragma unroll
    for (int k = 0; k < SZ; ++k) {
      const int f = k * 32 + lane;
      v[k] = T(0);
      ldg_stream_if(v[k], base + f, f < total);
    }
#pragma unroll
    for (int k = 0; k < SZ; ++k) {
      const int f = k * 32 + lane;
      sm[f + f / SZ] = v[k];                       // matrix f / SZ, element f % SZ -> (f / SZ) * LD + f % SZ
    }
    __syncwarp();
    T a[SZ];
#pragma unroll
    for (int e = 0; e < SZ; ++e) a[e] = sm[lane * LD + e];
    if (lane >= cnt) {                             // tail lanes factor the identity (finite arithmetic, never stored)
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = j; i < N; ++i) a[packed_col_off(N, j) + i - j] = (i == j) ? T(1) : T(0);
    }
#define KX_P(i_, j_) a[packed_col_off(N, (j_)) + (i_) - (j_)]
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T r = rsqrt_t(KX_P(j, j));
#pragma unroll
      for (int i = j; i < N; ++i) KX_P(i, j) *= r;
#pragma unroll
      for (int k = j + 1; k < N; ++k) {
        const T nk = -KX_P(k, j);
#pragma unroll
        for (int i = k; i < N; ++i) KX_P(i, k) = fma_t(KX_P(i, j), nk, KX_P(i, k));
      }
    }
#undef KX_P
    __syncwarp();
#pragma unroll
    for (int e = 0; e < SZ; ++e) sm[lane * LD + e] = a[e];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SZ; ++k) {
      const int f = k * 32 + lane;
      stg_stream_if(base + f, sm[f + f / SZ], f < total);
    }
    __syncwarp();
  }
}

// ---- pack / unpack: full column-major (lda, stride) <-> packed lower --------------------------------------------
// One warp per matrix, lane = packed element index (coalesced on the packed side).  UNPACK writes only the lower
// triangle of A (the strict upper triangle and the padding keep their bits, like potrf's output).
template <typename T, bool STRIDED, bool UNPACK>
__global__ void __launch_bounds__(256)
tri_pack_kernel(const int n, BatchRef<T, STRIDED> Aref, const int lda, BatchRef<T, STRIDED> APref, const int batchCount) {
  const int lane = threadIdx.x & 31;
  const long wpg = (long)gridDim.x * (blockDim.x >> 5);
  const int sz = packed_size(n);
  for (long b = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < batchCount; b += wpg) {
    T *A = Aref.at(b);
    T *AP = APref.at(b);
    // walk the packed index; (i, j) recovered incrementally per lane
    for (int e = lane; e < sz; e += 32) {
      // column j: largest j with col_off(j) <= e
      int j = 0, off = 0;
      while (off + (n - j) <= e) {
        off += n - j;
        ++j;
      }
      const int i = j + (e - off);
      if (UNPACK) A[i + (long)j * lda] = AP[e];
      else AP[e] = A[i + (long)j * lda];
    }
  }
}

}  // namespace kblasx
