// kernels/potrf_smem.cuh -- fp64 batched Cholesky for 32 < n <= 256 with the factor RESIDENT IN SHARED MEMORY (sm_100a).
//
// VERDICT round 1, item 3.  The one-warp-per-matrix kernel (kernels/potrf_panel_mma.cuh) fetches every DMMA operand from
// global memory; with ~1200 matrices in flight the factored columns do not survive in L2 and are re-read from DRAM
// (ncu, n = 256: 15.7 GB read for 4.3 GB compulsory, 40 % of the stalls on the scoreboard).  Here ONE CTA owns one
// matrix and every element crosses the memory system once:
//   * left-looking over 32-column panels; panel J+1 needs  L[I][K], I > J, K <= J  -- row block I of the factor is dead
//     as soon as panel I is done, so the LIVE part of L is at most (nblk-J-1)(J+2) blocks of 32 x 32 (20 of the 36 blocks of
//     a 256 x 256 matrix).  Blocks live in 8 KiB shared-memory slots handed out by a static interval colouring computed on
//     the host (SmemPotrfPlan): 24 slots = 192 KiB hold the live factor, the panel being processed AND the next panel,
//     which streams in with cp.async (LDGSTS) while the current one is being computed -- global loads never stall a warp.
//   * 32 x 32 blocks are stored column-major with the row index XOR-swizzled by 4*(col % 4): the DMMA m8n8k4 fragment
//     loads (8 rows x 4 columns per instruction), the row-per-lane accesses of the substitution and the 16-byte cp.async
//     writes are all bank-conflict free without padding.
//   * update  P[I] -= sum_K L[I][K] L[J][K]^T  on FP64 DMMA (mma.sync m8n8k4), operands from shared memory, in 8-row x
//     32-column strips with two K-interleaved accumulator sets (the dependent-DMMA latency is ~150 cycles on B200 and a
//     sub-partition runs one warp's DMMAs at a time: profiles/r01_microbench_pipes.txt).  Strips of the off-diagonal
//     blocks are handed out through a shared counter, so the late panels (one or two row blocks, long inner dimension)
//     still occupy every warp -- the first version, with whole 16-row units assigned statically, spent 48 % of its stall
//     samples in the CTA barrier (profiles/r02_ncu_dpotrf256_smem_potrf_smem_W8.json);
//   * look-ahead: the four strips of the NEXT diagonal block go to warps 0..3 first, warp 0 then factors it (row per lane;
//     the next column's pivot is formed in registers and its rsqrt overlaps the current column's shared-memory broadcast)
//     while the other warps update the rest of the panel; the rows below are solved against it (row per lane) and leave
//     for global memory as whole 256-byte column segments.
// Replaces the reference's recursion for these sizes: 14 / 41 / 108 launches with every tile making 5-8 global round trips
// (Xpotrf_batch_drivers.cuh:94-133, Xsyrk_batch_drivers.cuh:234-323).
#pragma once

#include <cstdint>
#include "common.cuh"
#include "potrf_panel_mma.cuh"  // dmma_m8n8k4

namespace kblasx {

// host + device view of the slot assignment: slot[I][K] for K <= I < nblk
struct SmemPotrfPlan {
  unsigned char slot[8][8];
  int nslots;
};

// Static allocation with one panel of look-ahead (the order of events of potrf_smem_kernel):
//   prologue: panels 0 and 1 are allocated;  iteration J: S(.,J) -> free (J,J) -> allocate panel J+2 ->
//   U(.,J+1), F(J+1) -> free (J+1, K <= J).
inline SmemPotrfPlan plan_potrf_slots(int nblk) {
  SmemPotrfPlan p;
  bool used[64] = {};
  int hi = 0;
  auto take = [&]() {
    int s = 0;
    while (used[s]) ++s;
    used[s] = true;
    if (s + 1 > hi) hi = s + 1;
    return (unsigned char)s;
  };
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 8; ++k) p.slot[i][k] = 0;
  for (int I = 0; I < nblk; ++I) p.slot[I][0] = take();
  for (int I = 1; I < nblk; ++I) p.slot[I][1] = take();
  for (int J = 0; J < nblk; ++J) {
    used[p.slot[J][J]] = false;
    for (int I = J + 2; I < nblk; ++I) p.slot[I][J + 2] = take();
    if (J + 1 < nblk)
      for (int K = 0; K <= J; ++K) used[p.slot[J + 1][K]] = false;
  }
  p.nslots = hi;
  return p;
}

struct PotrfSmemGeom {
  static constexpr int NB = 32;
  static constexpr int BLK = NB * NB;  // doubles per slot
  static size_t bytes(int nslots) { return (size_t)nslots * BLK * sizeof(double) + NB * sizeof(double) + 64 + 64; }  // blocks, 1/diag, slot table, strip counter
};

// element (r, c) of a swizzled 32 x 32 block
__device__ __forceinline__ int sw_idx(int r, int c) { return c * 32 + (r ^ ((c & 3) << 2)); }

__device__ __forceinline__ void cp_async16_zfill(double *smem_dst, const double *gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(double *smem_dst, const double *gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// KX_SMEM_TRACE (tools/trace_smem.cu only): thread 0 / lane 0 of warps 0 and 1 record clock64() at every phase boundary
#ifdef KX_SMEM_TRACE
#define KX_TRACE_PARAM , long long *__restrict__ trace
#define KX_T(slot_)                                                                     \
  do {                                                                                  \
    if (lane == 0 && warp < 2 && blockIdx.x < 4) trace[(blockIdx.x * 2 + warp) * 128 + (slot_)] = clock64(); \
  } while (0)
#else
#define KX_TRACE_PARAM
#define KX_T(slot_) do { } while (0)
#endif

template <int WARPS, int MINB, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, MINB)
potrf_smem_kernel(const int n, BatchRef<double, STRIDED> Aref, const int lda, const int batchCount, int *__restrict__ info,
                  const int info_mode, const SmemPotrfPlan plan KX_TRACE_PARAM) {
  constexpr int NB = 32, BLK = NB * NB, THREADS = WARPS * 32;
  extern __shared__ __align__(128) unsigned char smem_ps[];
  double *const blocks = reinterpret_cast<double *>(smem_ps);
  double *const invd = blocks + (size_t)plan.nslots * BLK;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (n + NB - 1) / NB;
  double *__restrict__ A = Aref.at(blockIdx.x);
  const bool al16 = ((reinterpret_cast<uintptr_t>(A) | ((uintptr_t)lda * sizeof(double))) & 15) == 0;
  int bad = 0;

  // the slot table goes to shared memory (indexing the kernel parameter dynamically would copy it to local memory)
  unsigned char *const slot_tab = reinterpret_cast<unsigned char *>(invd + NB);
  for (int i = tid; i < 64; i += THREADS) slot_tab[i] = plan.slot[i >> 3][i & 7];
  __syncthreads();
  auto blk = [&](int I, int K) -> double * { return blocks + (int)slot_tab[I * 8 + K] * BLK; };

  // ---- asynchronous load of the blocks (I, K), I = K .. nblk-1, of panel K (zero-filled outside the matrix) --------
  auto load_panel = [&](int K) {
    if (K < nblk) {
      const int nb = nblk - K;
      if (al16) {
        for (int idx = tid; idx < nb * 512; idx += THREADS) {
          const int b = idx >> 9, rem = idx & 511, c = rem >> 4, r = (rem & 15) * 2;
          const int grow = NB * (K + b) + r, gcol = NB * K + c;
          int bytes = (gcol < n) ? (grow + 1 < n ? 16 : (grow < n ? 8 : 0)) : 0;
          const double *src = A + (bytes ? grow : 0) + (long)(bytes ? gcol : 0) * lda;
          cp_async16_zfill(blk(K + b, K) + sw_idx(r, c), src, bytes);
        }
      } else {
        for (int idx = tid; idx < nb * 1024; idx += THREADS) {
          const int b = idx >> 10, rem = idx & 1023, c = rem >> 5, r = rem & 31;
          const int grow = NB * (K + b) + r, gcol = NB * K + c;
          const int bytes = (gcol < n && grow < n) ? 8 : 0;
          const double *src = A + (bytes ? grow : 0) + (long)(bytes ? gcol : 0) * lda;
          cp_async8_zfill(blk(K + b, K) + sw_idx(r, c), src, bytes);
        }
      }
    }
    cp_async_commit();  // one group per panel, also when empty: keeps the wait_group arithmetic uniform
  };

  // identity padding of a ragged last diagonal block (rows / columns >= n were zero-filled)
  auto pad_diag = [&](int J) {
    const int jb = n - NB * J;
    if (jb < NB) {
      double *D = blk(J, J);
      for (int r = jb + tid; r < NB; r += THREADS) D[sw_idx(r, r)] = 1.0;
    }
  };

  // ---- F(J): factor the diagonal block in place (one warp, lane = row), store its lower triangle -------------------
  // The pivot of column j+1 is formed from registers alone -- lane j+1 holds L[j+1][j] (its own p[j]) and p[j+1] -- and
  // broadcast BEFORE column j goes through shared memory, so the shuffle -> rsqrt (62 cycles) chain of the next column
  // overlaps the trailing update of this one instead of queueing behind its STS / LDS round trip.
  auto factor_diag = [&](int J) {
    double *D = blk(J, J);
    const int jb = (n - NB * J < NB) ? (n - NB * J) : NB;
    double p[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) p[c] = D[sw_idx(lane, c)];
    double dcur = __shfl_sync(0xffffffffu, p[0], 0);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (info_mode && bad == 0 && j < jb && !(dcur > 0.0)) bad = NB * J + j + 1;
      const double r = rsqrt(dcur);
      p[j] *= r;
      if (j + 1 < NB) dcur = __shfl_sync(0xffffffffu, fma(-p[j], p[j], p[j + 1]), j + 1);
      D[sw_idx(lane, j)] = p[j];
      if (lane == j) invd[j] = r;
      __syncwarp();
      const int sw = (j & 3) << 2;
#pragma unroll
      for (int k = (j + 1) & ~1; k < NB; k += 2) {
        const double2 l2 = lds_pair(D + j * 32 + (k ^ sw));  // L[k][j], L[k+1][j] (the swizzle keeps even pairs together)
        if (k > j) p[k] = fma(-p[j], l2.x, p[k]);
        p[k + 1] = fma(-p[j], l2.y, p[k + 1]);
      }
    }
    const int row = NB * J + lane;
    double *g = A + row + (long)(NB * J) * lda;
#pragma unroll
    for (int c = 0; c < NB; ++c) stg_stream_if(g + (long)c * lda, p[c], row < n && c <= lane && c < jb);
  };

  // ---- S(I, J): rows of block (I, J) solved against L[J][J]^T (one warp, lane = row), stored to smem and global ----
  auto solve_block = [&](int I, int J) {
    double *P = blk(I, J);
    const double *D = blk(J, J);
    double p[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) p[c] = P[sw_idx(lane, c)];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      p[j] *= lds_one(invd + j);
      const int sw = (j & 3) << 2;
#pragma unroll
      for (int k = (j + 1) & ~1; k < NB; k += 2) {
        const double2 l2 = lds_pair(D + j * 32 + (k ^ sw));
        if (k > j) p[k] = fma(-p[j], l2.x, p[k]);
        p[k + 1] = fma(-p[j], l2.y, p[k + 1]);
      }
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) P[sw_idx(lane, c)] = p[c];
    const int row = NB * I + lane;
    const int jb = (n - NB * J < NB) ? (n - NB * J) : NB;
    double *g = A + row + (long)(NB * J) * lda;
#pragma unroll
    for (int c = 0; c < NB; ++c) stg_stream_if(g + (long)c * lda, p[c], row < n && c < jb);
  };

  // ---- U strip: rows 8q .. 8q+7 of block (I, Jn) -= sum_{K < Jn} L[I][K] L[Jn][K]^T on DMMA.  8-row strips (4 per block)
  // keep all warps busy in the late panels, where only one or two row blocks are left but the inner dimension is long;
  // two K-interleaved accumulator sets x 4 column tiles = 8 independent DMMA chains per warp.
  auto update_strip = [&](int I, int q, int Jn) {
    const int fr = lane >> 2, fk = lane & 3, sw = fk << 2;
    double acc[2][4][2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) acc[s2][cb][0] = acc[s2][cb][1] = 0.0;
    const int arow = (8 * q + fr) ^ sw;
    for (int K = 0; K < Jn; ++K) {
      const double *a_blk = blk(I, K) + fk * 32 + arow;
      const double *b_blk = blk(Jn, K) + fk * 32;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        double bf[4];
        const double af = a_blk[ks * 128];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) bf[cb] = b_blk[ks * 128 + ((8 * cb + fr) ^ sw)];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma_m8n8k4(acc[ks & 1][cb][0], acc[ks & 1][cb][1], af, bf[cb]);
      }
    }
    double *P = blk(I, Jn);
#pragma unroll
    for (int cb = 0; cb < 4; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * q + fr, c = 8 * cb + 2 * fk + e;
        P[sw_idx(r, c)] -= acc[0][cb][e] + acc[1][cb][e];
      }
  };

  // off-diagonal strips of panel Jn, handed out through a shared counter (whoever is free takes the next one)
  int *const strip_ctr = reinterpret_cast<int *>(slot_tab + 64);
  auto offdiag_strips = [&](int Jn) {
    const int nstrips = 4 * (nblk - Jn - 1);
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(strip_ctr, 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= nstrips) break;
      update_strip(Jn + 1 + (u >> 2), u & 3, Jn);
    }
  };

  // ================================================================================================================
  constexpr int DW = WARPS < 4 ? WARPS : 4;  // warps that share the 4 strips of the next diagonal block
  KX_T(0);
  load_panel(0);
  load_panel(1);
  cp_async_wait<1>();  // panel 0 has landed (this thread's copies) ...
  __syncthreads();     // ... and everybody's
  KX_T(1);
  pad_diag(0);
  __syncthreads();
  if (warp == 0) factor_diag(0);
  KX_T(2);
  __syncthreads();
  KX_T(3);

  for (int J = 0; J < nblk; ++J) {
    // ---- S: rows below the diagonal block of panel J ---------------------------------------------------------------
    for (int I = J + 1 + warp; I < nblk; I += WARPS) solve_block(I, J);
    KX_T(8 + 8 * J + 0);  // own S blocks done
    if (tid == 0) *strip_ctr = 0;
    __syncthreads();
    KX_T(8 + 8 * J + 1);  // S phase over
    if (J + 1 >= nblk) break;
    // ---- look-ahead load of panel J+2 (its slots are free now), panel J+1 must have landed ----------------------
    load_panel(J + 2);
    cp_async_wait<1>();
    __syncthreads();
    pad_diag(J + 1);
    __syncthreads();
    KX_T(8 + 8 * J + 2);  // next panel resident
    // ---- U + F: panel J+1 -= L[.][0..J] L[J+1][0..J]^T.  The first DW warps take the 4 strips of the next diagonal block,
    //      warp 0 then factors it WHILE everybody else works through the off-diagonal strips -----------------------------
    const int Jn = J + 1;
    if (warp < DW) {
      for (int q = warp; q < 4; q += DW) update_strip(Jn, q, Jn);
      KX_T(8 + 8 * J + 3);  // own diagonal strip done
      if (DW > 1) asm volatile("bar.sync 1, %0;" ::"n"(32 * DW) : "memory");
      else __syncwarp();
      KX_T(8 + 8 * J + 4);  // diagonal block updated
      if (warp == 0) {
        factor_diag(Jn);
        if (WARPS == 1) offdiag_strips(Jn);  // one warp per matrix: nobody else to take them
      } else {
        offdiag_strips(Jn);
      }
    } else {
      offdiag_strips(Jn);
    }
    KX_T(8 + 8 * J + 5);  // F (warp 0) / off-diagonal strips (warp 1) done
    __syncthreads();
    KX_T(8 + 8 * J + 6);
  }
  cp_async_wait<0>();
  if (info_mode && tid == 0) info[blockIdx.x] = bad;
}

}  // namespace kblasx
