// kernels/trsm_dual.cuh -- batched triangular solves (all four side / trans variants) and the fused POTRS
// with a full NP x NP factor (NP = 16, 24, 32): TWO right-hand-side vectors per lane and two problems per
// warp (sm_100a).
//
// Why: ncu of the one-vector-per-lane kernel (kernels/trsm_small.cuh) on dpotrs n = 32 shows the
// shared-memory data pipe 86 % busy (profiles/r01_ncu_dpotrs32_tri_solve_small.json): every broadcast
// LDS of the factor feeds only 2 FMAs per lane, and staging the factor costs 32 LDG + 32 STS per
// problem.  Here
//   * a half-warp owns one (matrix, 32-vector slab) task and every lane owns rows lg and lg+16 of B, so one
//     broadcast LDS.128 (2 fp64 / 4 fp32 factor entries) feeds 4 / 8 FMAs per lane; the two
//     half-warps read their own factor copies, placed 16 B (mod 128) apart so that the two
//     addresses of one instruction never share a bank;
//   * the factor goes global -> shared memory with cp.async (LDGSTS): no staging registers, no STS,
//     and it is in flight together with the 64 predicated loads of B;
//   * forward and backward substitution run on the same registers (POTRS = one pass over B);
//   * side L (vectors = columns of B): B is read and written coalesced along its rows and transposed
//     through a padded shared-memory tile per problem, instead of the reference's stride-ldb per-lane
//     accesses (Xtrsm_batch_kernels.cuh:580-589).
// vec > 32 is handled by 32-vector slabs; ragged k goes to the older kernels in the dispatch.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"  // TriOp, sched_fence
#include "vec16.cuh"

namespace kblasx {

// L2 prefetch of everything one (matrix, 32-vector slab) task of a k <= NP solve reads: the lower part of the factor columns
// and the slab of B (side R: 32 rows of NP columns, side L: NP rows of 32 columns).  Hints only, one 128-byte line per
// instruction; `id` of `nl` cooperating lanes.  Measured on B200 (2^20 problems, two-vector kernel, with / without):
// dpotrs n = 32 4.27 / 4.97 ms, n = 24 2.52 / 2.87; dtrsm R n = 32 3.44 / 3.75; strsm R n = 32 1.90 / 2.72 -- but dpotrs
// n = 16 1.19 / 1.10: small fp64 tasks finish before the hint pays, so the launcher switches it off there.
template <typename T, int NP, bool LEFT>
__device__ __forceinline__ void prefetch_solve_task_l2(const T *A, const int lda, const T *B, const int ldb, const int pv0,
                                                       const int vec, const int id, const int nl) {
  constexpr int ES = (int)sizeof(T);
  const char *pa = reinterpret_cast<const char *>(A);
  const char *pb = reinterpret_cast<const char *>(B);
  for (int c = id; c < NP; c += nl) {  // factor: column c, rows c .. NP-1
    const char *col = pa + (long)c * lda * ES;
    for (int off = (c * ES) & ~127; off < NP * ES; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + off));
  }
  if (!LEFT) {  // B: rows pv0 .. pv0+31 of columns 0 .. NP-1
    const int nr = (vec - pv0 < 32) ? (vec - pv0) : 32;
    for (int c = id; c < NP; c += nl) {
      const char *col = pb + ((long)c * ldb + pv0) * ES;
      for (int off = 0; off < nr * ES; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + off));
    }
  } else {      // B: rows 0 .. NP-1 of columns pv0 .. pv0+31
    for (int c = id; c < 32; c += nl) {
      if (pv0 + c < vec) {
        const char *col = pb + (long)(pv0 + c) * ldb * ES;
        for (int off = 0; off < NP * ES; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + off));
      }
    }
  }
}

template <typename T, int NP, bool LEFT = false>
struct TriDualSmem {
  static constexpr int VW = 16 / (int)sizeof(T);
  static constexpr int per_problem = NP * NP + NP + VW;  // factor + reciprocal diagonal + bank-skew pad
  static constexpr int tile_stride = NP + 1;             // odd: conflict-free transposed reads
  static constexpr int tile_stride16 = NP + VW;          // 16-byte aligned pitch of the VEC16 path (conflict-free LDS.128 / STS.128)
  static constexpr int tile = LEFT ? 32 * tile_stride16 + (sizeof(T) == 4 ? 16 : 0) : 0;  // side L: 32 vectors x NP entries (+ bank skew between the two problems, fp32)
  static constexpr int per_warp = 2 * (per_problem + tile);
  static_assert((per_problem * sizeof(T)) % 16 == 0 && (per_problem * sizeof(T)) % 128 != 0,
                "the two factor copies of a warp must sit a non-zero multiple of 16 B apart (mod 128)");
};

#ifndef KX_DUAL_MINB64
#define KX_DUAL_MINB64 2
#endif
#ifndef KX_DUAL_FENCE64
#define KX_DUAL_FENCE64 2
#endif
#ifndef KX_DUAL_FENCE32
#define KX_DUAL_FENCE32 4
#endif
// VEC16 (side L only; lda, ldb multiples of 16 bytes, 16-byte aligned operands -- the launcher checks): the slab of B moves
// global <-> shared as 16-byte chunks (cp.async in, LDS.128 + streaming STG.128 out) and a lane reads / writes its two
// columns with 128-bit shared accesses, instead of one 4- or 8-byte LDG / STS / LDS / STG per element each way.
template <typename T, int NP, bool LEFT, int OP, int WARPS, bool STRIDED, bool VEC16 = false>
__global__ void __launch_bounds__(WARPS * 32, (sizeof(T) == 8 ? KX_DUAL_MINB64 : (NP > 24 ? 3 : 4)) * 4 / WARPS)
tri_solve_dual_kernel(const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda, BatchRef<T, STRIDED> Bref,
                      const int ldb, const int batchCount, const int slabs, const int ahead) {
  constexpr int VW = 16 / (int)sizeof(T);
  constexpr int NV = NP / VW;
  constexpr int SE = SectorElems<T>::value;
  constexpr int FSZ = TriDualSmem<T, NP>::per_problem;
  constexpr int FENCE = sizeof(T) == 8 ? KX_DUAL_FENCE64 : KX_DUAL_FENCE32;  // columns between scheduling fences (bounds ptxas' LDS look-ahead)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 4, lg = lane & 15;
  T *Ls = reinterpret_cast<T *>(smem_raw) + warp * TriDualSmem<T, NP, LEFT>::per_warp + g * FSZ;
  T *invd = Ls + NP * NP;
  constexpr int TS = TriDualSmem<T, NP, LEFT>::tile_stride;
  T *tile = reinterpret_cast<T *>(smem_raw) + warp * TriDualSmem<T, NP, LEFT>::per_warp + 2 * FSZ + g * TriDualSmem<T, NP, LEFT>::tile;

  const long ntask = (long)batchCount * slabs;
  const long task = ((long)blockIdx.x * WARPS + warp) * 2 + g;  // (matrix, 32-vector slab)
  const bool live = task < ntask;
  const long tsafe = live ? task : ntask - 1;
  const long mat = tsafe / slabs;
  const int v0 = (int)(tsafe % slabs) * 32;
  const T *__restrict__ A = Aref.at(mat);
  T *__restrict__ B = Bref.at(mat);

  // ---- factor: global -> shared, asynchronously; sectors wholly above the diagonal are skipped ----------
#pragma unroll
  for (int c = 0; c < NP; ++c)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = lg + 16 * hh;
      if (16 * hh < NP) cp_async_elem(Ls + r + c * NP, A + r + (long)c * lda, r < NP && ((r | (SE - 1)) >= c));
    }
  // ---- my two vectors (v0 + lg and v0 + lg + 16) ------------------------------------------------------------
  T x0[NP], x1[NP];
  const int my0 = v0 + lg, my1 = v0 + lg + 16;
  const bool h0 = live && my0 < vec, h1 = live && my1 < vec;
  if (!LEFT) {
    // side R: vector = row of B; all 2 NP loads in flight
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      x0[j] = T(0);
      x1[j] = T(0);
      ldg_stream_if(x0[j], B + my0 + (long)j * ldb, h0);
      ldg_stream_if(x1[j], B + my1 + (long)j * ldb, h1);
    }
  } else if (VEC16) {
    constexpr int TSV = TriDualSmem<T, NP, LEFT>::tile_stride16;
    const T *Bv = B + (long)v0 * ldb;
    const int ncol = !live ? 0 : ((vec - v0 < 32) ? (vec - v0) : 32);
#pragma unroll
    for (int i0 = 0; i0 < 32 * NV; i0 += 16) {
      const int i = i0 + lg, c = i / NV, r0 = (i % NV) * VW;
      cp_async16_if(tile + c * TSV + r0, Bv + (long)c * ldb + r0, c < ncol);
    }
  } else {
    // side L: vector = column of B.  Lane lg reads rows lg and lg+16 of 16 columns at a time (coalesced) and
    // parks them in the tile as tile[column][row]; the vectors are then read back along the rows.
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      T t0[16], t1[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const bool hc = live && (v0 + c0 + c) < vec;
        t0[c] = T(0);
        t1[c] = T(0);
        ldg_stream_if(t0[c], B + lg + (long)(v0 + c0 + c) * ldb, hc);
        ldg_stream_if(t1[c], B + lg + 16 + (long)(v0 + c0 + c) * ldb, hc && (lg + 16 < NP));
      }
      sched_fence();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        tile[(c0 + c) * TS + lg] = t0[c];
        if (NP > 16 && lg + 16 < NP) tile[(c0 + c) * TS + lg + 16] = t1[c];
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      x0[j] = tile[lg * TS + j];
      x1[j] = tile[(lg + 16) * TS + j];
    }
  }
  // Pull the operands of the tasks that a LATER CTA of this grid will own into L2 (`ahead` = CTAs resident on the whole
  // GPU: the CTA that takes this one's place starts with L2 hits instead of DRAM round trips).
  if (ahead > 0) {
    const long ptask = task + (long)ahead * WARPS * 2;
    if (ptask < ntask) {
      const long pmat = ptask / slabs;
      prefetch_solve_task_l2<T, NP, LEFT>(Aref.at(pmat), lda, Bref.at(pmat), ldb, (int)(ptask % slabs) * 32, vec, lg, 16);
    }
  }
  cp_async_wait_all();
  __syncwarp();
  if (LEFT && VEC16) {  // my two columns of the slab = rows lg and lg + 16 of the tile
    constexpr int TSV = TriDualSmem<T, NP, LEFT>::tile_stride16;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      T c0[VW], c1[VW];
      lds_vec(c0, tile + lg * TSV + q * VW);
      lds_vec(c1, tile + (lg + 16) * TSV + q * VW);
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        x0[q * VW + e] = c0[e];
        x1[q * VW + e] = c1[e];
      }
    }
  }
  // reciprocal diagonal: one division per lane and diagonal entry
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int d = lg + 16 * hh;
    if (d < NP) invd[d] = T(1) / Ls[d + d * NP];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    x0[j] *= alpha;
    x1[j] *= alpha;
  }

  if (OP == TRI_FORWARD || OP == TRI_BOTH) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (j % FENCE == 0) sched_fence();
      const T dinv = lds_one(invd + j);
      x0[j] *= dinv;
      x1[j] *= dinv;
      const T n0 = -x0[j], n1 = -x1[j];
#pragma unroll
      for (int v = (j + 1) / VW; v < NV; ++v) {
        T c[VW];
        lds_vec(c, Ls + v * VW + j * NP);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const int i = v * VW + e;
          if (i > j) {
            x0[i] = fma_t(n0, c[e], x0[i]);
            x1[i] = fma_t(n1, c[e], x1[i]);
          }
        }
      }
    }
  }
  if (OP == TRI_BACKWARD || OP == TRI_BOTH) {
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
      if ((NP - 1 - j) % FENCE == 0) sched_fence();
      const T dinv = lds_one(invd + j);
      T a0[2] = {x0[j], T(0)}, a1[2] = {x1[j], T(0)};
#pragma unroll
      for (int v = (j + 1) / VW; v < NV; ++v) {
        T c[VW];
        lds_vec(c, Ls + v * VW + j * NP);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const int i = v * VW + e;
          if (i > j) {
            a0[i & 1] = fma_t(-x0[i], c[e], a0[i & 1]);
            a1[i & 1] = fma_t(-x1[i], c[e], a1[i & 1]);
          }
        }
      }
      x0[j] = (a0[0] + a0[1]) * dinv;
      x1[j] = (a1[0] + a1[1]) * dinv;
    }
  }

  T *Bs = launder(B);  // fresh addresses for the stores (see launder)
  if (!LEFT) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      stg_stream_if(Bs + my0 + (long)j * ldb, x0[j], h0);
      stg_stream_if(Bs + my1 + (long)j * ldb, x1[j], h1);
    }
  } else if (VEC16) {
    constexpr int TSV = TriDualSmem<T, NP, LEFT>::tile_stride16;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      T c0[VW], c1[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        c0[e] = x0[q * VW + e];
        c1[e] = x1[q * VW + e];
      }
      sts_vec(tile + lg * TSV + q * VW, c0);
      sts_vec(tile + (lg + 16) * TSV + q * VW, c1);
    }
    __syncwarp();
    T *Bv = Bs + (long)v0 * ldb;
    const int ncol = !live ? 0 : ((vec - v0 < 32) ? (vec - v0) : 32);
#pragma unroll
    for (int i0 = 0; i0 < 32 * NV; i0 += 16) {
      const int i = i0 + lg, c = i / NV, r0 = (i % NV) * VW;
      if (c < ncol) {
        T o[VW];
        lds_vec(o, tile + c * TSV + r0);
        stg_vec_stream(Bv + (long)c * ldb + r0, o);
      }
    }
  } else {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      tile[lg * TS + j] = x0[j];
      tile[(lg + 16) * TS + j] = x1[j];
    }
    __syncwarp();
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      T t0[16], t1[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        t0[c] = tile[(c0 + c) * TS + lg];
        t1[c] = (NP > 16 && lg + 16 < NP) ? tile[(c0 + c) * TS + lg + 16] : T(0);
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const bool hc = live && (v0 + c0 + c) < vec;
        stg_stream_if(Bs + lg + (long)(v0 + c0 + c) * ldb, t0[c], hc);
        stg_stream_if(Bs + lg + 16 + (long)(v0 + c0 + c) * ldb, t1[c], hc && (lg + 16 < NP));
      }
    }
  }
}

}  // namespace kblasx
