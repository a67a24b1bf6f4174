// kernels/trsm_dual.cuh -- side-R batched triangular solves / fused POTRS with a full NP x NP factor
// (NP = 16, 24, 32), TWO right-hand-side vectors per lane and two problems per warp (sm_100a).
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
//   * forward and backward substitution run on the same registers (POTRS = one pass over B).
// Ragged k, side L and vec > 32 handled by slabs here / by the older kernels in the dispatch.
#pragma once

#include "common.cuh"
#include "trsm_small.cuh"  // TriOp, sched_fence

namespace kblasx {

__device__ __forceinline__ void lds_vec(double (&v)[2], const double *p) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void lds_vec(float (&v)[4], const float *p) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}

template <typename T, int NP>
struct TriDualSmem {
  static constexpr int VW = 16 / (int)sizeof(T);
  static constexpr int per_problem = NP * NP + NP + VW;  // factor + reciprocal diagonal + bank-skew pad
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
template <typename T, int NP, int OP, int WARPS, bool STRIDED>
__global__ void __launch_bounds__(WARPS * 32, sizeof(T) == 8 ? KX_DUAL_MINB64 : (NP > 24 ? 3 : 4))
tri_solve_dual_kernel(const int vec, const T alpha, BatchRef<const T, STRIDED> Aref, const int lda, BatchRef<T, STRIDED> Bref,
                      const int ldb, const int batchCount, const int slabs) {
  constexpr int VW = 16 / (int)sizeof(T);
  constexpr int NV = NP / VW;
  constexpr int SE = SectorElems<T>::value;
  constexpr int FSZ = TriDualSmem<T, NP>::per_problem;
  constexpr int FENCE = sizeof(T) == 8 ? KX_DUAL_FENCE64 : KX_DUAL_FENCE32;  // columns between scheduling fences (bounds ptxas' LDS look-ahead)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 4, lg = lane & 15;
  T *Ls = reinterpret_cast<T *>(smem_raw) + (warp * 2 + g) * FSZ;
  T *invd = Ls + NP * NP;

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
  // ---- my two rows of B (vectors v0 + lg and v0 + lg + 16), all loads in flight --------------------------
  T x0[NP], x1[NP];
  const int my0 = v0 + lg, my1 = v0 + lg + 16;
  const bool h0 = live && my0 < vec, h1 = live && my1 < vec;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    x0[j] = T(0);
    x1[j] = T(0);
    ldg_stream_if(x0[j], B + my0 + (long)j * ldb, h0);
    ldg_stream_if(x1[j], B + my1 + (long)j * ldb, h1);
  }
  cp_async_wait_all();
  __syncwarp();
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
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    stg_stream_if(Bs + my0 + (long)j * ldb, x0[j], h0);
    stg_stream_if(Bs + my1 + (long)j * ldb, x1[j], h1);
  }
}

}  // namespace kblasx
