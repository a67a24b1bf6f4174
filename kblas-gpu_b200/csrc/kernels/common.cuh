// kernels/common.cuh -- device helpers shared by the sm_100a batch kernels.
#pragma once

#include <cuda_runtime.h>

namespace kblasx {

// which substitution a triangular-solve kernel runs (trsm: one of the two; potrs: both, fused)
enum TriOp { TRI_FORWARD = 0, TRI_BACKWARD = 1, TRI_BOTH = 2 };
// flags (the variants the reference answers KBLAS_NotImplemented for; generic kernels only):
//   TRI_FLAG_UPPER  the factor is stored in the UPPER triangle: what is staged is L = U^T, L[r][c] = A[c + r*lda] (the caller
//                   flips the transposition: op(U) = op'(L)); lanes then read with stride lda -- correct, not fast
//   TRI_FLAG_UNIT   unit diagonal: the diagonal entries are not referenced, 1 / L_jj = 1
enum { TRI_FLAG_UPPER = 1, TRI_FLAG_UNIT = 2 };


template <typename T> struct Vec2T;
template <> struct Vec2T<double> { typedef double2 type; };
template <> struct Vec2T<float>  { typedef float2 type; };

// elements per 32-byte DRAM/L2 sector
template <typename T> struct SectorElems { static constexpr int value = 32 / (int)sizeof(T); };

// ---- where matrix b of a batch lives --------------------------------------------------
// strided: base + b*stride (reference "T* A, long strideA"); pointer array: base[b]
// (reference "T** A": a DEVICE array of device pointers, dereferenced inside the kernel,
//  Xpotrf_batch_kernels.cuh:92-97).  64-bit offsets throughout (8M x 1024 elements).
template <typename T, bool STRIDED> struct BatchRef;
template <typename T> struct BatchRef<T, true> {
  T *base;
  long stride;
  __device__ __forceinline__ T *at(long b) const { return base + b * stride; }
};
template <typename T> struct BatchRef<T, false> {
  T *const *base;
  long stride;  // pointer-array mode: ELEMENT OFFSET added to every entry (the reference's A_row_off + A_col_off*lda
                // of Xpotrf_batch_offset & co., Xpotrf_batch.cu:44-48; its drivers launch pointer fix-up kernels instead)
  __device__ __forceinline__ T *at(long b) const { return base[b] + stride; }
};

// ---- streaming global access ----------------------------------------------------------
// Every byte of a matrix is touched exactly once per kernel, so bypass L1 allocation.
__device__ __forceinline__ double ldg_stream(const double *p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float *p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// predicated forms: a real predicated LDG (no branch, no convergence barrier), so that a whole
// batch of loads can be in flight before the first use; `v` keeps its value when pred is false.
__device__ __forceinline__ void ldg_stream_if(double &v, const double *p, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.L1::no_allocate.f64 %0, [%1]; }"
               : "+d"(v) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg_stream_if(float &v, const float *p, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.L1::no_allocate.f32 %0, [%1]; }"
               : "+f"(v) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void stg_stream(double *p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(float *p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ void stg_stream_if(double *p, double v, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.global.L1::no_allocate.f64 [%0], %1; }"
               ::"l"(p), "d"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void stg_stream_if(float *p, float v, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.global.L1::no_allocate.f32 [%0], %1; }"
               ::"l"(p), "f"(v), "r"((int)pred) : "memory");
}

// Hide a pointer's provenance from the optimiser.  Without it NVVM keeps every address it has computed
// alive for a later re-use (the 64 column addresses of a forward pass for the backward pass of POTRS, load
// addresses for the stores): hundreds of registers, or kilobytes of spills under a cap.
template <typename T>
__device__ __forceinline__ T *launder(T *p) {
  asm volatile("" : "+l"(p));
  return p;
}

// ---- shared-memory pair loads -----------------------------------------------------------
// asm volatile on purpose: the staged factor is immutable, and a plain (const __restrict__)
// load lets the compiler keep every value it has ever read alive in registers across the
// forward and the backward substitution (observed: 15 KB of local-memory spills).
__device__ __forceinline__ double2 lds_pair(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ float2 lds_pair(const float *p) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ double lds_one(const double *p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ float lds_one(const float *p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

// ---- arithmetic -----------------------------------------------------------------------
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }
__device__ __forceinline__ float sqrt_t(float a) { return sqrtf(a); }
__device__ __forceinline__ double rsqrt_t(double a) { return rsqrt(a); }
__device__ __forceinline__ float rsqrt_t(float a) { return rsqrtf(a); }

// c = fma(a, b, c) / c = c * a under a predicate, as ONE predicated instruction (a C++ `if`
// becomes DFMA + 2 FSEL: +20 % instructions in the n=32 Cholesky kernel)
__device__ __forceinline__ void fma_if(double &c, double a, double b, bool pred) {
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q fma.rn.f64 %0, %1, %2, %0; }" : "+d"(c) : "d"(a), "d"(b), "r"((int)pred));
}
__device__ __forceinline__ void fma_if(float &c, float a, float b, bool pred) {
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q fma.rn.f32 %0, %1, %2, %0; }" : "+f"(c) : "f"(a), "f"(b), "r"((int)pred));
}
__device__ __forceinline__ void mul_if(double &c, double a, bool pred) {
  asm("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q mul.rn.f64 %0, %0, %1; }" : "+d"(c) : "d"(a), "r"((int)pred));
}
__device__ __forceinline__ void mul_if(float &c, float a, bool pred) {
  asm("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q mul.rn.f32 %0, %0, %1; }" : "+f"(c) : "f"(a), "r"((int)pred));
}

// 8-/4-byte asynchronous global -> shared copy (LDGSTS) and its completion wait
__device__ __forceinline__ void cp_async_elem(double *smem_dst, const double *gsrc, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q cp.async.ca.shared.global [%0], [%1], 8; }"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float *smem_dst, const float *gsrc, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q cp.async.ca.shared.global [%0], [%1], 4; }"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// width-G broadcast from lane `src` of each G-lane segment
template <int G>
__device__ __forceinline__ double shfl_seg(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src, G);
}
template <int G>
__device__ __forceinline__ float shfl_seg(float v, int src) {
  return __shfl_sync(0xffffffffu, v, src, G);
}

}  // namespace kblasx
