// kernels/vec16.cuh -- 16-byte access helpers shared by the solve kernels: cp.async (LDGSTS.128, zero-filled when predicated
// off), 128-bit shared-memory loads / stores and 128-bit streaming global stores of 2 doubles / 4 floats.
#pragma once

#include "common.cuh"

namespace kblasx {

__device__ __forceinline__ void lds_vec(double (&v)[2], const double *p) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void lds_vec(float (&v)[4], const float *p) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"((unsigned)__cvta_generic_to_shared(p)));
}

__device__ __forceinline__ void cp_async16_if(void *smem_dst, const void *gsrc, bool pred) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(pred ? 16 : 0) : "memory");
}
__device__ __forceinline__ void sts_vec(double *p, const double (&v)[2]) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void sts_vec(float *p, const float (&v)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void stg_vec_stream(double *p, const double (&v)[2]) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void stg_vec_stream(float *p, const float (&v)[4]) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
}

}  // namespace kblasx
