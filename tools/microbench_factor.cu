// tools/microbench_factor.cu -- how long does one warp need for the 32 x 32 diagonal-block Cholesky (lane = row, the
// "F" step of the n > 32 kernels), and which part of the per-pivot dependency chain costs what?
//   V0  shfl pivot -> rsqrt() -> scale -> STS column -> LDS broadcast -> trailing FMAs         (the round-1 code)
//   V1  V0 with a branch-free rsqrt (MUFU.RSQ64H + one third-order step, no slow-path call)
//   V2  V1 + pivot look-ahead: pivot j+1 = p[j+1] - p[j]^2 is formed in lane j+1 from its own registers and shuffled
//       before column j goes through shared memory (STS/LDS leave the chain)
//   V3  V2 with the column broadcast by SHFL instead of shared memory
//   V4  2 x 2 pivot blocks: rsqrt(a) and rsqrt(a c - b^2) are independent, so one rsqrt latency covers two pivots
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench_factor tools/microbench_factor.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double rsqrt_fast(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y * y, 1.0);
  return fma(fma(0.375, e, 0.5), e * y, y);
}
__device__ __forceinline__ double lds_one(const double *p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

__device__ __forceinline__ double2 lds_pair(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
// the library's fast path, with its special cases inline instead of behind a call
__device__ __forceinline__ double rsqrt_inl(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const unsigned hi = (unsigned)__double2hiint(a);
  if (__builtin_expect(hi - 0x00100000u >= 0x7fe00000u, 0)) {  // zero, subnormal, negative, inf, nan
    if (!(a > 0.0 && a < __longlong_as_double(0x7ff0000000000000ll))) return y;  // MUFU already returns inf / nan / 0
    a *= 0x1p108;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y * y, 1.0);
    return fma(fma(0.375, e, 0.5), e * y, y) * 0x1p54;
  }
  const double e = fma(-a, y * y, 1.0);
  return fma(fma(0.375, e, 0.5), e * y, y);
}

template <int V>
__device__ __forceinline__ void factor(double (&p)[32], double *Lkk, double *invd, const int lane) {
  if (V == 0 || V == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double d = __shfl_sync(0xffffffffu, p[j], j);
      const double r = V == 0 ? rsqrt(d) : rsqrt_fast(d);
      p[j] *= r;
      Lkk[lane + j * 32] = p[j];
      if (lane == j) invd[j] = r;
      __syncwarp();
#pragma unroll
      for (int k = j + 1; k < 32; ++k) p[k] = fma(-p[j], lds_one(Lkk + k + j * 32), p[k]);
    }
  } else if (V == 2) {
    double d = __shfl_sync(0xffffffffu, p[0], 0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double r = rsqrt_fast(d);
      p[j] *= r;
      if (j + 1 < 32) d = __shfl_sync(0xffffffffu, fma(-p[j], p[j], p[j + 1]), j + 1);
      Lkk[lane + j * 32] = p[j];
      if (lane == j) invd[j] = r;
      __syncwarp();
#pragma unroll
      for (int k = j + 1; k < 32; ++k) p[k] = fma(-p[j], lds_one(Lkk + k + j * 32), p[k]);
    }
  } else if (V == 5 || V == 6) {
    double d = __shfl_sync(0xffffffffu, p[0], 0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double r = V == 5 ? rsqrt_fast(d) : rsqrt_inl(d);
      p[j] *= r;
      if (j + 1 < 32) d = __shfl_sync(0xffffffffu, fma(-p[j], p[j], p[j + 1]), j + 1);
      Lkk[lane + j * 32] = p[j];
      if (lane == j) invd[j] = r;
      __syncwarp();
#pragma unroll
      for (int k = (j + 1) & ~1; k < 32; k += 2) {
        const double2 l2 = lds_pair(Lkk + k + j * 32);
        if (k > j) p[k] = fma(-p[j], l2.x, p[k]);
        p[k + 1] = fma(-p[j], l2.y, p[k + 1]);
      }
    }
  } else if (V == 3) {
    double d = __shfl_sync(0xffffffffu, p[0], 0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double r = rsqrt_fast(d);
      p[j] *= r;
      if (j + 1 < 32) d = __shfl_sync(0xffffffffu, fma(-p[j], p[j], p[j + 1]), j + 1);
      if (lane == j) invd[j] = r;
#pragma unroll
      for (int k = j + 1; k < 32; ++k) p[k] = fma(-p[j], __shfl_sync(0xffffffffu, p[j], k), p[k]);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) Lkk[lane + j * 32] = p[j];
    __syncwarp();
  } else if (V == 4) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const double a = __shfl_sync(0xffffffffu, p[j], j);
      const double b = __shfl_sync(0xffffffffu, p[j], j + 1);
      const double c = __shfl_sync(0xffffffffu, p[j + 1], j + 1);
      const double ra = rsqrt_fast(a);
      const double bb = b * b;
      const double det = fma(a, c, -bb) - fma(b, b, -bb);  // a c - b^2 with the rounding of b^2 compensated
      const double rdet = rsqrt_fast(det);
      const double l21 = b * ra;
      const double inv22 = rdet * (a * ra);  // 1 / sqrt(c - b^2 / a)
      p[j] *= ra;
      p[j + 1] = fma(-p[j], l21, p[j + 1]) * inv22;
      Lkk[lane + j * 32] = p[j];
      Lkk[lane + (j + 1) * 32] = p[j + 1];
      if (lane == j) invd[j] = ra;
      if (lane == j + 1) invd[j + 1] = inv22;
      __syncwarp();
#pragma unroll
      for (int k = j + 2; k < 32; ++k) {
        p[k] = fma(-p[j], lds_one(Lkk + k + j * 32), p[k]);
        p[k] = fma(-p[j + 1], lds_one(Lkk + k + (j + 1) * 32), p[k]);
      }
    }
  }
}

template <int V>
__global__ void __launch_bounds__(512) bench(const double *A0, double *out, long long *cycles, int reps) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *Lkk = sm + warp * (2 * 32 * 32 + 32), *invd = Lkk + 32 * 32, *P0 = invd + 32;
  double p[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { P0[lane + c * 32] = A0[lane + c * 32]; p[c] = 0.0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
#pragma unroll
    for (int c = 0; c < 32; ++c) p[c] = fma(1e-30, p[c], P0[lane + c * 32]);  // depends on the previous round: no overlap between factorizations
    factor<V>(p, Lkk, invd, lane);
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  if (blockIdx.x == 0 && warp == 0) {
#pragma unroll
    for (int c = 0; c < 32; ++c) out[lane + c * 32] = c <= lane ? p[c] : 0.0;
  }
}

template <int V>
void run(const char *name, const double *A0, double *out, long long *cyc, const double *Lref_h) {
  const int reps = 40;
  static double Lh[1024];
  printf("%-34s", name);
  for (int warps : {1, 4, 8, 12}) {
    const int smem = warps * (2 * 32 * 32 + 32) * 8;
    CK(cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    bench<V><<<148, warps * 32, smem>>>(A0, out, cyc, reps);
    bench<V><<<148, warps * 32, smem>>>(A0, out, cyc, reps);
    CK(cudaDeviceSynchronize());
    static long long h[148 * 16];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost));
    double s = 0;
    for (int i = 0; i < 148 * warps; ++i) s += (double)h[i];
    printf("  %2d warps/SM: %6.0f cyc/F (%5.0f per F per SM)", warps, s / (148.0 * warps) / reps, s / (148.0 * warps) / reps / warps);
  }
  CK(cudaMemcpy(Lh, out, sizeof(Lh), cudaMemcpyDeviceToHost));
  double err = 0;
  for (int i = 0; i < 1024; ++i) err = fmax(err, fabs(Lh[i] - Lref_h[i]));
  printf("  max|L - Lref| %.2e\n", err);
}

int main() {
  static double Ah[1024], Lr[1024];
  srand(7);
  static double G[1024];
  for (int i = 0; i < 1024; ++i) G[i] = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      double s = i == j ? 32.0 : 0.0;
      for (int k = 0; k < 32; ++k) s += G[i + 32 * k] * G[j + 32 * k];
      Ah[i + 32 * j] = s;
    }
  for (int i = 0; i < 1024; ++i) Lr[i] = 0;
  {  // host Cholesky
    static double W[1024];
    for (int i = 0; i < 1024; ++i) W[i] = Ah[i];
    for (int j = 0; j < 32; ++j) {
      const double d = sqrt(W[j + 32 * j]);
      for (int i = j; i < 32; ++i) W[i + 32 * j] = i == j ? d : W[i + 32 * j] / d;
      for (int k = j + 1; k < 32; ++k)
        for (int i = k; i < 32; ++i) W[i + 32 * k] -= W[i + 32 * j] * W[k + 32 * j];
    }
    for (int j = 0; j < 32; ++j)
      for (int i = j; i < 32; ++i) Lr[i + 32 * j] = W[i + 32 * j];
  }
  double *A0, *out;
  long long *cyc;
  CK(cudaMalloc(&A0, sizeof(Ah)));
  CK(cudaMalloc(&out, sizeof(Ah)));
  CK(cudaMalloc(&cyc, sizeof(long long) * 148 * 16));
  CK(cudaMemcpy(A0, Ah, sizeof(Ah), cudaMemcpyHostToDevice));
  run<0>("V0 shfl, rsqrt(), STS/LDS", A0, out, cyc, Lr);
  run<1>("V1 branch-free rsqrt", A0, out, cyc, Lr);
  run<2>("V2 + pivot look-ahead", A0, out, cyc, Lr);
  run<5>("V5 V2 with LDS.128", A0, out, cyc, Lr);
  run<6>("V6 V5, inline special cases", A0, out, cyc, Lr);
  run<3>("V3 V2 with SHFL broadcast", A0, out, cyc, Lr);
  run<4>("V4 2x2 pivot blocks", A0, out, cyc, Lr);
  return 0;
}
