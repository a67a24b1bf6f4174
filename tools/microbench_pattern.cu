// tools/microbench_pattern.cu -- memory-system ceiling for the lower-triangle access pattern of
// dpotrf n=32 (lda=32): no arithmetic, just the loads/stores, several write policies.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double ldg_s(const double *p) { double v; asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ void stg_s(double *p, double v) { asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// one warp per matrix, lane = row.  RD/WR: 0 = lower elements only (element predicate),
// 1 = sectors holding a lower element, 2 = 128-byte lines holding a lower element, 3 = everything
template <int RD, int WR, int NC>
__global__ void __launch_bounds__(256) pattern(double *A, long batch) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long m = warp; m < batch; m += nw) {
    double *M = A + m * 1024;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += NC) {
      double v[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = c0 + c;
        const int first = RD == 0 ? col : RD == 1 ? (col & ~3) : RD == 2 ? (col & ~15) : RD == 4 ? (col & ~7) : 0;
        v[c] = 0;
        if (lane >= first) v[c] = ldg_s(M + lane + col * 32);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = c0 + c;
        const int first = WR == 0 ? col : WR == 1 ? (col & ~3) : WR == 2 ? (col & ~15) : WR == 4 ? (col & ~7) : 0;
        if (lane >= first) stg_s(M + lane + col * 32, v[c] + 1.0);
      }
    }
  }
}

template <int RD, int WR, int NC>
void run(const char *name, double *A, long batch, double algo_bytes, double dram_bytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    pattern<RD, WR, NC><<<148 * 8, 256>>>(A, batch);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-46s %7.3f ms  %7.1f Mmat/s  algo %6.0f GB/s (frac %.3f)  est. DRAM %6.0f GB/s\n", name, best, batch / best / 1e3,
         batch * algo_bytes / best / 1e6, batch * algo_bytes / best / 1e6 / 6554.6, batch * dram_bytes / best / 1e6);
}

int main() {
  const long batch = 1 << 20;
  double *A;
  CK(cudaMalloc(&A, batch * 8192));
  CK(cudaMemset(A, 0, batch * 8192));
  // DRAM estimates: reads are line granular (48 lines when only the lower triangle is touched)
  run<1, 0, 8>("rd sectors / wr lower elems (current kernel)", A, batch, 8448, 6144 + 4608);
  run<1, 1, 8>("rd sectors / wr full sectors", A, batch, 8448, 6144 + 4608);
  run<2, 2, 8>("rd lines   / wr full lines", A, batch, 8448, 6144 + 6144);
  run<3, 3, 8>("rd all     / wr all (plain copy in place)", A, batch, 8448, 8192 + 8192);
  run<2, 1, 8>("rd lines   / wr full sectors", A, batch, 8448, 6144 + 4608);
  run<4, 4, 8>("rd 64B pairs / wr 64B pairs", A, batch, 8448, 6144 + 5120);
  run<1, 4, 8>("rd sectors / wr 64B pairs (upper sector garbage)", A, batch, 8448, 6144 + 5120);
  run<4, 4, 16>("rd 64B pairs / wr 64B pairs, 16 cols in flight", A, batch, 8448, 6144 + 5120);
  run<4, 4, 32>("rd 64B pairs / wr 64B pairs, 32 cols in flight", A, batch, 8448, 6144 + 5120);
  run<1, 0, 4>("rd sectors / wr lower elems, 4 cols in flight", A, batch, 8448, 6144 + 4608);
  run<1, 0, 16>("rd sectors / wr lower elems, 16 cols in flight", A, batch, 8448, 6144 + 4608);
  run<1, 1, 16>("rd sectors / wr full sectors, 16 cols in flight", A, batch, 8448, 6144 + 4608);
  run<3, 3, 16>("rd all / wr all, 16 cols in flight", A, batch, 8448, 8192 + 8192);
  return 0;
}
