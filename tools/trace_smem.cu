// tools/trace_smem.cu -- phase timeline of potrf_smem_kernel (developer tool): clock64() stamps of warps 0 and 1 of the first
// CTAs at every phase boundary, for a grid that fills the GPU once (one CTA per SM) and for a full batch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DKX_SMEM_TRACE -Iinclude -Ikblas-gpu_b200/csrc tools/trace_smem.cu -o tools/bin/trace_smem
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kblas.h"
#include "kblas_common.h"
#include "kernels/potrf_smem.cuh"

using namespace kblasx;

template <int W, int MB>
static void run(int n, int batch) {
  const long sA = (long)n * n;
  std::vector<double> h(sA * 64);
  srand(1);
  for (int b = 0; b < 64; ++b)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) h[b * sA + i + (long)j * n] = (i == j) ? n + 1.0 : 0.5 * rand() / RAND_MAX;
  for (int b = 0; b < 64; ++b)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < j; ++i) h[b * sA + i + (long)j * n] = h[b * sA + j + (long)i * n];
  double *dA;
  cudaMalloc(&dA, sizeof(double) * sA * batch);
  for (long b = 0; b < batch; b += 64)
    cudaMemcpy(dA + b * sA, h.data(), sizeof(double) * sA * (batch - b < 64 ? batch - b : 64), cudaMemcpyHostToDevice);
  long long *dT;
  cudaMalloc(&dT, sizeof(long long) * 8 * 128);
  cudaMemset(dT, 0, sizeof(long long) * 8 * 128);
  const SmemPotrfPlan plan = plan_potrf_slots((n + 31) / 32);
  const size_t smem = PotrfSmemGeom::bytes(plan.nslots);
  auto kern = potrf_smem_kernel<W, MB, true>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BatchRef<double, true> ref = {dA, sA};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<batch, W * 32, smem>>>(n, ref, n, batch, nullptr, 0, plan, dT);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> t(8 * 128);
  cudaMemcpy(t.data(), dT, sizeof(long long) * 8 * 128, cudaMemcpyDeviceToHost);
  const int nblk = (n + 31) / 32;
  printf("== n=%d W=%d batch=%d: %s, %.3f ms, smem %zu B\n", n, W, batch, cudaGetErrorString(err), ms, smem);
  for (int w = 0; w < 2; ++w) {
    const long long *q = t.data() + (0 * 2 + w) * 128;
    const long long t0 = t[0];
    printf(" CTA0 warp%d: load %lld  F0 %lld  sync %lld |", w, q[1] - t0, q[2] - q[1], q[3] - q[2]);
    long long prev = q[3];
    for (int J = 0; J < nblk; ++J) {
      const long long *p = q + 8 + 8 * J;
      printf("\n   J=%d: S_own %lld  S_phase_end +%lld", J, p[0] - prev, p[1] - p[0]);
      if (J + 1 < nblk)
        printf("  loadwait %lld  diag_strip %lld  diag_bar +%lld  %s %lld  end_sync +%lld", p[2] - p[1], p[3] - p[2], p[4] - p[3],
               w == 0 ? "F" : "offdiag", p[5] - p[4], p[6] - p[5]);
      prev = p[6];
    }
    printf("\n   total %lld cycles\n", (nblk > 1 ? q[8 + 8 * (nblk - 1) + 1] : q[3]) - t0);
  }
  cudaFree(dA);
  cudaFree(dT);
}

int main() {
  run<8, 1>(256, 148);
  run<8, 1>(256, 8192);
  run<4, 4>(128, 148);
  run<4, 4>(128, 16384);
  run<4, 4>(64, 148);
  run<4, 4>(64, 65536);
  return 0;
}
