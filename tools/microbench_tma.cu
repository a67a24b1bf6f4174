// tools/microbench_tma.cu -- does a TMA 1-D bulk copy (cp.async.bulk, UBLKCP) fetch DRAM at sector
// or at line granularity?  Each lane copies `bytes` (32/64/128) from the start of its own 128-byte line.
// run under: ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void tma_sectors(const double *__restrict__ buf, long lines, int bytes, int offset_bytes, double *sink) {
  __shared__ __align__(128) unsigned char sm[32 * 128];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + lane * 128);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;
  double acc = 0;
  for (long base = (long)blockIdx.x * 32; base < lines; base += (long)gridDim.x * 32) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32 * bytes) : "memory");
    __syncwarp();
    const char *src = reinterpret_cast<const char *>(buf) + (base + lane) * 128 + offset_bytes;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
    }
    phase ^= 1;
    acc += *reinterpret_cast<double *>(sm + lane * 128);
    __syncwarp();
  }
  if (acc == 123.456) *sink = acc;
}

int main() {
  const long bytes = 1L << 30, lines = bytes / 128;
  double *buf, *sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(buf, 0, bytes));
  CK(cudaDeviceSynchronize());
  tma_sectors<<<148 * 16, 32>>>(buf, lines, 32, 0, sink);    // 1: sector 0
  tma_sectors<<<148 * 16, 32>>>(buf, lines, 32, 96, sink);   // 2: sector 3
  tma_sectors<<<148 * 16, 32>>>(buf, lines, 64, 64, sink);   // 3: sectors 2,3
  tma_sectors<<<148 * 16, 32>>>(buf, lines, 16, 0, sink);    // 4: 16 bytes
  tma_sectors<<<148 * 16, 32>>>(buf, lines, 128, 0, sink);   // 5: whole line
  CK(cudaDeviceSynchronize());
  printf("tma microbench done\n");
  return 0;
}
