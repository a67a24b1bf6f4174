// tools/microbench.cu -- B200 micro-measurements that decide the kernel design (developer tool).
//   part A (run under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum): DRAM fetch / fill
//           granularity for partial-line reads and partial-sector writes;
//   part B (plain run): FP64 pipe latency / throughput, MUFU.RSQ64H, SHFL, broadcast LDS.128, DMMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- part A
// every thread reads `nsec` consecutive 32-byte sectors starting at sector `first` of its own 128-byte line
template <int MODE>
__global__ void read_sectors(const double *__restrict__ buf, long lines, int first, int nsec, double *sink) {
  long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0;
  for (; line < lines; line += (long)gridDim.x * blockDim.x) {
    const double *p = buf + line * 16 + first * 4;
    for (int s = 0; s < nsec; ++s) {
      double v;
      if (MODE == 0) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p + 4 * s));
      if (MODE == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p + 4 * s));
      if (MODE == 2) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p + 4 * s));
      if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p + 4 * s));
      if (MODE == 4) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p + 4 * s));
      acc += v;
    }
  }
  if (acc == 123.456) *sink = acc;
}

// warp-coalesced variant: a warp reads rows [first_row, 32) of consecutive 256-byte columns (the potrf pattern)
__global__ void read_cols(const double *__restrict__ buf, long cols, int first_row, double *sink) {
  long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  double acc = 0;
  for (long c = warp; c < cols; c += nwarps) {
    if (lane >= first_row) {
      double v;
      asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(buf + c * 32 + lane));
      acc += v;
    }
  }
  if (acc == 123.456) *sink = acc;
}

// every thread writes `nbytes8` doubles (8 B each) at the start of sector `sec` of its own 128-byte line
__global__ void write_partial(double *buf, long lines, int sec, int ndoubles) {
  long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; line < lines; line += (long)gridDim.x * blockDim.x) {
    double *p = buf + line * 16 + sec * 4;
    for (int i = 0; i < ndoubles; ++i) asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p + i), "d"(1.0 + i) : "memory");
  }
}

// ---------------------------------------------------------------- part B
__global__ void k_dfma_lat(double *out, int iters, long long *cyc) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) c = fma(a, b, c);
  }
  long long t1 = clock64();
  out[3] = c;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_dfma_tput(double *out, int iters, long long *cyc) {
  double a = out[0], b = out[1];
  double c[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) c[u] = out[2] + u;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) c[u] = fma(a, b, c[u]);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += c[u];
  out[3 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_rsqrt_lat(double *out, int iters, long long *cyc, int mode) {
  double x = out[0] + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) x = rsqrt(x) + 1.5;
    else if (mode == 1) x = 1.0 / sqrt(x) + 1.5;
    else if (mode == 2) x = sqrt(x) + 1.5;
    else x = 1.0 / x + 1.5;
  }
  long long t1 = clock64();
  out[3] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(double *out, int iters, long long *cyc, int dep) {
  double x = out[0] + threadIdx.x, y = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (dep) {
#pragma unroll
      for (int u = 0; u < 8; ++u) x = __shfl_sync(0xffffffffu, x, (u + 1) & 7, 8);
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) y += __shfl_sync(0xffffffffu, x, (u + i) & 7, 8);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  out[3 + threadIdx.x] = x + y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds_bcast(double *out, int iters, long long *cyc, int dep) {
  __shared__ __align__(16) double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7) & 1022);
  __syncthreads();
  int g = (threadIdx.x & 31) >> 3;
  double acc0 = 0, acc1 = 0;
  int idx = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double2 v = *reinterpret_cast<const double2 *>(&sm[((idx + u * 8 + g * 2) & 1022)]);
      acc0 += v.x;
      acc1 += v.y;
      if (dep) idx = (int)v.x;
    }
    if (!dep) idx += 64;
  }
  long long t1 = clock64();
  out[3 + threadIdx.x] = acc0 + acc1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dmma(double *out, int iters, long long *cyc, int ilp) {
  double a = out[0] + threadIdx.x, b = out[1] - threadIdx.x;
  double c[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u < ilp)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
  out[3 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// legacy tensor path, TF32 m16n8k8 (fp32 accumulate): rate and latency (3 of these emulate one fp32 product)
__global__ void k_tf32mma(float *out, int iters, long long *cyc, int ilp) {
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(out[i] + threadIdx.x) & 0xffffe000u;
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(out[i] - threadIdx.x) & 0xffffe000u;
  float c[8][4];
#pragma unroll
  for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u < ilp)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
  out[8 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main(int argc, char **argv) {
  const char *part = argc > 1 ? argv[1] : "B";
  int gran = argc > 2 ? atoi(argv[2]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("set L2 fetch granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), g);
  } else {
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("default L2 fetch granularity %zu\n", g);
  }
  double *sink;
  CK(cudaMalloc(&sink, 8));
  if (part[0] == 'A') {
    const long bytes = 1L << 30;  // 1 GiB >> L2
    const long lines = bytes / 128;
    double *buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    CK(cudaDeviceSynchronize());
    // order of launches (ncu lists them in this order)
    read_sectors<0><<<148 * 8, 256>>>(buf, lines, 0, 1, sink);   // 1: 1 of 4 sectors, no_allocate
    read_sectors<0><<<148 * 8, 256>>>(buf, lines, 1, 1, sink);   // 2: sector 1 only
    read_sectors<0><<<148 * 8, 256>>>(buf, lines, 0, 2, sink);   // 3: sectors 0,1
    read_sectors<0><<<148 * 8, 256>>>(buf, lines, 1, 3, sink);   // 4: sectors 1,2,3
    read_sectors<0><<<148 * 8, 256>>>(buf, lines, 0, 4, sink);   // 5: all 4
    read_sectors<1><<<148 * 8, 256>>>(buf, lines, 0, 1, sink);   // 6: .cg 1 of 4
    read_sectors<2><<<148 * 8, 256>>>(buf, lines, 0, 1, sink);   // 7: .cs 1 of 4
    read_sectors<3><<<148 * 8, 256>>>(buf, lines, 0, 1, sink);   // 8: nc evict_first 1 of 4
    read_sectors<4><<<148 * 8, 256>>>(buf, lines, 0, 1, sink);   // 9: plain 1 of 4
    read_cols<<<148 * 8, 256>>>(buf, bytes / 256, 28, sink);     // 10: rows 28..31 of each column (1 of 8 sectors)
    read_cols<<<148 * 8, 256>>>(buf, bytes / 256, 16, sink);     // 11: rows 16..31 (line B only)
    read_cols<<<148 * 8, 256>>>(buf, bytes / 256, 12, sink);     // 12: rows 12..31 (1 sector of line A + line B)
    read_cols<<<148 * 8, 256>>>(buf, bytes / 256, 4, sink);      // 13: rows 4..31
    write_partial<<<148 * 8, 256>>>(buf, lines, 0, 4);           // 14: full sector 0 of each line
    write_partial<<<148 * 8, 256>>>(buf, lines, 1, 1);           // 15: 8 bytes of sector 1 of each line
    write_partial<<<148 * 8, 256>>>(buf, lines, 2, 3);           // 16: 24 bytes of sector 2
    CK(cudaDeviceSynchronize());
    printf("part A done: buffer %ld bytes, %ld lines\n", bytes, lines);
    return 0;
  }
  // ---- part B
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  double *out;
  long long *cyc, hc[4];
  CK(cudaMalloc(&out, 8 * 2048));
  CK(cudaMemset(out, 0, 8 * 2048));
  CK(cudaMalloc(&cyc, 8 * 1024));
  const int it = 2000;
  k_dfma_lat<<<1, 32>>>(out, it, cyc);
  CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
  printf("DFMA dependent-chain latency: %.2f cycles\n", (double)hc[0] / (it * 16.0));
#define TPUT(ILP, THREADS)                                                                        \
  k_dfma_tput<ILP><<<1, THREADS>>>(out, it, cyc);                                                 \
  CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));                                             \
  printf("DFMA ILP=%d warps/SM=%d: %.2f cycles per warp-instr per SMSP-equivalent, %.1f lane-FMA/clk/SM\n", ILP, THREADS / 32, \
         (double)hc[0] / (it * (double)ILP) / ((THREADS / 32 + 3) / 4), (double)it * ILP * THREADS / hc[0]);
  TPUT(1, 32) TPUT(4, 32) TPUT(8, 32) TPUT(16, 32) TPUT(8, 128) TPUT(8, 256) TPUT(8, 512) TPUT(16, 1024)
  const char *nm[4] = {"rsqrt(x)", "1/sqrt(x)", "sqrt(x)", "1/x"};
  for (int m = 0; m < 4; ++m) {
    k_rsqrt_lat<<<1, 32>>>(out, it, cyc, m);
    CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%s (+1 DADD) dependent latency: %.1f cycles\n", nm[m], (double)hc[0] / it);
  }
  for (int dep = 1; dep >= 0; --dep)
    for (int th = 32; th <= 1024; th *= 4) {
      k_shfl<<<1, th>>>(out, it, cyc, dep);
      CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
      printf("SHFL fp64 (2x SHFL.IDX) %s, %d warps: %.2f cycles per fp64 shuffle per warp, %.3f fp64-shfl/clk/SM\n",
             dep ? "dependent" : "independent", th / 32, (double)hc[0] / (it * 8.0), it * 8.0 * (th / 32) / hc[0]);
    }
  for (int dep = 1; dep >= 0; --dep)
    for (int th = 32; th <= 1024; th *= 4) {
      k_lds_bcast<<<1, th>>>(out, it, cyc, dep);
      CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
      printf("LDS.128 4-address broadcast %s, %d warps: %.2f cycles per LDS per warp, %.3f LDS/clk/SM\n",
             dep ? "dependent" : "independent", th / 32, (double)hc[0] / (it * 8.0), it * 8.0 * (th / 32) / hc[0]);
    }
  for (int ilp = 1; ilp <= 8; ilp *= 2)
    for (int th = 32; th <= 1024; th *= 4) {
      k_dmma<<<1, th>>>(out, it, cyc, ilp);
      CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
      printf("DMMA m8n8k4 ilp=%d, %d warps: %.2f cycles per DMMA per warp, %.1f FMA/clk/SM\n", ilp, th / 32,
             (double)hc[0] / (it * (double)ilp), it * (double)ilp * (th / 32) * 256.0 / hc[0]);
    }
  for (int ilp = 1; ilp <= 8; ilp *= 2)
    for (int th = 32; th <= 1024; th *= 4) {
      k_tf32mma<<<1, th>>>(reinterpret_cast<float *>(out), it, cyc, ilp);
      CK(cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost));
      printf("TF32 mma.sync m16n8k8 ilp=%d, %d warps: %.2f cycles per MMA per warp, %.1f MAC/clk/SM\n", ilp, th / 32,
             (double)hc[0] / (it * (double)ilp), it * (double)ilp * (th / 32) * 1024.0 / hc[0]);
    }
  printf("SM clock %d kHz, SMs %d\n", prop.clockRate, prop.multiProcessorCount);
  return 0;
}
