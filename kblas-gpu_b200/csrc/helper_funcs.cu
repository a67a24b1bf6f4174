// helper_funcs.cu -- device pointer-array builders.
//
// Counterpart of reference src/batch_triangular/Xhelper_funcs.cu:74-105 and the kernels
// in Xhelper_funcs.cuh:280-305.  The reference launches grid = batchCount blocks of ONE
// thread each (Xhelper_funcs.cuh:299); here a single grid-stride kernel with 256-thread
// blocks, capped at a few waves of the 148 SMs, writes all (up to three) arrays.
// None of this library's compute kernels need these arrays -- they exist because callers
// (and the reference's test binaries) use them to build their own pointer-array arguments.
#include "kblas.h"
#include "kblas_common.h"

namespace {

template <typename T, int K>
struct PtrJob {
  T **out[K];
  const T *in[K];
  long off[K];
};

template <typename T, int K>
__global__ void set_pointer_kernel(PtrJob<T, K> job, long count) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long step = (long)gridDim.x * blockDim.x;
  for (; i < count; i += step) {
#pragma unroll
    for (int k = 0; k < K; ++k) job.out[k][i] = const_cast<T *>(job.in[k]) + i * job.off[k];
  }
}

// pointer-array flavour (reference Xhelper_funcs.cuh:140-233): out_k[i] = in_k[i] + row_off_k + col_off_k * ld_k,
// ld either one value per array (uniform) or one value per matrix (ld_k[i], the non-uniform overload of _1)
template <typename T, int K>
struct PtrOffJob {
  T **out[K];
  const T *const *in[K];
  long off[K];        // uniform: row_off + col_off * ld
  int row_off[K], col_off[K];
  const int *ld[K];   // per-matrix leading dimensions (NULL = uniform, `off` is final)
};

template <typename T, int K>
__global__ void set_pointer_off_kernel(PtrOffJob<T, K> job, long count) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long step = (long)gridDim.x * blockDim.x;
  for (; i < count; i += step) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long o = job.ld[k] ? (long)job.row_off[k] + (long)job.col_off[k] * job.ld[k][i] : job.off[k];
      job.out[k][i] = const_cast<T *>(job.in[k][i]) + o;
    }
  }
}

template <typename T, int K>
int launch_off(const PtrOffJob<T, K> &job, long count, cudaStream_t s) {
  if (count <= 0) return KBLAS_Success;
  long blocks = (count + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  set_pointer_off_kernel<T, K><<<(unsigned)blocks, 256, 0, s>>>(job, count);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

template <typename T, int K>
int launch(const PtrJob<T, K> &job, long count, cudaStream_t s) {
  if (count <= 0) return KBLAS_Success;
  long blocks = (count + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  set_pointer_kernel<T, K><<<(unsigned)blocks, 256, 0, s>>>(job, count);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}

}  // namespace

namespace kblasx {

template <typename T>
int set_pointer_1(T **out, const T *in, int /*lda*/, long batch_offset, long batchCount, cudaStream_t s) {
  PtrJob<T, 1> j = {{out}, {in}, {batch_offset}};
  return launch(j, batchCount, s);
}
template <typename T>
int set_pointer_2(T **out1, const T *in1, int, long off1, T **out2, const T *in2, int, long off2, long batchCount,
                  cudaStream_t s) {
  PtrJob<T, 2> j = {{out1, out2}, {in1, in2}, {off1, off2}};
  return launch(j, batchCount, s);
}
template <typename T>
int set_pointer_3(T **out1, const T *in1, int, long off1, T **out2, const T *in2, int, long off2, T **out3,
                  const T *in3, int, long off3, long batchCount, cudaStream_t s) {
  PtrJob<T, 3> j = {{out1, out2, out3}, {in1, in2, in3}, {off1, off2, off3}};
  return launch(j, batchCount, s);
}

#define INST(T)                                                                                            \
  template int set_pointer_1<T>(T **, const T *, int, long, long, cudaStream_t);                            \
  template int set_pointer_2<T>(T **, const T *, int, long, T **, const T *, int, long, long, cudaStream_t); \
  template int set_pointer_3<T>(T **, const T *, int, long, T **, const T *, int, long, T **, const T *, int, \
                                long, long, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace kblasx

// ---- C++-linkage names the reference's test binaries call (src/Xhelper_funcs.ch:48-60)
#define PUB(T)                                                                                             \
  int Xset_pointer_1(T **output_array, const T *input, int lda, long batch_offset, long batchCount,         \
                     cudaStream_t cuda_stream) {                                                            \
    return kblasx::set_pointer_1<T>(output_array, input, lda, batch_offset, batchCount, cuda_stream);       \
  }                                                                                                        \
  int Xset_pointer_2(T **output_array1, const T *input1, int ld1, long batch_offset1, T **output_array2,    \
                     const T *input2, int ld2, long batch_offset2, long batchCount,                         \
                     cudaStream_t cuda_stream) {                                                            \
    return kblasx::set_pointer_2<T>(output_array1, input1, ld1, batch_offset1, output_array2, input2, ld2,  \
                                    batch_offset2, batchCount, cuda_stream);                                \
  }                                                                                                        \
  int Xset_pointer_3(T **output_array1, const T *input1, int ld1, long batch_offset1, T **output_array2,    \
                     const T *input2, int ld2, long batch_offset2, T **output_array3, const T *input3,      \
                     int ld3, long batch_offset3, long batchCount, cudaStream_t cuda_stream) {              \
    return kblasx::set_pointer_3<T>(output_array1, input1, ld1, batch_offset1, output_array2, input2, ld2,  \
                                    batch_offset2, output_array3, input3, ld3, batch_offset3, batchCount,   \
                                    cuda_stream);                                                           \
  }
PUB(float)
PUB(double)
#undef PUB

// pointer-array inputs: sub-matrix (offset_r, offset_c) of every entry (src/Xhelper_funcs.ch:32-46)
#define PUB_OFF(T)                                                                                          \
  int Xset_pointer_1(T **output_array, T **input, int offset_r, int offset_c, int *lda, long batchCount,     \
                     cudaStream_t cuda_stream) {                                                            \
    PtrOffJob<T, 1> j = {{output_array}, {input}, {0}, {offset_r}, {offset_c}, {lda}};                      \
    return launch_off(j, batchCount, cuda_stream);                                                          \
  }                                                                                                         \
  int Xset_pointer_2(T **output_array1, const T **input1, int offset_r1, int offset_c1, int lda1,           \
                     T **output_array2, const T **input2, int offset_r2, int offset_c2, int lda2,           \
                     long batchCount, cudaStream_t cuda_stream) {                                           \
    PtrOffJob<T, 2> j = {{output_array1, output_array2}, {input1, input2},                                  \
                         {offset_r1 + (long)offset_c1 * lda1, offset_r2 + (long)offset_c2 * lda2},          \
                         {0, 0}, {0, 0}, {nullptr, nullptr}};                                               \
    return launch_off(j, batchCount, cuda_stream);                                                          \
  }                                                                                                         \
  int Xset_pointer_3(T **output_array1, const T **input1, int offset_r1, int offset_c1, int lda1,           \
                     T **output_array2, const T **input2, int offset_r2, int offset_c2, int lda2,           \
                     T **output_array3, const T **input3, int offset_r3, int offset_c3, int lda3,           \
                     long batchCount, cudaStream_t cuda_stream) {                                           \
    PtrOffJob<T, 3> j = {{output_array1, output_array2, output_array3}, {input1, input2, input3},           \
                         {offset_r1 + (long)offset_c1 * lda1, offset_r2 + (long)offset_c2 * lda2,           \
                          offset_r3 + (long)offset_c3 * lda3},                                              \
                         {0, 0, 0}, {0, 0, 0}, {nullptr, nullptr, nullptr}};                                \
    return launch_off(j, batchCount, cuda_stream);                                                          \
  }
PUB_OFF(float)
PUB_OFF(double)
#undef PUB_OFF
