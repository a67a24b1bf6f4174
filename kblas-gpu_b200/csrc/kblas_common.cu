// kblas_common.cu -- handle life cycle, workspace, timer, error strings.
// Counterpart of reference src/kblas_common.cu:35-268 + the method bodies of
// src/kblas_struct.h:93-456.  No cuBLAS on this path: the cuBLAS handle that
// kblasGetCublasHandle() must return is created lazily through dlopen so that
// libkblas-gpu.so has no link-time dependency on libcublas.
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

#include "kblas.h"
#include "kblas_common.h"

// ---------------------------------------------------------------------------------------------
// error reporting (reference src/kblas_common.cu:208-238)
int _kblas_error(cudaError_t err, const char *func, const char *file, int line) {
  if (err != cudaSuccess) {
    fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(err), (int)err, func, file,
            line);
    return 0;
  }
  return 1;
}

int _kblas_error(int err, const char *func, const char *file, int line) {
  if (err != KBLAS_Success) {
    fprintf(stderr, "KBLAS error: %s (%d) in %s at %s:%d\n", kblasx::error_string(err), err, func, file, line);
    return 0;
  }
  return 1;
}

const char *kblasx::error_string(int error) {
  switch (error) {
    case KBLAS_UnknownError: return "KBLAS: unknown error";
    case KBLAS_NotSupported: return "Operation not supported";
    case KBLAS_NotImplemented: return "Operation not implemented yet";
    case KBLAS_cuBLAS_Error: return "cuBLAS error";
    case KBLAS_WrongConfig: return "Wrong compilation flags configuration";
    case KBLAS_CUDA_Error: return "CUDA error";
    case KBLAS_InsufficientWorkspace: return "Insufficient workspace supplied to function";
    case KBLAS_Error_Allocation: return "Error allocating memory";
    case KBLAS_Error_Deallocation: return "Error de-allocating memory";
    case KBLAS_Error_NotInitialized: return "KBLAS handle not initialized";
    case KBLAS_Error_WrongInput: return "One of input parameter's value is wrong";
    case KBLAS_MAGMA_Error: return "MAGMA error";
    case KBLAS_SVD_NoConvergence: return "SVD-gram operation did not converge.";
    default: return "unknown KBLAS error code";
  }
}

// ---------------------------------------------------------------------------------------------
// size helpers (reference src/kblas_common.cu:241-268)
bool REG_SIZE(int n) { return (n > 0) && !(n & (n - 1)); }

int CLOSEST_REG_SIZE(int n) {
  if (n <= 0) return 0;
  int res = 1;
  while (res < n) res <<= 1;
  return res >> 1;
}

extern "C" int kblas_roundup(int x, int y) { return int((x + y - 1) / y) * y; }
long kblas_roundup_l(long x, long y) { return long((x + y - 1) / y) * y; }
size_t kblas_roundup_s(size_t x, size_t y) { return size_t((x + y - 1) / y) * y; }

// ---------------------------------------------------------------------------------------------
// iset_value_{1,2,4,5} (reference src/kblas_common.cu:344-386): output_k[i] = input_k for K arrays in ONE
// grid-stride launch (the reference: grid = batchCount/256 blocks per call, one kernel per arity)
namespace {
template <int K>
struct ISetJob {
  int *out[K];
  int v[K];
};
template <int K>
__global__ void kblasx_iset_value_kernel(ISetJob<K> job, long count) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long step = (long)gridDim.x * blockDim.x;
  for (; i < count; i += step) {
#pragma unroll
    for (int k = 0; k < K; ++k) job.out[k][i] = job.v[k];
  }
}
template <int K>
int iset_launch(const ISetJob<K> &job, long batchCount, cudaStream_t s) {
  if (batchCount <= 0) return KBLAS_Success;
  long blocks = (batchCount + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  kblasx_iset_value_kernel<K><<<(unsigned)blocks, 256, 0, s>>>(job, batchCount);
  check_error_ret(cudaGetLastError(), KBLAS_UnknownError);
  return KBLAS_Success;
}
}  // namespace

int iset_value_1(int *output_array, int input, long batchCount, cudaStream_t cuda_stream) {
  ISetJob<1> j = {{output_array}, {input}};
  return iset_launch(j, batchCount, cuda_stream);
}
int iset_value_2(int *output_array1, int input1, int *output_array2, int input2, long batchCount,
                 cudaStream_t cuda_stream) {
  ISetJob<2> j = {{output_array1, output_array2}, {input1, input2}};
  return iset_launch(j, batchCount, cuda_stream);
}
int iset_value_4(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, long batchCount, cudaStream_t cuda_stream) {
  ISetJob<4> j = {{output_array1, output_array2, output_array3, output_array4}, {input1, input2, input3, input4}};
  return iset_launch(j, batchCount, cuda_stream);
}
int iset_value_5(int *output_array1, int input1, int *output_array2, int input2, int *output_array3, int input3,
                 int *output_array4, int input4, int *output_array5, int input5, long batchCount,
                 cudaStream_t cuda_stream) {
  ISetJob<5> j = {{output_array1, output_array2, output_array3, output_array4, output_array5},
                  {input1, input2, input3, input4, input5}};
  return iset_launch(j, batchCount, cuda_stream);
}

// ---------------------------------------------------------------------------------------------
// workspace (reference src/kblas_struct.h:93-290)
void KBlasWorkspace::reset() {
  allocated = false;
  h_data = NULL;
  h_ptrs = NULL;
  d_data = NULL;
  d_ptrs = NULL;
  allocated_ws_state.reset();
  requested_ws_state.reset();
  consumed_ws_state.reset();
}

KBlasWorkspaceState KBlasWorkspace::getAvailable() const {
  return KBlasWorkspaceState(allocated_ws_state.h_data_bytes - consumed_ws_state.h_data_bytes,
                             allocated_ws_state.h_ptrs_bytes - consumed_ws_state.h_ptrs_bytes,
                             allocated_ws_state.d_data_bytes - consumed_ws_state.d_data_bytes,
                             allocated_ws_state.d_ptrs_bytes - consumed_ws_state.d_ptrs_bytes);
}

#define KBLASX_PUSH_POP(region, field)                                                          \
  void *KBlasWorkspace::push_##region(size_t bytes) {                                           \
    assert(bytes + consumed_ws_state.field <= allocated_ws_state.field);                        \
    void *p = (unsigned char *)region + consumed_ws_state.field;                                \
    consumed_ws_state.field += bytes;                                                           \
    return p;                                                                                   \
  }                                                                                             \
  void KBlasWorkspace::pop_##region(size_t bytes) {                                             \
    assert(consumed_ws_state.field >= bytes);                                                   \
    consumed_ws_state.field -= bytes;                                                           \
  }
KBLASX_PUSH_POP(d_data, d_data_bytes)
KBLASX_PUSH_POP(d_ptrs, d_ptrs_bytes)
KBLASX_PUSH_POP(h_data, h_data_bytes)
KBLASX_PUSH_POP(h_ptrs, h_ptrs_bytes)
#undef KBLASX_PUSH_POP

// grow-only: a region is reallocated only when more was requested than is held
int KBlasWorkspace::allocate() {
  KBlasWorkspaceState &req = requested_ws_state, &have = allocated_ws_state;
  if (req.h_data_bytes > have.h_data_bytes) {
    if (h_data) check_error(cudaFreeHost(h_data));
    h_data = NULL;
    check_error_ret(cudaHostAlloc(&h_data, req.h_data_bytes, cudaHostAllocPortable), KBLAS_Error_Allocation);
    have.h_data_bytes = req.h_data_bytes;
  }
  if (req.h_ptrs_bytes > have.h_ptrs_bytes) {
    if (h_ptrs) check_error(cudaFreeHost(h_ptrs));
    h_ptrs = NULL;
    check_error_ret(cudaHostAlloc((void **)&h_ptrs, req.h_ptrs_bytes, cudaHostAllocPortable),
                    KBLAS_Error_Allocation);
    have.h_ptrs_bytes = req.h_ptrs_bytes;
  }
  if (req.d_data_bytes > have.d_data_bytes) {
    if (d_data) check_error_ret(cudaFree(d_data), KBLAS_Error_Deallocation);
    d_data = NULL;
    check_error_ret(cudaMalloc(&d_data, req.d_data_bytes), KBLAS_Error_Allocation);
    have.d_data_bytes = req.d_data_bytes;
  }
  if (req.d_ptrs_bytes > have.d_ptrs_bytes) {
    if (d_ptrs) check_error_ret(cudaFree(d_ptrs), KBLAS_Error_Deallocation);
    d_ptrs = NULL;
    check_error_ret(cudaMalloc((void **)&d_ptrs, req.d_ptrs_bytes), KBLAS_Error_Allocation);
    have.d_ptrs_bytes = req.d_ptrs_bytes;
  }
  req.reset();
  allocated = true;
  return KBLAS_Success;
}

int KBlasWorkspace::deallocate() {
  if (h_data) cudaFreeHost(h_data);
  if (h_ptrs) cudaFreeHost(h_ptrs);
  if (d_data) check_error_ret(cudaFree(d_data), KBLAS_Error_Deallocation);
  if (d_ptrs) check_error_ret(cudaFree(d_ptrs), KBLAS_Error_Deallocation);
  reset();
  return KBLAS_Success;
}

// ---------------------------------------------------------------------------------------------
// timer (reference src/kblas_gpu_timer.h:24-73)
void kblas_gpu_timer::init() {
  check_error(cudaEventCreate(&start_event));
  check_error(cudaEventCreate(&stop_event));
  elapsed_time = 0;
  recorded_end = false;
}
void kblas_gpu_timer::destroy() {
  check_error(cudaEventDestroy(start_event));
  check_error(cudaEventDestroy(stop_event));
}
void kblas_gpu_timer::start(cudaStream_t s) {
  check_error(cudaEventRecord(start_event, s));
  recorded_end = false;
}
void kblas_gpu_timer::recordEnd(cudaStream_t s) {
  check_error(cudaEventRecord(stop_event, s));
  recorded_end = true;
}
double kblas_gpu_timer::stop(cudaStream_t s) {
  if (!recorded_end) recordEnd(s);
  check_error(cudaEventSynchronize(stop_event));
  check_error(cudaEventElapsedTime(&elapsed_time, start_event, stop_event));
  recorded_end = false;
  return elapsed_time / 1000.0;
}

// ---------------------------------------------------------------------------------------------
// handle (reference src/kblas_struct.h:311-456)
KBlasHandle::KBlasHandle(int /*use_magma*/, cudaStream_t stream_, int device_id_) {
  cublas_handle = NULL;
  create_cublas = 0;  // becomes 1 when kblasGetCublasHandle() creates it
  use_magma = 0;
  device_id = device_id_;
  stream = stream_;
  nStreams = 0;
  timer.init();
  work_space.reset();

  sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device_id);
  smem_optin_max = 227 * 1024;
  cudaDeviceGetAttribute(&smem_optin_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id);
  const char *im = getenv("KBLAS_B200_INFO_MODE");
  info_mode = (im && !strcmp(im, "lapack")) ? KBLASX_INFO_LAPACK : KBLASX_INFO_COMPAT;
  const char *vo = getenv("KBLAS_B200_VARIANT");
  variant_override = vo ? atoi(vo) : -1;
  tri_flags = 0;
  const char *xs = getenv("KBLAS_B200_ELEMENT_EXACT_STORES");
  exact_stores = (xs && atoi(xs) != 0) ? 1 : 0;
  launch_count = 0;
  last_kernel = "none";
  host_pipe = NULL;
  n_kernel_notes = 0;
}

KBlasHandle::KernelNote *KBlasHandle::kernel_note(const void *fn) {
  for (int i = 0; i < n_kernel_notes; ++i)
    if (kernel_notes[i].fn == fn) return &kernel_notes[i];
  // table full (cannot happen with the kernels of this library): recycle the last slot, costing a re-query
  KernelNote *k = &kernel_notes[n_kernel_notes < KERNEL_NOTES ? n_kernel_notes++ : KERNEL_NOTES - 1];
  k->fn = fn;
  k->ctas_per_sm = 0;
  k->smem_limit = 0;
  return k;
}

// minimal lazily-bound cuBLAS (only for kblasGetCublasHandle / kblasSetStream parity)
namespace {
struct LazyCublas {
  void *lib = NULL;
  int (*create)(cublasHandle_t *) = NULL;
  int (*destroy)(cublasHandle_t) = NULL;
  int (*set_stream)(cublasHandle_t, cudaStream_t) = NULL;
  bool load() {
    if (lib) return create != NULL;
    const char *names[] = {"libcublas.so.12", "libcublas.so.13", "libcublas.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) return false;
    create = (int (*)(cublasHandle_t *))dlsym(lib, "cublasCreate_v2");
    destroy = (int (*)(cublasHandle_t))dlsym(lib, "cublasDestroy_v2");
    set_stream = (int (*)(cublasHandle_t, cudaStream_t))dlsym(lib, "cublasSetStream_v2");
    return create && destroy && set_stream;
  }
} g_cublas;
}  // namespace

namespace kblasx { void host_pipe_destroy(void *p); }  // host_pipeline.cu

KBlasHandle::~KBlasHandle() {
  kblasx::host_pipe_destroy(host_pipe);
  if (cublas_handle != NULL && create_cublas && g_cublas.destroy) g_cublas.destroy(cublas_handle);
  for (int i = 0; i < nStreams; ++i) check_error(cudaStreamDestroy(streams[i]));
  timer.destroy();
}

int KBlasHandle::SetStream(cudaStream_t s) {
  stream = s;
  if (cublas_handle && g_cublas.set_stream) g_cublas.set_stream(cublas_handle, s);
  return KBLAS_Success;
}

int KBlasHandle::CreateStreams(int n) {
  if (n > KBLAS_NSTREAMS) return KBLAS_WrongConfig;
  for (int i = 0; i < n; ++i) check_error(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
  nStreams = n;
  return KBLAS_Success;
}

int kblasx::create(KBlasHandle **handle) {
  int dev_id = 0;
  check_error_ret(cudaGetDevice(&dev_id), KBLAS_CUDA_Error);
  *handle = new KBlasHandle(0, 0, dev_id);
  return KBLAS_Success;
}

int kblasx::destroy(KBlasHandle **handle) {
  delete *handle;
  *handle = NULL;
  return KBLAS_Success;
}

void kblasx::set_stream(KBlasHandle *handle, cudaStream_t stream) { handle->SetStream(stream); }

cublasHandle_t kblasx::get_cublas(KBlasHandle *handle) {
  if (handle->cublas_handle == NULL && g_cublas.load()) {
    if (g_cublas.create(&handle->cublas_handle) == 0) {
      handle->create_cublas = 1;
      g_cublas.set_stream(handle->cublas_handle, handle->stream);
    } else {
      handle->cublas_handle = NULL;
    }
  }
  return handle->cublas_handle;
}

// ---------------------------------------------------------------------------------------------
// C++-linkage public API (same mangled names as the reference, src/kblas_common.cu:35-129)
int kblasCreate(kblasHandle_t *handle) { return kblasx::create(handle); }
int kblasDestroy(kblasHandle_t *handle) { return kblasx::destroy(handle); }
int kblasAllocateWorkspace(kblasHandle_t handle) { return handle->work_space.allocate(); }
int kblasFreeWorkspace(kblasHandle_t handle) { return handle->work_space.deallocate(); }
void kblasTimerTic(kblasHandle_t handle) { handle->tic(); }
void kblasTimerRecordEnd(kblasHandle_t handle) { handle->recordEnd(); }
double kblasTimerToc(kblasHandle_t handle) { return handle->toc(); }
int kblasCreateStreams(kblasHandle_t handle, int nStreams) { return handle->CreateStreams(nStreams); }
cudaStream_t kblasGetStream(kblasHandle_t handle) { return handle->stream; }
void kblasSetStream(kblasHandle_t handle, cudaStream_t stream) { kblasx::set_stream(handle, stream); }
cublasHandle_t kblasGetCublasHandle(kblasHandle_t handle) { return kblasx::get_cublas(handle); }
int kblasEnableMagma(kblasHandle_t) {
  printf("ERROR: KBLAS is compiled without magma!\n");
  return KBLAS_Error_NotInitialized;
}
const char *kblasGetErrorString(int error) { return kblasx::error_string(error); }
