// kblas_struct.h -- handle + workspace plumbing of the B200-native KBLAS batch path.
//
// Replaces reference src/kblas_struct.h:43-456.  Field names that the reference's
// own test binaries reach into (they add src/ to their include path) are kept
// source-compatible: handle->stream, ->cublas_handle, ->use_magma, ->device_id,
// ->work_space.{allocated,requested,consumed}_ws_state.{h_data,h_ptrs,d_data,d_ptrs}_bytes.
// Everything else is new: the kernels on this path need no scratch memory, so the
// workspace is pure bookkeeping kept for API parity (wsquery -> allocate -> call,
// KBLAS_InsufficientWorkspace when skipped), cuBLAS is created lazily, and the handle
// carries launch accounting plus the tuning knobs of the sm_100a kernels.
#pragma once

#include <cstddef>
#include <cuda_runtime.h>
#include "kblas_defs.h"

// Same opaque types as include/kblas.h.  That header is deliberately NOT included here:
// ffi.cu defines C-linkage twins of the functions it declares with C++ linkage.
struct cublasContext;
typedef struct cublasContext *cublasHandle_t;
struct KBlasHandle;
struct KBlasWorkspace;
typedef struct KBlasWorkspace *kblasWorkspace_t;
typedef struct KBlasHandle *kblasHandle_t;

#define kblasx_min(a, b) ((a) > (b) ? (b) : (a))
#define kblasx_max(a, b) ((a) < (b) ? (b) : (a))

// reference src/kblas_struct.h:43-91
struct KBlasWorkspaceState {
  size_t h_data_bytes, h_ptrs_bytes;  // host data / host pointer arrays
  size_t d_data_bytes, d_ptrs_bytes;  // device data / device pointer arrays

  KBlasWorkspaceState() { reset(); }
  KBlasWorkspaceState(size_t hd, size_t hp, size_t dd, size_t dp)
      : h_data_bytes(hd), h_ptrs_bytes(hp), d_data_bytes(dd), d_ptrs_bytes(dp) {}
  void reset() { h_data_bytes = h_ptrs_bytes = d_data_bytes = d_ptrs_bytes = 0; }
  // element-wise max: what several queries need together
  void pad(const KBlasWorkspaceState *o) {
    h_data_bytes = kblasx_max(h_data_bytes, o->h_data_bytes);
    h_ptrs_bytes = kblasx_max(h_ptrs_bytes, o->h_ptrs_bytes);
    d_data_bytes = kblasx_max(d_data_bytes, o->d_data_bytes);
    d_ptrs_bytes = kblasx_max(d_ptrs_bytes, o->d_ptrs_bytes);
  }
  void set(const KBlasWorkspaceState *o) { *this = *o; }
  bool isSufficient(const KBlasWorkspaceState *have) const {
    return h_data_bytes <= have->h_data_bytes && h_ptrs_bytes <= have->h_ptrs_bytes &&
           d_data_bytes <= have->d_data_bytes && d_ptrs_bytes <= have->d_ptrs_bytes;
  }
};
typedef KBlasWorkspaceState *kblasWorkspaceState_t;

// reference src/kblas_struct.h:93-290 (four grow-only regions, FILO push/pop)
struct KBlasWorkspace {
  void *h_data;
  void **h_ptrs;
  void *d_data;
  void **d_ptrs;
  KBlasWorkspaceState allocated_ws_state;  // what is allocated now
  KBlasWorkspaceState requested_ws_state;  // max of the queries since the last allocate()
  KBlasWorkspaceState consumed_ws_state;   // pushed by a routine holding the handle
  bool allocated;

  KBlasWorkspace() { reset(); }
  ~KBlasWorkspace() {
    if (allocated) deallocate();
  }
  void reset();
  int allocate();    // KBLAS_Success / KBLAS_Error_Allocation / KBLAS_Error_Deallocation
  int deallocate();  // KBLAS_Success / KBLAS_Error_Deallocation
  KBlasWorkspaceState getAvailable() const;
  void *push_d_data(size_t bytes);
  void pop_d_data(size_t bytes);
  void *push_d_ptrs(size_t bytes);
  void pop_d_ptrs(size_t bytes);
  void *push_h_data(size_t bytes);
  void pop_h_data(size_t bytes);
  void *push_h_ptrs(size_t bytes);
  void pop_h_ptrs(size_t bytes);
};

// reference src/kblas_struct.h:292-309
struct KBlasWorkspaceGuard {
  KBlasWorkspaceState pushed_ws;
  KBlasWorkspace *ws_ptr;
  KBlasWorkspaceGuard(const KBlasWorkspaceState &pushed, KBlasWorkspace &ws) : pushed_ws(pushed), ws_ptr(&ws) {}
  ~KBlasWorkspaceGuard() {
    ws_ptr->pop_d_data(pushed_ws.d_data_bytes);
    ws_ptr->pop_d_ptrs(pushed_ws.d_ptrs_bytes);
    ws_ptr->pop_h_data(pushed_ws.h_data_bytes);
    ws_ptr->pop_h_ptrs(pushed_ws.h_ptrs_bytes);
  }
};

// two-event timer, reference src/kblas_gpu_timer.h:24-73
struct kblas_gpu_timer {
  cudaEvent_t start_event, stop_event;
  float elapsed_time;
  bool recorded_end;
  void init();
  void destroy();
  void start(cudaStream_t s);
  void recordEnd(cudaStream_t s);
  double stop(cudaStream_t s);  // seconds
};

// How info_array is treated (SURVEY §0 finding 1).
enum KBlasxInfoMode {
  KBLASX_INFO_COMPAT = 0,  // never written (bit-identical to the reference)
  KBLASX_INFO_LAPACK = 1   // info[b] = 0 / (j+1) of first non-positive pivot
};

// reference src/kblas_struct.h:311-456
struct KBlasHandle {
  cublasHandle_t cublas_handle;  // lazily created (kblasGetCublasHandle)
  cudaStream_t stream;
  cudaStream_t streams[KBLAS_NSTREAMS];
  int nStreams;
  int use_magma, device_id, create_cublas;
  kblas_gpu_timer timer;
  KBlasWorkspace work_space;

  // ---- B200-native additions -------------------------------------------------
  int sm_count;            // multiProcessorCount of device_id
  int smem_optin_max;      // cudaDevAttrMaxSharedMemoryPerBlockOptin of device_id (227 KiB on B200)
  int info_mode;           // KBlasxInfoMode
  int variant_override;    // -1 = auto; tuning / ablation hook (env KBLAS_B200_VARIANT)
  int tri_flags;           // TRI_FLAG_UPPER | TRI_FLAG_UNIT of the triangular solve being dispatched (0 outside such a call)
  int exact_stores;        // env KBLAS_B200_ELEMENT_EXACT_STORES=1: potrf never stores a strict-upper element
  long launch_count;       // kernels launched through this handle
  const char *last_kernel; // name of the last dispatched kernel variant
  void *host_pipe;         // staging buffers / streams of the host-memory entry points (host_pipeline.cu), lazily created
  // per-handle (= per-device, single-threaded by contract) launch cache, keyed by kernel address: resident CTAs
  // per SM and "dynamic shared-memory limit raised".  Function-local statics would be shared by every device
  // and every host thread of the process.
  struct KernelNote {
    const void *fn;
    int ctas_per_sm;   // 0 = not queried yet
    int smem_limit;    // bytes the MaxDynamicSharedMemorySize attribute was raised to (0 = never)
  };
  static const int KERNEL_NOTES = 128;
  KernelNote kernel_notes[KERNEL_NOTES];
  int n_kernel_notes;
  KernelNote *kernel_note(const void *fn);

  explicit KBlasHandle(int use_magma, cudaStream_t stream = 0, int device_id = 0);
  ~KBlasHandle();
  void tic() { timer.start(stream); }
  void recordEnd() { timer.recordEnd(stream); }
  double toc() { return timer.stop(stream); }
  int SetStream(cudaStream_t s);
  int CreateStreams(int n);
  void note_launch(const char *name) {
    ++launch_count;
    last_kernel = name;
  }
};
