/*
 * kblas.h -- public entry header of the B200-native KBLAS batched
 * very-small-matrix Cholesky path (potrf / trsm / potrs / posv, s + d,
 * strided + pointer-array, uniform batch).
 *
 * Drop-in for the reference's include/kblas.h restricted to that path:
 * the handle type, the management calls and the workspace calls keep the
 * reference's names, argument meaning, return codes AND linkage.  In the
 * reference these management calls have C++ linkage (include/kblas.h:54-108
 * declares them outside any extern "C" block), so this header declares them
 * the same way and the library exports the same mangled symbols.  The very
 * same functions are ALSO exported with C linkage for FFI users (ctypes, cgo,
 * JNI ...): see include/kblas_ffi.h.
 */
#ifndef KBLAS_B200_H
#define KBLAS_B200_H

#include <cuda_runtime_api.h>
#include "kblas_defs.h"

/* cuBLAS is not used by this implementation; the opaque type is declared so
 * that kblasGetCublasHandle() keeps its signature without forcing
 * <cublas_v2.h> on every includer (identical typedef to cublas_api.h). */
struct cublasContext;
typedef struct cublasContext *cublasHandle_t;

/* reference include/kblas.h:33-36 */
struct KBlasHandle;
struct KBlasWorkspace;
typedef struct KBlasWorkspace *kblasWorkspace_t;
typedef struct KBlasHandle    *kblasHandle_t;

#ifdef __cplusplus

/* ---- handle life cycle (reference include/kblas.h:54-59, src/kblas_common.cu:35-82) */
/** Create a handle bound to the CURRENT device, stream 0. Returns KBLAS_Success (1). */
int kblasCreate(kblasHandle_t *handle);
/** Destroy the handle, its streams, timer events and workspace; sets *handle = NULL. */
int kblasDestroy(kblasHandle_t *handle);

/* ---- timer: two CUDA events recorded on handle->stream
 *      (reference include/kblas.h:60-62, src/kblas_gpu_timer.h:24-73) */
void   kblasTimerTic(kblasHandle_t handle);
void   kblasTimerRecordEnd(kblasHandle_t handle);
/** seconds between Tic and RecordEnd (records End itself if it was not). */
double kblasTimerToc(kblasHandle_t handle);

/* ---- streams (reference include/kblas.h:67-76, src/kblas_common.cu:111-125) */
int          kblasCreateStreams(kblasHandle_t handle, int nStreams);
cudaStream_t kblasGetStream(kblasHandle_t handle);
void         kblasSetStream(kblasHandle_t handle, cudaStream_t stream);

/** cuBLAS handle owned by the KBLAS handle (reference include/kblas.h:81).
 *  Created lazily on first call: nothing on the hot path uses cuBLAS. */
cublasHandle_t kblasGetCublasHandle(kblasHandle_t handle);

/** MAGMA is never compiled in: returns KBLAS_Error_NotInitialized like a
 *  reference build without USE_MAGMA would (src/kblas_common.cu:53-62). */
int kblasEnableMagma(kblasHandle_t handle);

/** reference include/kblas.h:92, src/kblas_common.cu:171-202 */
const char *kblasGetErrorString(int error);

/* ---- workspace (reference include/kblas.h:102-108, src/kblas_common.cu:88-97) */
/** Allocate (grow-only) what the *_wsquery calls since the last allocation asked for. */
int kblasAllocateWorkspace(kblasHandle_t handle);
/** Free every workspace buffer and reset the recorded sizes. */
int kblasFreeWorkspace(kblasHandle_t handle);

#endif /* __cplusplus */

#include "kblas_batch.h"

#endif /* KBLAS_B200_H */
