// host_pipeline.cu -- kblasx{S,D}potrf_batch_strided_host: the strided batch Cholesky fed from HOST memory.
//
// No reference counterpart: the reference API takes device pointers only, and its callers
// (testing/batch_triangular/test_Xpotrf_batch.cpp:170-206) move whole matrices with cudaMemcpy before
// and after the call.  For 8 KiB matrices the PCIe link, not the kernel, sets the end-to-end rate
// (2 x 8 GiB cross the bus per 2^20-matrix batch; the kernel needs 2 ms of the ~185 ms).  This entry point
// is the one-call form of that workflow: the batch is cut into 256 MiB chunks that stream through three
// device staging buffers on three streams (H2D / factorisation / D2H), so both PCIe directions and the
// kernel overlap; measured on B200 (PCIe Gen5 x16): 45.7 GB/s per direction, both directions at once.
//
// Two ways of putting fewer bytes on the bus were built and MEASURED SLOWER, and are therefore not the default:
//   * KBLAS_B200_HOSTCOPY=tri: only the lower triangle travels, in 8-column groups (rows 8g.. of columns
//     8g..8g+7: 62.5 % of a 32 x 32 matrix) as strided 3-D copies.  The copy engines move the 64..256-byte
//     rows at 20 GB/s per direction: 263 ms per 2^20 matrices against 188 ms for whole-array copies.
//     (The diagonal 8 x 8 blocks travel whole, so the strict-upper elements the kernel re-writes carry the
//     caller's bits both ways; in place the result is still bit-identical.)
//   * zero-copy (the kernel run directly on pinned host memory, tools/zero_copy_test.py): 350 ms.
// In place (A_out == A_in) the result is bit-identical to cudaMemcpy + kblas_potrf_batch + cudaMemcpy.
// Out of place, whole-array mode: A_out receives a copy of A_in's storage (padding between columns and between
// matrices included, up to the last element of the last matrix) with the lower triangles replaced by the factors.
// Out of place in tri mode, elements of A_out above the diagonal blocks are not written.
// Host buffers must hold (batchCount-1)*strideA + lda*(n-1) + n elements -- the minimal strided allocation.
#include <cstdlib>

#include "kblas_common.h"
#include "kblas_struct.h"
#include "potrf_batch.h"

namespace kblasx {

static const int HP_NBUF = 3;

struct HostPipe {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[HP_NBUF], ev_k[HP_NBUF], ev_out[HP_NBUF];
  void *dbuf[HP_NBUF] = {nullptr, nullptr, nullptr};
  int *dinfo[HP_NBUF] = {nullptr, nullptr, nullptr};
  size_t cap = 0;       // bytes per staging buffer
  size_t info_cap = 0;  // ints per info buffer
  bool ready = false;
};

static int hp_init(HostPipe *p) {
  if (p->ready) return KBLAS_Success;
  check_error_ret(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking), KBLAS_CUDA_Error);
  check_error_ret(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking), KBLAS_CUDA_Error);
  for (int i = 0; i < HP_NBUF; ++i) {
    check_error_ret(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming), KBLAS_CUDA_Error);
    check_error_ret(cudaEventCreateWithFlags(&p->ev_k[i], cudaEventDisableTiming), KBLAS_CUDA_Error);
    check_error_ret(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming), KBLAS_CUDA_Error);
  }
  p->ready = true;
  return KBLAS_Success;
}

static int hp_reserve(HostPipe *p, size_t bytes, size_t ninfo) {
  if (bytes > p->cap) {
    for (int i = 0; i < HP_NBUF; ++i) {
      if (p->dbuf[i]) cudaFree(p->dbuf[i]);
      p->dbuf[i] = nullptr;
      check_error_ret(cudaMalloc(&p->dbuf[i], bytes), KBLAS_Error_Allocation);
    }
    p->cap = bytes;
  }
  if (ninfo > p->info_cap) {
    for (int i = 0; i < HP_NBUF; ++i) {
      if (p->dinfo[i]) cudaFree(p->dinfo[i]);
      p->dinfo[i] = nullptr;
      check_error_ret(cudaMalloc(&p->dinfo[i], ninfo * sizeof(int)), KBLAS_Error_Allocation);
    }
    p->info_cap = ninfo;
  }
  return KBLAS_Success;
}

void host_pipe_destroy(void *pp) {
  HostPipe *p = static_cast<HostPipe *>(pp);
  if (!p) return;
  for (int i = 0; i < HP_NBUF; ++i) {
    if (p->dbuf[i]) cudaFree(p->dbuf[i]);
    if (p->dinfo[i]) cudaFree(p->dinfo[i]);
    if (p->ready) {
      cudaEventDestroy(p->ev_in[i]);
      cudaEventDestroy(p->ev_k[i]);
      cudaEventDestroy(p->ev_out[i]);
    }
  }
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
}

// Copy `count` matrices between a host and a device array of identical (lda, stride) geometry.
// tri: lower triangle in 8-column groups through strided 3-D copies; otherwise one contiguous copy.
static cudaError_t copy_matrices(void *dst, const void *src, size_t es, int n, int lda, long stride, long count, bool tri,
                                 bool last, cudaMemcpyKind kind, cudaStream_t s, long last_elems = -1) {
  // whole-array mode: count*stride elements, except that the chunk holding the batch's LAST matrix stops at that
  // matrix's last element ((count-1)*stride + lda*(n-1) + n elements): a caller with stride > lda*n owns no storage
  // behind it (the usual minimal strided allocation)
  if (!tri) {
    const size_t tail = last_elems >= 0 ? (size_t)last_elems : (size_t)lda * (n - 1) + n;  // elements of the last matrix
    const size_t elems = last ? ((size_t)(count - 1) * stride + tail) : (size_t)count * stride;
    return cudaMemcpyAsync(dst, src, elems * es, kind, s);
  }
  const size_t pitch = (size_t)lda * es;
  const size_t cols_per_matrix = (size_t)(stride / lda);  // tri requires stride % lda == 0
  for (int c0 = 0; c0 < n; c0 += 8) {
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), pitch, pitch, cols_per_matrix);
    p.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, cols_per_matrix);
    p.srcPos = make_cudaPos((size_t)c0 * es, (size_t)c0, 0);
    p.dstPos = p.srcPos;
    p.extent = make_cudaExtent((size_t)(n - c0) * es, (size_t)((n - c0) < 8 ? (n - c0) : 8), (size_t)count);
    p.kind = kind;
    cudaError_t e = cudaMemcpy3DAsync(&p, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// PACKED: A_in / A_out hold LAPACK packed lower matrices (n(n+1)/2 elements each, stride strideA, lda ignored) and the
// chunks are factored by the packed kernels -- half the bytes of the full layout cross PCIe in each direction.
template <typename T, bool PACKED>
int potrf_batch_strided_host(KBlasHandle *h, char uplo, int n, const T *A_in, T *A_out, int lda, long strideA, int batchCount,
                             int *info_host) {
  if (PACKED) lda = 1;  // a packed matrix is one "column" of n(n+1)/2 elements for the copy arithmetic below
  const int nrows = PACKED ? (n * (n + 1)) / 2 : n;
  if (uplo == KBLAS_Upper) {
    printf("Upper POTRF_BATCH is not implemented yet\n");
    return KBLAS_NotImplemented;
  }
  if (batchCount <= 0) return KBLAS_UnknownError;  // same code as the device entry point's empty batch
  if (n <= 0) return KBLAS_Success;
  if (PACKED ? (strideA < (long)nrows || n > 32) : (lda < n || strideA < (long)lda * n)) return PACKED && n > 32 ? KBLAS_NotImplemented : KBLAS_Error_WrongInput;
  if (!h->host_pipe) h->host_pipe = new HostPipe();
  HostPipe *p = static_cast<HostPipe *>(h->host_pipe);
  int rc = hp_init(p);
  if (rc != KBLAS_Success) return rc;

  // transfer mode: whole-array copies unless KBLAS_B200_HOSTCOPY=tri (see the header: measured slower)
  const char *mode = getenv("KBLAS_B200_HOSTCOPY");
  const bool want_tri = mode && mode[0] == 't';
  const bool tri = !PACKED && want_tri && n > 8 && (strideA % lda == 0);
  const bool want_info = info_host && h->info_mode == KBLASX_INFO_LAPACK;

  // chunk: ~256 MiB of matrices per staging buffer (large enough that the per-copy latency
  // vanishes, small enough that the pipeline fills quickly)
  const size_t mat_bytes = (size_t)strideA * sizeof(T);
  size_t chunk_bytes = 256ull << 20;
  if (const char *cb = getenv("KBLAS_B200_HOSTCHUNK_MB")) chunk_bytes = (size_t)(atof(cb) * (1 << 20));  // tests / tuning
  long chunk = (long)(chunk_bytes / mat_bytes);
  if (chunk < 1) chunk = 1;
  if (chunk > batchCount) chunk = batchCount;
  rc = hp_reserve(p, (size_t)chunk * mat_bytes, want_info ? (size_t)chunk : 0);
  if (rc != KBLAS_Success) return rc;

  const long nchunks = ((long)batchCount + chunk - 1) / chunk;
  for (long c = 0; c < nchunks; ++c) {
    const long lo = c * chunk;
    const long cnt = (lo + chunk <= batchCount) ? chunk : (long)batchCount - lo;
    const int b = (int)(c % HP_NBUF);
    T *d = static_cast<T *>(p->dbuf[b]);
    // H2D once the previous result has left this buffer
    if (c >= HP_NBUF) check_error_ret(cudaStreamWaitEvent(p->s_in, p->ev_out[b], 0), KBLAS_CUDA_Error);
    check_error_ret(copy_matrices(d, A_in + lo * strideA, sizeof(T), n, lda, strideA, cnt, tri, c + 1 == nchunks, cudaMemcpyHostToDevice, p->s_in,
                                  PACKED ? (long)nrows : -1L),
                    KBLAS_CUDA_Error);
    check_error_ret(cudaEventRecord(p->ev_in[b], p->s_in), KBLAS_CUDA_Error);
    // factorise on the handle's stream
    check_error_ret(cudaStreamWaitEvent(h->stream, p->ev_in[b], 0), KBLAS_CUDA_Error);
    {
      BatchRef<T, true> ref = {d, strideA};  // no workspace protocol here: the kernels use none
      if (PACKED)
        rc = pptrf_batch_core<T, true>(h, uplo, n, ref, (int)cnt, want_info ? p->dinfo[b] : nullptr,
                                       ((size_t)strideA * sizeof(T)) % 16 == 0);   // cudaMalloc'ed staging buffers are aligned
      else
        rc = potrf_batch_core<T, true>(h, uplo, n, ref, lda, (int)cnt, want_info ? p->dinfo[b] : nullptr);
    }
    if (rc != KBLAS_Success) return rc;
    check_error_ret(cudaEventRecord(p->ev_k[b], h->stream), KBLAS_CUDA_Error);
    // D2H
    check_error_ret(cudaStreamWaitEvent(p->s_out, p->ev_k[b], 0), KBLAS_CUDA_Error);
    check_error_ret(copy_matrices(A_out + lo * strideA, d, sizeof(T), n, lda, strideA, cnt, tri, c + 1 == nchunks, cudaMemcpyDeviceToHost, p->s_out,
                                  PACKED ? (long)nrows : -1L),
                    KBLAS_CUDA_Error);
    if (want_info)
      check_error_ret(cudaMemcpyAsync(info_host + lo, p->dinfo[b], (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, p->s_out),
                      KBLAS_CUDA_Error);
    check_error_ret(cudaEventRecord(p->ev_out[b], p->s_out), KBLAS_CUDA_Error);
  }
  // the host result is complete when this returns
  check_error_ret(cudaStreamSynchronize(p->s_out), KBLAS_CUDA_Error);
  return KBLAS_Success;
}

template int potrf_batch_strided_host<float, false>(KBlasHandle *, char, int, const float *, float *, int, long, int, int *);
template int potrf_batch_strided_host<double, false>(KBlasHandle *, char, int, const double *, double *, int, long, int, int *);
template int potrf_batch_strided_host<float, true>(KBlasHandle *, char, int, const float *, float *, int, long, int, int *);
template int potrf_batch_strided_host<double, true>(KBlasHandle *, char, int, const double *, double *, int, long, int, int *);

}  // namespace kblasx

extern "C" {
int kblasxSpotrf_batch_strided_host(KBlasHandle *handle, char uplo, int n, const float *A_in, float *A_out, int lda,
                                    long strideA, int batchCount, int *info_host) {
  return kblasx::potrf_batch_strided_host<float, false>(handle, uplo, n, A_in, A_out, lda, strideA, batchCount, info_host);
}
int kblasxDpotrf_batch_strided_host(KBlasHandle *handle, char uplo, int n, const double *A_in, double *A_out, int lda,
                                    long strideA, int batchCount, int *info_host) {
  return kblasx::potrf_batch_strided_host<double, false>(handle, uplo, n, A_in, A_out, lda, strideA, batchCount, info_host);
}
// packed lower storage in host memory (kblasx?pptrf_batch_strided semantics, same pipeline)
int kblasxSpptrf_batch_strided_host(KBlasHandle *handle, char uplo, int n, const float *AP_in, float *AP_out, long strideAP,
                                    int batchCount, int *info_host) {
  return kblasx::potrf_batch_strided_host<float, true>(handle, uplo, n, AP_in, AP_out, 1, strideAP, batchCount, info_host);
}
int kblasxDpptrf_batch_strided_host(KBlasHandle *handle, char uplo, int n, const double *AP_in, double *AP_out, long strideAP,
                                    int batchCount, int *info_host) {
  return kblasx::potrf_batch_strided_host<double, true>(handle, uplo, n, AP_in, AP_out, 1, strideAP, batchCount, info_host);
}
}
