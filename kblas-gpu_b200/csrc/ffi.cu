// ffi.cu -- C-linkage twins of the management / workspace API (include/kblas_ffi.h).
//
// The reference declares these calls with C++ linkage only (include/kblas.h:54-108,
// kblas_batch.h:773,786,1380,1391,2077,2089,2772,2785), which no FFI can bind by name.
// This translation unit therefore never includes kblas.h: it re-exports the same
// implementations (namespace kblasx) under unmangled names.
#include <cstring>
#include "kblas_common.h"
#include "kernels/potrf_smem.cuh"

#define KBLASX_VERSION "kblas-b200 0.1.0 (sm_100a; potrf/trsm/potrs/posv batch; drop-in for KBLAS-GPU 3.0.0 API)"

extern "C" {

int kblasCreate(kblasHandle_t *handle) { return kblasx::create(handle); }
int kblasDestroy(kblasHandle_t *handle) { return kblasx::destroy(handle); }
void kblasTimerTic(kblasHandle_t handle) { handle->tic(); }
void kblasTimerRecordEnd(kblasHandle_t handle) { handle->recordEnd(); }
double kblasTimerToc(kblasHandle_t handle) { return handle->toc(); }
int kblasCreateStreams(kblasHandle_t handle, int nStreams) { return handle->CreateStreams(nStreams); }
void *kblasGetStream(kblasHandle_t handle) { return (void *)handle->stream; }
void kblasSetStream(kblasHandle_t handle, void *stream) { kblasx::set_stream(handle, (cudaStream_t)stream); }
void *kblasGetCublasHandle(kblasHandle_t handle) { return (void *)kblasx::get_cublas(handle); }
int kblasEnableMagma(kblasHandle_t) {
  printf("ERROR: KBLAS is compiled without magma!\n");
  return KBLAS_Error_NotInitialized;
}
const char *kblasGetErrorString(int error) { return kblasx::error_string(error); }
int kblasAllocateWorkspace(kblasHandle_t handle) { return handle->work_space.allocate(); }
int kblasFreeWorkspace(kblasHandle_t handle) { return handle->work_space.deallocate(); }

#define REQ(h) (&((h)->work_space.requested_ws_state))
void kblas_trsm_batch_wsquery(kblasHandle_t h, char side, int m, int n, int batchCount) {
  kblasx::trsm_batch_wsquery_core(false, batchCount, side, m, n, REQ(h));
}
void kblas_trsm_batch_strided_wsquery(kblasHandle_t h, char side, int m, int n, int batchCount) {
  kblasx::trsm_batch_wsquery_core(true, batchCount, side, m, n, REQ(h));
}
void kblas_potrf_batch_wsquery(kblasHandle_t h, int n, int batchCount) {
  kblasx::potrf_batch_wsquery_core(false, n, batchCount, REQ(h));
}
void kblas_potrf_batch_strided_wsquery(kblasHandle_t h, int n, int batchCount) {
  kblasx::potrf_batch_wsquery_core(true, n, batchCount, REQ(h));
}
void kblas_potrs_batch_wsquery(kblasHandle_t h, int m, int n, int batchCount) {
  kblasx::potrs_batch_wsquery_core(false, m, n, batchCount, REQ(h));
}
void kblas_potrs_batch_strided_wsquery(kblasHandle_t h, int m, int n, int batchCount) {
  kblasx::potrs_batch_wsquery_core(true, m, n, batchCount, REQ(h));
}
void kblas_posv_batch_wsquery(kblasHandle_t h, char side, int m, int n, int batchCount) {
  kblasx::posv_batch_wsquery_core(false, m, n, side, batchCount, REQ(h));
}
void kblas_posv_batch_strided_wsquery(kblasHandle_t h, char side, int m, int n, int batchCount) {
  kblasx::posv_batch_wsquery_core(true, m, n, side, batchCount, REQ(h));
}
void kblas_gemm_batch_strided_wsquery(kblasHandle_t h, int batchCount) {
  kblasx::gemm_batch_strided_wsquery_core(batchCount, REQ(h));
}
void kblas_syrk_batch_wsquery(kblasHandle_t h, int m, int batchCount) { kblasx::syrk_batch_wsquery_core(m, batchCount, REQ(h)); }
#define KX_INV_WS_C(NAME)                                                                                                   \
  void kblas_##NAME##_batch_wsquery(kblasHandle_t h, int n, int batchCount) { kblasx::NAME##_batch_wsquery_core(false, n, batchCount, REQ(h)); } \
  void kblas_##NAME##_batch_strided_wsquery(kblasHandle_t h, int n, int batchCount) {                                       \
    kblasx::NAME##_batch_wsquery_core(true, n, batchCount, REQ(h));                                                         \
  }
KX_INV_WS_C(trtri)
KX_INV_WS_C(lauum)
KX_INV_WS_C(potri)
KX_INV_WS_C(poti)
#undef REQ

int kblasSset_pointer_1(float **out, const float *in, int lda, long off, long batchCount, void *stream) {
  return kblasx::set_pointer_1<float>(out, in, lda, off, batchCount, (cudaStream_t)stream);
}
int kblasDset_pointer_1(double **out, const double *in, int lda, long off, long batchCount, void *stream) {
  return kblasx::set_pointer_1<double>(out, in, lda, off, batchCount, (cudaStream_t)stream);
}
int kblasSset_pointer_2(float **out1, const float *in1, int ld1, long off1, float **out2, const float *in2, int ld2,
                        long off2, long batchCount, void *stream) {
  return kblasx::set_pointer_2<float>(out1, in1, ld1, off1, out2, in2, ld2, off2, batchCount, (cudaStream_t)stream);
}
int kblasDset_pointer_2(double **out1, const double *in1, int ld1, long off1, double **out2, const double *in2,
                        int ld2, long off2, long batchCount, void *stream) {
  return kblasx::set_pointer_2<double>(out1, in1, ld1, off1, out2, in2, ld2, off2, batchCount, (cudaStream_t)stream);
}
int kblasSset_pointer_3(float **out1, const float *in1, int ld1, long off1, float **out2, const float *in2, int ld2,
                        long off2, float **out3, const float *in3, int ld3, long off3, long batchCount, void *stream) {
  return kblasx::set_pointer_3<float>(out1, in1, ld1, off1, out2, in2, ld2, off2, out3, in3, ld3, off3, batchCount,
                                      (cudaStream_t)stream);
}
int kblasDset_pointer_3(double **out1, const double *in1, int ld1, long off1, double **out2, const double *in2,
                        int ld2, long off2, double **out3, const double *in3, int ld3, long off3, long batchCount,
                        void *stream) {
  return kblasx::set_pointer_3<double>(out1, in1, ld1, off1, out2, in2, ld2, off2, out3, in3, ld3, off3, batchCount,
                                       (cudaStream_t)stream);
}
int kblas_iset_value_1(int *output_array, int input, long batchCount, void *stream) {
  return iset_value_1(output_array, input, batchCount, (cudaStream_t)stream);
}
int kblas_iset_value_2(int *output_array1, int input1, int *output_array2, int input2, long batchCount, void *stream) {
  return iset_value_2(output_array1, input1, output_array2, input2, batchCount, (cudaStream_t)stream);
}
int kblas_iset_value_4(int *o1, int i1, int *o2, int i2, int *o3, int i3, int *o4, int i4, long batchCount, void *stream) {
  return iset_value_4(o1, i1, o2, i2, o3, i3, o4, i4, batchCount, (cudaStream_t)stream);
}
int kblas_iset_value_5(int *o1, int i1, int *o2, int i2, int *o3, int i3, int *o4, int i4, int *o5, int i5,
                       long batchCount, void *stream) {
  return iset_value_5(o1, i1, o2, i2, o3, i3, o4, i4, o5, i5, batchCount, (cudaStream_t)stream);
}

// ---- introspection (no reference counterpart) ---------------------------------------------
static void ws_out(const KBlasWorkspaceState &s, size_t out[4]) {
  out[0] = s.h_data_bytes;
  out[1] = s.h_ptrs_bytes;
  out[2] = s.d_data_bytes;
  out[3] = s.d_ptrs_bytes;
}

int kblasx_workspace_state(kblasHandle_t handle, int which, size_t out[4]) {
  if (!handle || which < 0 || which > 2) return KBLAS_Error_WrongInput;
  const KBlasWorkspace &w = handle->work_space;
  ws_out(which == 0 ? w.requested_ws_state : which == 1 ? w.allocated_ws_state : w.consumed_ws_state, out);
  return KBLAS_Success;
}

int kblasx_wsquery_bytes(int op, int strided, char side, int m, int n, int batchCount, size_t out[4]) {
  KBlasWorkspaceState s;
  switch (op) {
    case kblasx::WS_TRSM: kblasx::trsm_batch_wsquery_core(strided != 0, batchCount, side, m, n, &s); break;
    case kblasx::WS_POTRF: kblasx::potrf_batch_wsquery_core(strided != 0, n, batchCount, &s); break;
    case kblasx::WS_POTRS: kblasx::potrs_batch_wsquery_core(strided != 0, m, n, batchCount, &s); break;
    case kblasx::WS_POSV: kblasx::posv_batch_wsquery_core(strided != 0, m, n, side, batchCount, &s); break;
    default: return KBLAS_Error_WrongInput;
  }
  ws_out(s, out);
  return KBLAS_Success;
}

// shared-memory slot plan of the 32 < n <= 256 fp64 Cholesky (kernels/potrf_smem.cuh): out[I*8+K] = slot of block (I,K),
// returns the number of 8 KiB slots (host logic, no GPU needed)
int kblasx_potrf_smem_plan(int nblk, unsigned char out[64]) {
  if (nblk < 1 || nblk > 8) return KBLAS_Error_WrongInput;
  const kblasx::SmemPotrfPlan p = kblasx::plan_potrf_slots(nblk);
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 8; ++k) out[i * 8 + k] = p.slot[i][k];
  return p.nslots;
}

long kblasx_launch_count(kblasHandle_t handle) { return handle ? handle->launch_count : -1; }
const char *kblasx_last_kernel(kblasHandle_t handle) { return handle ? handle->last_kernel : "none"; }
const char *kblasx_version(void) { return KBLASX_VERSION; }
int kblasx_reg_size(int n) { return REG_SIZE(n) ? 1 : 0; }
int kblasx_closest_reg_size(int n) { return CLOSEST_REG_SIZE(n); }

}  // extern "C"
