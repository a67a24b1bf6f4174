// potrf_batch.h -- internal interface of the potrf dispatch (used by posv).
#pragma once
#include "kblas_struct.h"
#include "kernels/common.cuh"

namespace kblasx {
template <typename T, bool STRIDED>
int potrf_batch_core(KBlasHandle *h, char uplo, int n, BatchRef<T, STRIDED> A, int lda, int batchCount, int *info);
template <typename T>
int potrf_batch_strided(KBlasHandle *h, char uplo, int n, T *A, int lda, long strideA, int batchCount, int *info);
template <typename T>
int potrf_batch_ptrs(KBlasHandle *h, char uplo, int n, T **A, long elem_off, int lda, int batchCount, int *info);
// packed lower storage (pptrf_batch.cu); aligned = every matrix starts on a 16-byte boundary (TMA path allowed)
template <typename T, bool STRIDED>
int pptrf_batch_core(KBlasHandle *h, char uplo, int n, BatchRef<T, STRIDED> AP, int batchCount, int *info, bool aligned);
}  // namespace kblasx
