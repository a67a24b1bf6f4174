// tri_batch.h -- internal interface of the triangular-solve dispatch (trsm, potrs, posv).
#pragma once
#include "kblas_struct.h"
#include "kernels/common.cuh"

namespace kblasx {
// op: TRI_FORWARD / TRI_BACKWARD / TRI_BOTH (kernels/trsm_small.cuh); left: factor acts on the
// left of B (k = m) or on the right (k = n).
template <typename T, bool STRIDED>
int tri_solve_core(KBlasHandle *h, bool left, int op, int m, int n, T alpha, BatchRef<const T, STRIDED> A, int lda,
                   BatchRef<T, STRIDED> B, int ldb, int batchCount);

template <typename T, bool STRIDED>
int potrs_batch_core(KBlasHandle *h, char side, char uplo, int m, int n, BatchRef<const T, STRIDED> A, int lda,
                     BatchRef<T, STRIDED> B, int ldb, int batchCount);
}  // namespace kblasx
