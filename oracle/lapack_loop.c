/*
 * lapack_loop.c -- the reference test harness's CPU check loop, restated.
 *
 * *** TEST INFRASTRUCTURE / REPORTED BASELINE ONLY *** (same rules as kblas_oracle.c)
 *
 * Reference: testing/batch_triangular/test_Xpotrf_batch.cpp:308-321 -- a serial loop
 *     for (s = 0; s < batchCount; s++) LAPACK_POTRF("L", &N, h_R + s*lda*N, &lda, &info);
 * timed with gettimeofday.  Nothing defines USE_OPENMP there and MKL is linked sequential
 * (testing/Makefile:44), so the harness's CPU baseline is ONE core; `threads` > 1 runs the
 * same loop as an OpenMP parallel-for over matrices (the USE_OPENMP branch, :299-305).
 * LAPACK is the OpenBLAS bundled with scipy (symbols scipy_dpotrf_ ..., LP64), resolved
 * by the caller with dlopen and passed in as function pointers -- no link-time dependency.
 */
#include <stddef.h>
#include <sys/time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*potrf_fn)(const char *uplo, const int *n, void *a, const int *lda, int *info);
typedef void (*trsm_fn)(const char *side, const char *uplo, const char *trans, const char *diag, const int *m,
                        const int *n, const void *alpha, const void *a, const int *lda, void *b, const int *ldb);

static double now_s(void) {
  struct timeval t;
  gettimeofday(&t, NULL);
  return t.tv_sec + 1e-6 * t.tv_usec;
}

/* returns wall seconds; *bad_info = number of matrices with info != 0 */
double lapack_potrf_loop(potrf_fn potrf, int elem_size, int n, void *A, int lda, long strideA, long batchCount,
                         int threads, long *bad_info) {
  long bad = 0;
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) reduction(+ : bad) schedule(static)
#endif
  for (long s = 0; s < batchCount; s++) {
    int info = 0;
    potrf("L", &n, (char *)A + (size_t)s * strideA * elem_size, &lda, &info);
    bad += (info != 0);
  }
  const double t1 = now_s();
  (void)threads;
  if (bad_info) *bad_info = bad;
  return t1 - t0;
}

/* POSV check path of the harness (test_Xposv_batch.cpp:313-326): potrf + trsm(R,L,T) + trsm(R,L,N) */
double lapack_posv_loop(potrf_fn potrf, trsm_fn trsm, int elem_size, int m, int n, void *A, int lda, long strideA,
                        void *B, int ldb, long strideB, long batchCount, int threads, const void *one) {
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
  for (long s = 0; s < batchCount; s++) {
    int info = 0;
    char *a = (char *)A + (size_t)s * strideA * elem_size;
    char *b = (char *)B + (size_t)s * strideB * elem_size;
    potrf("L", &n, a, &lda, &info);
    trsm("R", "L", "T", "N", &m, &n, one, a, &lda, b, &ldb);
    trsm("R", "L", "N", "N", &m, &n, one, a, &lda, b, &ldb);
  }
  (void)threads;
  return now_s() - t0;
}

/* TRSM / POTRS-given-factor loops for the other bench configurations */
double lapack_trsm_loop(trsm_fn trsm, int elem_size, char side, char trans, int m, int n, const void *alpha, void *A,
                        int lda, long strideA, void *B, int ldb, long strideB, long batchCount, int threads) {
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
  for (long s = 0; s < batchCount; s++) {
    char *a = (char *)A + (size_t)s * strideA * elem_size;
    char *b = (char *)B + (size_t)s * strideB * elem_size;
    trsm(&side, "L", &trans, "N", &m, &n, alpha, a, &lda, b, &ldb);
  }
  (void)threads;
  return now_s() - t0;
}
