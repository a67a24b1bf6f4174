/* mkl.h -- SHIM, test infrastructure only (oracle/): lets the reference's own test programs
 * (testing/batch_triangular/test_X{potrf,trsm,potrs,posv}_batch.cpp, compiled with -DUSE_MKL as the authors do,
 * make.inc:28-29) build in an image that has no MKL: the handful of Fortran-style BLAS/LAPACK names those programs
 * call are forwarded to the LP64 OpenBLAS that ships inside scipy (symbols scipy_<name>_).  Nothing here is
 * reference code; the prototypes are the standard BLAS/LAPACK Fortran interfaces. */
#ifndef KBLAS_B200_SHIM_MKL_H
#define KBLAS_B200_SHIM_MKL_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int lapack_int;
typedef int MKL_INT;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

#define KX_SHIM_DECL(p, T)                                                                                         \
  void scipy_##p##potrf_(const char *uplo, const int *n, T *a, const int *lda, int *info);                         \
  void scipy_##p##axpy_(const int *n, const T *alpha, const T *x, const int *incx, T *y, const int *incy);         \
  T scipy_##p##lansy_(const char *norm, const char *uplo, const int *n, const T *a, const int *lda, T *work);      \
  T scipy_##p##lange_(const char *norm, const int *m, const int *n, const T *a, const int *lda, T *work);          \
  void scipy_##p##trsm_(const char *side, const char *uplo, const char *transa, const char *diag, const int *m,    \
                        const int *n, const T *alpha, const T *a, const int *lda, T *b, const int *ldb);           \
  lapack_int scipy_LAPACKE_##p##latms(int matrix_layout, lapack_int m, lapack_int n, char dist, lapack_int *iseed, \
                                      char sym, T *d, lapack_int mode, T cond, T dmax, lapack_int kl,              \
                                      lapack_int ku, char pack, T *a, lapack_int lda);                             \
  static inline void p##potrf(const char *uplo, const int *n, T *a, const int *lda, int *info) {                   \
    scipy_##p##potrf_(uplo, n, a, lda, info);                                                                      \
  }                                                                                                                \
  static inline void p##axpy(const int *n, const T *alpha, const T *x, const int *incx, T *y, const int *incy) {   \
    scipy_##p##axpy_(n, alpha, x, incx, y, incy);                                                                  \
  }                                                                                                                \
  static inline T p##lansy(const char *norm, const char *uplo, const int *n, const T *a, const int *lda, T *work) { \
    return scipy_##p##lansy_(norm, uplo, n, a, lda, work);                                                         \
  }                                                                                                                \
  static inline T p##lange(const char *norm, const int *m, const int *n, const T *a, const int *lda, T *work) {    \
    return scipy_##p##lange_(norm, m, n, a, lda, work);                                                            \
  }                                                                                                                \
  static inline void p##trsm(const char *side, const char *uplo, const char *transa, const char *diag,             \
                             const int *m, const int *n, const T *alpha, const T *a, const int *lda, T *b,         \
                             const int *ldb) {                                                                     \
    scipy_##p##trsm_(side, uplo, transa, diag, m, n, alpha, a, lda, b, ldb);                                       \
  }                                                                                                                \
  static inline lapack_int LAPACKE_##p##latms(int matrix_layout, lapack_int m, lapack_int n, char dist,            \
                                              lapack_int *iseed, char sym, T *d, lapack_int mode, T cond, T dmax,  \
                                              lapack_int kl, lapack_int ku, char pack, T *a, lapack_int lda) {     \
    return scipy_LAPACKE_##p##latms(matrix_layout, m, n, dist, iseed, sym, d, mode, cond, dmax, kl, ku, pack, a,   \
                                    lda);                                                                          \
  }

KX_SHIM_DECL(s, float)
KX_SHIM_DECL(d, double)
#undef KX_SHIM_DECL

void scipy_openblas_set_num_threads(int);

#ifdef __cplusplus
}
#endif
#endif
