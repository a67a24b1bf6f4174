/*
 * kblas_oracle_impl.h -- precision-generic body of the CPU oracle (included twice by
 * kblas_oracle.c with ORA_T = float / double and ORA_(name) = name##_s / name##_d).
 *
 * TEST INFRASTRUCTURE ONLY.  See kblas_oracle.c for the rules about who may use it.
 *
 * Every routine restates, element by element, the arithmetic one matrix goes through in the
 * reference GPU path: same recursion, same blocking, same order of fma / division / sqrt,
 * so that for the parts of the path that do not go through cuBLAS (potrf n <= 32, trsm /
 * syrk kernels up to 16) the oracle reproduces the reference's roundings.  Where the
 * reference calls cuBLAS batched GEMM (closed source) the oracle uses a k-sequential fma dot
 * product and parity is tolerance-based.
 */

#define A_(i, j) A[(size_t)(i) + (size_t)(j) * lda]
#define B_(i, j) B[(size_t)(i) + (size_t)(j) * ldb]
#define C_(i, j) C[(size_t)(i) + (size_t)(j) * ldc]

/* ---- unblocked right-looking Cholesky of an n x n block, n <= 8.
 * Reference: dev_potrf_U_registers_fixN / _varN, Xpotrf_batch_kernels.cuh:37-78, 104-150:
 *   s = sqrt(a_jj); column j /= s (division, rows >= j incl. the diagonal: a_jj = a_jj / s);
 *   a_ik = fma(a_ij, -a_kj, a_ik) for k > j, rows i >= k. */
static void ORA_(potrf_unblocked)(int n, ORA_T *A, int lda) {
  for (int j = 0; j < n; j++) {
    ORA_T s = ORA_SQRT(A_(j, j));
    for (int i = j; i < n; i++) A_(i, j) = A_(i, j) / s;
    for (int k = j + 1; k < n; k++) {
      ORA_T nk = -A_(k, j);
      for (int i = k; i < n; i++) A_(i, k) = ORA_FMA(A_(i, j), nk, A_(i, k));
    }
  }
}

/* ---- 2x2-blocked Cholesky (block 8) of an n x n matrix, 8 < n <= 16.
 * Reference: dev_potrf_U_registers_fixN_blocked_2 / _varN_blocked_2,
 * Xpotrf_batch_kernels.cuh:176-300, 329-459:
 *   1. potrf A00 (8x8)                         (:190-209)
 *   2. A10 := A10 * A00^-T row by row:        (:228-238)  r_k /= a_kk ; r_j = fma(-a_jk, r_k, r_j), j > k
 *   3. A11 -= A10 * A10^T with  s = sum_i fma(r_i, q_i, s) from 0, then a -= s   (:257-266)
 *   4. potrf A11                               (:270-290) */
static void ORA_(potrf_blocked2)(int n, ORA_T *A, int lda) {
  const int BS = 8, n2 = n - BS;
  ORA_(potrf_unblocked)(BS, A, lda);
  for (int r = 0; r < n2; r++) {
    for (int k = 0; k < BS; k++) {
      A_(BS + r, k) = A_(BS + r, k) / A_(k, k);
      for (int j = k + 1; j < BS; j++) A_(BS + r, j) = ORA_FMA(-A_(j, k), A_(BS + r, k), A_(BS + r, j));
    }
  }
  for (int j = 0; j < n2; j++)
    for (int r = j; r < n2; r++) {
      ORA_T s = 0;
      for (int i = 0; i < BS; i++) s = ORA_FMA(A_(BS + r, i), A_(BS + j, i), s);
      A_(BS + r, BS + j) = A_(BS + r, BS + j) - s;
    }
  ORA_(potrf_unblocked)(n2, &A_(BS, BS), lda);
}

/* ---- GEMM as the reference gets it from cuBLAS (Xgemm_batch_core.cuh:260-267, 549-556):
 * C(m x n) = alpha * op(A) * op(B) + beta * C, inner dimension k.  cuBLAS's accumulation
 * order is not observable; the oracle uses a k-sequential fma chain. */
static void ORA_(gemm)(int transA, int transB, int m, int n, int k, ORA_T alpha, const ORA_T *A, int lda,
                       const ORA_T *B, int ldb, ORA_T beta, ORA_T *C, int ldc) {
  for (int j = 0; j < n; j++)
    for (int i = 0; i < m; i++) {
      ORA_T s = 0;
      for (int l = 0; l < k; l++) {
        ORA_T a = transA ? A_(l, i) : A_(i, l);
        ORA_T b = transB ? B_(j, l) : B_(l, j);
        s = ORA_FMA(a, b, s);
      }
      C_(i, j) = ORA_FMA(alpha, s, beta * C_(i, j));
    }
}

/* ---- SYRK trailing update C(m x m, lower) = alpha * A(m x n) * A^T + beta * C.
 * Reference: Xsyrk_batch_strided_core / Xsyrk_batch_core, Xsyrk_batch_drivers.cuh:32-228:
 * diagonal 16x16 blocks by the register kernels K11-K14 (Xsyrk_batch_kernels.cuh:116-212,
 * 423-543): C is first scaled by beta, then for every chunk of 8 columns of A
 *   s = sum_{i in chunk} fma(a_ri, a_ci, s) from 0 ;  c = fma(alpha, s, c);
 * everything below the diagonal blocks goes through batched GEMM (:234-323). */
static void ORA_(syrk)(int m, int n, ORA_T alpha, const ORA_T *A, int lda, ORA_T beta, ORA_T *C, int ldc) {
  for (int c = 0; c < m; c++)
    for (int r = c; r < m; r++) {
      if (r / 16 == c / 16) {
        ORA_T v = C_(r, c) * beta;
        for (int b = 0; b < n; b += 8) {
          ORA_T s = 0;
          for (int i = b; i < b + 8 && i < n; i++) s = ORA_FMA(A_(r, i), A_(c, i), s);
          v = ORA_FMA(alpha, s, v);
        }
        C_(r, c) = v;
      } else {
        ORA_T s = 0;
        for (int i = 0; i < n; i++) s = ORA_FMA(A_(r, i), A_(c, i), s);
        C_(r, c) = ORA_FMA(alpha, s, beta * C_(r, c));
      }
    }
}

static int ORA_(reg_size)(int n) { return (n > 0) && !(n & (n - 1)); }
static int ORA_(closest_reg_size)(int n) {
  if (n <= 0) return 0;
  int r = 1;
  while (r < n) r <<= 1;
  return r >> 1;
}
/* split rule shared by every driver (Xpotrf_batch_drivers.cuh:94-101, kblas_common.cu:241-255) */
static void ORA_(split)(int n, int *n1, int *n2) {
  if (ORA_(reg_size)(n)) *n1 = *n2 = n / 2;
  else { *n1 = ORA_(closest_reg_size)(n); *n2 = n - *n1; }
}

/* ---- TRSM, A lower non-unit, one problem.  Reference: Xtrsm_batch_core,
 * Xtrsm_batch_drivers.cuh:54-272 (recursion) + kernels K5-K9 (Xtrsm_batch_kernels.cuh):
 *  R/T (:36-101)  rows of B: for j up:  b_j /= a_jj ; b_i = fma(b_j, -a_ij, b_i), i > j ; store alpha*b
 *  R/N (:36-101)  b = alpha*b ; for j down: b_j = fma(b_i, -a_ij, b_j), i > j (i ascending) ; b_j /= a_jj
 *  L/N (:590-604) columns of B: b = alpha*b ; for j up: b_j /= a_jj ; b_i = fma(-a_ij, b_j, b_i), i > j
 *  L/T (:605-618) b = alpha*b ; for j down: b_j = fma(-a_ij, b_i, b_j), i > j (i ascending) ; b_j /= a_jj
 * Above 16 the driver recurses with a GEMM in between (:127-266), including the
 * -1/alpha trick of the R/T branch (:154-163). */
static void ORA_(trsm)(char side, char trans, int m, int n, ORA_T alpha, const ORA_T *A, int lda, ORA_T *B, int ldb) {
  const int left = (side == 'L'), tr = (trans == 'T');
  const int k = left ? m : n;
  if (k <= 0) return;
  if (k <= 16) {
    if (!left) {
      for (int r = 0; r < m; r++) {
        if (tr) {
          for (int j = 0; j < n; j++) {
            B_(r, j) = B_(r, j) / A_(j, j);
            for (int i = j + 1; i < n; i++) B_(r, i) = ORA_FMA(B_(r, j), -A_(i, j), B_(r, i));
          }
          for (int j = 0; j < n; j++) B_(r, j) = alpha * B_(r, j);
        } else {
          for (int j = 0; j < n; j++) B_(r, j) = alpha * B_(r, j);
          for (int j = n - 1; j >= 0; j--) {
            for (int i = j + 1; i < n; i++) B_(r, j) = ORA_FMA(B_(r, i), -A_(i, j), B_(r, j));
            B_(r, j) = B_(r, j) / A_(j, j);
          }
        }
      }
    } else {
      for (int c = 0; c < n; c++) {
        for (int i = 0; i < m; i++) B_(i, c) = alpha * B_(i, c);
        if (!tr) {
          for (int j = 0; j < m; j++) {
            B_(j, c) = B_(j, c) / A_(j, j);
            for (int i = j + 1; i < m; i++) B_(i, c) = ORA_FMA(-A_(i, j), B_(j, c), B_(i, c));
          }
        } else {
          for (int j = m - 1; j >= 0; j--) {
            for (int i = j + 1; i < m; i++) B_(j, c) = ORA_FMA(-A_(i, j), B_(i, c), B_(j, c));
            B_(j, c) = B_(j, c) / A_(j, j);
          }
        }
      }
    }
    return;
  }
  int k1, k2;
  ORA_(split)(k, &k1, &k2);
  const ORA_T one = 1, mone = -1;
  if (!left) {
    if (tr) { /* :140-171 */
      ORA_(trsm)(side, trans, m, k1, alpha, A, lda, B, ldb);
      ORA_(gemm)(0, 1, m, k2, k1, mone / alpha, B, ldb, &A_(k1, 0), lda, one, &B_(0, k1), ldb);
      ORA_(trsm)(side, trans, m, k2, alpha, &A_(k1, k1), lda, &B_(0, k1), ldb);
    } else { /* :173-199 */
      ORA_(trsm)(side, trans, m, k2, alpha, &A_(k1, k1), lda, &B_(0, k1), ldb);
      ORA_(gemm)(0, 0, m, k1, k2, mone, &B_(0, k1), ldb, &A_(k1, 0), lda, alpha, B, ldb);
      ORA_(trsm)(side, trans, m, k1, one, A, lda, B, ldb);
    }
  } else {
    if (tr) { /* :213-238 */
      ORA_(trsm)(side, trans, k2, n, alpha, &A_(k1, k1), lda, &B_(k1, 0), ldb);
      ORA_(gemm)(1, 0, k1, n, k2, mone, &A_(k1, 0), lda, &B_(k1, 0), ldb, alpha, B, ldb);
      ORA_(trsm)(side, trans, k1, n, one, A, lda, B, ldb);
    } else { /* :240-265 */
      ORA_(trsm)(side, trans, k1, n, alpha, A, lda, B, ldb);
      ORA_(gemm)(0, 0, k2, n, k1, mone, &A_(k1, 0), lda, B, ldb, alpha, &B_(k1, 0), ldb);
      ORA_(trsm)(side, trans, k2, n, one, &A_(k1, k1), lda, &B_(k1, 0), ldb);
    }
  }
}

/* ---- POTRF driver recursion, one matrix.  Reference: Xpotrf_batch_core,
 * Xpotrf_batch_drivers.cuh:30-137: n <= 16 -> register kernels; otherwise
 * potrf(n1) -> trsm(R,L,T, n2 x n1, alpha 1) -> syrk(L,N, n2, n1, -1, 1) -> potrf(n2). */
static void ORA_(potrf)(int n, ORA_T *A, int lda) {
  if (n <= 0) return;
  if (n <= 8) { ORA_(potrf_unblocked)(n, A, lda); return; }
  if (n <= 16) { ORA_(potrf_blocked2)(n, A, lda); return; }
  int n1, n2;
  ORA_(split)(n, &n1, &n2);
  ORA_(potrf)(n1, A, lda);
  ORA_(trsm)('R', 'T', n2, n1, (ORA_T)1, A, lda, &A_(n1, 0), lda);
  ORA_(syrk)(n2, n1, (ORA_T)-1, &A_(n1, 0), lda, (ORA_T)1, &A_(n1, n1), lda);
  ORA_(potrf)(n2, &A_(n1, n1), lda);
}

/* ---- POTRS (side R): X (L L^T) = B.  Reference: Xpotrs_batch_core,
 * Xpotrs_batch_drivers.cuh:80-171 -- always the split form, even for n <= 16:
 *   trsm(R,L,T, m x n1) ; B1 -= B0 A10^T ; trsm(R,L,T, m x n2) ; trsm(R,L,N, m x n2) ;
 *   B0 -= B1 A10 ; trsm(R,L,N, m x n1).
 * n == 1 gives n1 = 0 and the reference fails with KBLAS_NotImplemented (:85-98 with
 * Xtrsm_batch_drivers.cuh:267-270); the oracle solves it as the obvious scalar problem. */
static void ORA_(potrs)(int m, int n, const ORA_T *A, int lda, ORA_T *B, int ldb) {
  const ORA_T one = 1, mone = -1;
  if (n == 1) {
    for (int r = 0; r < m; r++) B_(r, 0) = (B_(r, 0) / A_(0, 0)) / A_(0, 0);
    return;
  }
  int n1, n2;
  ORA_(split)(n, &n1, &n2);
  ORA_(trsm)('R', 'T', m, n1, one, A, lda, B, ldb);
  ORA_(gemm)(0, 1, m, n2, n1, mone, B, ldb, &A_(n1, 0), lda, one, &B_(0, n1), ldb);
  ORA_(trsm)('R', 'T', m, n2, one, &A_(n1, n1), lda, &B_(0, n1), ldb);
  ORA_(trsm)('R', 'N', m, n2, one, &A_(n1, n1), lda, &B_(0, n1), ldb);
  ORA_(gemm)(0, 0, m, n1, n2, mone, &B_(0, n1), ldb, &A_(n1, 0), lda, one, B, ldb);
  ORA_(trsm)('R', 'N', m, n1, one, A, lda, B, ldb);
}

/* =============================== batch entry points =============================== */

/* kblas{S,D}potrf_batch_strided (Xpotrf_batch.cu:107-160): lower only; info untouched. */
int ORA_(oracle_potrf_batch_strided)(char uplo, int n, ORA_T *A, int lda, long strideA, int batchCount) {
  if (uplo == 'U') return -2;
  for (long b = 0; b < batchCount; b++) ORA_(potrf)(n, A + b * strideA, lda);
  return 1;
}

/* kblas{S,D}trsm_batch_strided (Xtrsm_batch.cu:189-257) */
int ORA_(oracle_trsm_batch_strided)(char side, char uplo, char trans, char diag, int m, int n, ORA_T alpha,
                                    const ORA_T *A, int lda, long strideA, ORA_T *B, int ldb, long strideB,
                                    int batchCount) {
  if (uplo == 'U' || diag == 'U') return -2;
  if ((side == 'L' ? m : n) <= 0) return -2;
  for (long b = 0; b < batchCount; b++) ORA_(trsm)(side, trans, m, n, alpha, A + b * strideA, lda, B + b * strideB, ldb);
  return 1;
}

/* kblas{S,D}potrs_batch_strided (Xpotrs_batch.cu:104-164) */
int ORA_(oracle_potrs_batch_strided)(char side, char uplo, int m, int n, const ORA_T *A, int lda, long strideA,
                                     ORA_T *B, int ldb, long strideB, int batchCount) {
  if (side == 'L' || uplo == 'U') return -2;
  for (long b = 0; b < batchCount; b++) ORA_(potrs)(m, n, A + b * strideA, lda, B + b * strideB, ldb);
  return 1;
}

/* kblas{S,D}posv_batch_strided (Xposv_batch.cu:112-178, Xposv_batch_drivers.cuh:84-114) */
int ORA_(oracle_posv_batch_strided)(char side, char uplo, int m, int n, ORA_T *A, int lda, long strideA, ORA_T *B,
                                    int ldb, long strideB, int batchCount) {
  if (side == 'L' || uplo == 'U') return -2;
  for (long b = 0; b < batchCount; b++) {
    ORA_(potrf)(n, A + b * strideA, lda);
    ORA_(potrs)(m, n, A + b * strideA, lda, B + b * strideB, ldb);
  }
  return 1;
}

/* ---- side = 'L' EXTENSION: (L L^T) X = B, A of order m, B m x n.  NOT in the reference -- its potrs /
 * posv answer KBLAS_NotImplemented for side L (Xpotrs_batch_drivers.cuh:40-43, Xposv_batch_drivers.cuh:41-44), which is
 * what the two functions above restate.  The extension is DEFINED as the composition of the restated reference TRSMs:
 *   L Y = B  (trsm L,L,N)  then  L^T X = Y  (trsm L,L,T);  posv = restated potrf of order m, then that. */
int ORA_(oracle_potrs_left_batch_strided)(int m, int n, const ORA_T *A, int lda, long strideA, ORA_T *B, int ldb,
                                          long strideB, int batchCount) {
  for (long b = 0; b < batchCount; b++) {
    ORA_(trsm)('L', 'N', m, n, (ORA_T)1, A + b * strideA, lda, B + b * strideB, ldb);
    ORA_(trsm)('L', 'T', m, n, (ORA_T)1, A + b * strideA, lda, B + b * strideB, ldb);
  }
  return 1;
}
int ORA_(oracle_posv_left_batch_strided)(int m, int n, ORA_T *A, int lda, long strideA, ORA_T *B, int ldb, long strideB,
                                         int batchCount) {
  for (long b = 0; b < batchCount; b++) ORA_(potrf)(m, A + b * strideA, lda);
  return ORA_(oracle_potrs_left_batch_strided)(m, n, A, lda, strideA, B, ldb, strideB, batchCount);
}

/* kblas{S,D}gemm_batch_strided (Xgemm_batch.cu:298-363 -> cuBLAS batched GEMM, Xgemm_batch_core.cuh:549-556) and
 * kblas{S,D}syrk_batch_strided (Xsyrk_batch.cu:137-189 -> Xsyrk_batch_strided_core, drivers.cuh:125-228; trans = 'T' reads
 * A as n x m).  cuBLAS's accumulation order is not observable: k-sequential fma, tolerance-based parity. */
int ORA_(oracle_gemm_batch_strided)(char transA, char transB, int m, int n, int k, ORA_T alpha, const ORA_T *A, int lda,
                                    long strideA, const ORA_T *B, int ldb, long strideB, ORA_T beta, ORA_T *C, int ldc,
                                    long strideC, int batchCount) {
  if (batchCount < 1) return -10; /* KBLAS_Error_WrongInput, Xgemm_batch_core.cuh:181-182 */
  for (long b = 0; b < batchCount; b++)
    ORA_(gemm)(transA == 'T' || transA == 't', transB == 'T' || transB == 't', m, n, k, alpha, A + b * strideA, lda,
               B + b * strideB, ldb, beta, C + b * strideC, ldc);
  return 1;
}
int ORA_(oracle_syrk_batch_strided)(char uplo, char trans, int m, int n, ORA_T alpha, const ORA_T *A, int lda, long strideA,
                                    ORA_T beta, ORA_T *B, int ldb, long strideB, int batchCount) {
  if (uplo == 'U' || uplo == 'u') return -2; /* Xsyrk_batch_drivers.cuh:133-136 */
  for (long b = 0; b < batchCount; b++) {
    const ORA_T *Ab = A + b * strideA;
    ORA_T *C = B + b * strideB;
    const int ldc = ldb;
    if (trans == 'T' || trans == 't') {
      for (int c = 0; c < m; c++)
        for (int r = c; r < m; r++) {
          ORA_T s = 0;
          for (int i = 0; i < n; i++) s = ORA_FMA(Ab[(size_t)i + (size_t)r * lda], Ab[(size_t)i + (size_t)c * lda], s);
          C_(r, c) = ORA_FMA(alpha, s, beta * C_(r, c));
        }
    } else {
      ORA_(syrk)(m, n, alpha, Ab, lda, beta, C, ldc);
    }
  }
  return 1;
}

/* ---- the consumers of the factor: kblas{S,D}{trtri,lauum,potri,poti}_batch_strided
 * (reference src/batch_triangular/X{trtri,lauum,potri,poti}_batch.cu + drivers: trtri = register kernels up to 16 + two
 * TRSMs per recursion level, Xtrtri_batch_drivers.cuh:31-125; lauum = register kernels + TRMM + SYRK,
 * Xlauum_batch_drivers.cuh; potri = trtri then lauum, Xpotri_batch_drivers.cuh; poti = potrf then potri,
 * Xpoti_batch_drivers.cuh:82-89).  The recursion's operation order goes through cuBLAS GEMM for n > 16, so parity is
 * tolerance-based; the oracle states the definitions with k-sequential fma chains:
 *   trtri: column j of X = L^-1 by forward substitution of e_j;  lauum: r_ij = sum_{k>=i} l_ki l_kj, i >= j. */
static void ORA_(trtri)(int n, ORA_T *A, int lda) {
  static ORA_T X[256 * 256];
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) X[i + (size_t)j * n] = (i == j) ? (ORA_T)1 : (ORA_T)0;
    for (int k = j; k < n; k++) {
      X[k + (size_t)j * n] = X[k + (size_t)j * n] / A_(k, k);
      for (int i = k + 1; i < n; i++) X[i + (size_t)j * n] = ORA_FMA(-A_(i, k), X[k + (size_t)j * n], X[i + (size_t)j * n]);
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) A_(i, j) = X[i + (size_t)j * n];
}
static void ORA_(lauum)(int n, ORA_T *A, int lda) {
  static ORA_T R[256 * 256];
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) {
      ORA_T s = 0;
      for (int k = i; k < n; k++) s = ORA_FMA(A_(k, i), A_(k, j), s);
      R[i + (size_t)j * n] = s;
    }
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) A_(i, j) = R[i + (size_t)j * n];
}
/* which: 0 trtri, 1 lauum, 2 potri, 3 poti */
int ORA_(oracle_inv_batch_strided)(int which, char uplo, char diag, int n, ORA_T *A, int lda, long strideA, int batchCount) {
  if (uplo == 'U' || diag == 'U') return -2; /* Xtrtri_batch_drivers.cuh:96-99 and siblings */
  if (n > 256) return -2;
  for (long b = 0; b < batchCount; b++) {
    ORA_T *Ab = A + b * strideA;
    if (which == 3) ORA_(potrf)(n, Ab, lda);
    if (which == 0 || which == 2 || which == 3) ORA_(trtri)(n, Ab, lda);
    if (which == 1 || which == 2 || which == 3) ORA_(lauum)(n, Ab, lda);
  }
  return 1;
}

/* ---- packed lower storage (LAPACK ?pptrf, uplo = 'L'): AP[j*n - j(j-1)/2 + (i-j)] = A(i,j), i >= j.
 * The reference has no packed batch routine (its batch_pstrf, src/batch_svd/batch_pstrf.cu:226-246,
 * is pivoted Cholesky on full storage), so the oracle for kblasx?pptrf_batch is DEFINED as: unpack, factor with the
 * restated reference potrf above (same recursion and roundings, Xpotrf_batch_drivers.cuh:30-137), re-pack.  Parity
 * anchor: pptrf(pack(A)) == pack(potrf(A)) with potrf pinned by the reference's golden vectors. */
int ORA_(oracle_pptrf_batch_strided)(char uplo, int n, ORA_T *AP, long strideAP, int batchCount) {
  if (uplo != 'L' && uplo != 'l') return -2; /* KBLAS_NotImplemented */
  if (n <= 0) return 1;
  if (n > 256) return -2;
  static ORA_T W[256 * 256];
  const int lda = n;
  for (long b = 0; b < batchCount; b++) {
    ORA_T *P = AP + b * strideAP;
    ORA_T *A = W;
    size_t e = 0;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) A_(i, j) = (i >= j) ? P[e++] : (ORA_T)0;
    ORA_(potrf)(n, A, lda);
    e = 0;
    for (int j = 0; j < n; j++)
      for (int i = j; i < n; i++) P[e++] = A_(i, j);
  }
  return 1;
}


#undef A_
#undef B_
#undef C_
