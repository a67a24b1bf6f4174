/*
 * kblas_oracle.c -- CPU restatement of the reference KBLAS-GPU batched Cholesky path.
 *
 * *** TEST INFRASTRUCTURE ONLY ***
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call anything under oracle/ -- and there only as the checker or the
 * reported CPU baseline, never as the thing measured or shipped.  The product
 * (kblas-gpu_b200/) never links, imports or falls back to this code.
 *
 * What it is: plain C, one matrix at a time, following the reference's algorithm for
 *   kblas{S,D}potrf_batch_strided, trsm_batch_strided, potrs_batch_strided, posv_batch_strided
 * (ecrc/kblas-gpu src/batch_triangular/X{potrf,trsm,syrk,potrs,posv}_batch_{drivers,kernels}.cuh;
 * each function in kblas_oracle_impl.h cites the file:line it follows).
 *
 * Pinning: the reference ships no golden vectors and its tests assert nothing (SURVEY.md §0
 * finding 9), so the oracle is pinned against outputs of the REFERENCE ITSELF: the unmodified
 * reference sources are compiled into oracle/_ref/libkblas_ref.so (oracle/build_ref.sh), run on
 * a B200 by tests/golden/make_golden.py, and the resulting input/output pairs are committed
 * under tests/golden/.  tests/test_oracle.py checks this file against them (bit-exact where
 * the reference path has no cuBLAS call, tolerance-based elsewhere) and against LAPACK.
 *
 * Build: make -C oracle        (gcc -O2 -ffp-contract=off: every fma is an explicit fma())
 */
#include <math.h>
#include <stddef.h>

#define ORA_CAT_(a, b) a##b
#define ORA_CAT(a, b) ORA_CAT_(a, b)

#define ORA_T double
#define ORA_(name) ORA_CAT(name, _d)
#define ORA_SQRT(x) sqrt(x)
#define ORA_FMA(a, b, c) fma((a), (b), (c))
#include "kblas_oracle_impl.h"
#undef ORA_T
#undef ORA_
#undef ORA_SQRT
#undef ORA_FMA

#define ORA_T float
#define ORA_(name) ORA_CAT(name, _s)
#define ORA_SQRT(x) sqrtf(x)
#define ORA_FMA(a, b, c) fmaf((a), (b), (c))
#include "kblas_oracle_impl.h"
#undef ORA_T
#undef ORA_
#undef ORA_SQRT
#undef ORA_FMA

const char *oracle_version(void) { return "kblas-oracle 1 (restates KBLAS-GPU 3.0.0 batch potrf/trsm/potrs/posv)"; }
