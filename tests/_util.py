"""Shared test plumbing: the CPU oracle, LAPACK, the reference GPU library, data generators.

Everything under oracle/ is TEST INFRASTRUCTURE: it is loaded here (and by bench.py's baseline
legs and __graft_entry__.smoke()) as the checker only.
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libkblas_oracle.so")
LAPACK_LOOP_SO = os.path.join(ORACLE_DIR, "liblapack_loop.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libkblas_ref.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

EPS = {np.float32: float(np.finfo(np.float32).eps), np.float64: float(np.finfo(np.float64).eps)}
SUFFIX = {np.float32: "s", np.float64: "d"}
CT = {np.float32: C.c_float, np.float64: C.c_double}


def kblas():
    """the product package (its directory name has a hyphen)"""
    return importlib.import_module("kblas-gpu_b200")


# --------------------------------------------------------------------------------------------
# CPU oracle (oracle/kblas_oracle.c)
_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libkblas_oracle.so"], stdout=subprocess.DEVNULL)
        _oracle = C.CDLL(ORACLE_SO)
        i, l, c, P = C.c_int, C.c_long, C.c_char, C.c_void_p
        for s, t in (("s", C.c_float), ("d", C.c_double)):
            getattr(_oracle, f"oracle_potrf_batch_strided_{s}").argtypes = [c, i, P, i, l, i]
            getattr(_oracle, f"oracle_trsm_batch_strided_{s}").argtypes = [c, c, c, c, i, i, t, P, i, l, P, i, l, i]
            getattr(_oracle, f"oracle_potrs_batch_strided_{s}").argtypes = [c, c, i, i, P, i, l, P, i, l, i]
            getattr(_oracle, f"oracle_posv_batch_strided_{s}").argtypes = [c, c, i, i, P, i, l, P, i, l, i]
            getattr(_oracle, f"oracle_pptrf_batch_strided_{s}").argtypes = [c, i, P, l, i]
            getattr(_oracle, f"oracle_inv_batch_strided_{s}").argtypes = [i, c, c, i, P, i, l, i]
            getattr(_oracle, f"oracle_gemm_batch_strided_{s}").argtypes = [c, c, i, i, i, t, P, i, l, P, i, l, t, P, i, l, i]
            getattr(_oracle, f"oracle_syrk_batch_strided_{s}").argtypes = [c, c, i, i, t, P, i, l, t, P, i, l, i]
            getattr(_oracle, f"oracle_potrs_left_batch_strided_{s}").argtypes = [i, i, P, i, l, P, i, l, i]
            getattr(_oracle, f"oracle_posv_left_batch_strided_{s}").argtypes = [i, i, P, i, l, P, i, l, i]
    return _oracle


def _np_ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _dt(a):
    return np.float64 if a.dtype == np.float64 else np.float32


# Batches are numpy arrays of shape (batch, ncols, ld): matrix b, element (i, j) at [b, j, i]
# (i.e. column-major matrices with leading dimension ld, stride ncols*ld) -- exactly the strided
# layout of the KBLAS API.

def oracle_potrf(A, n, uplo="L"):
    """in place on A[(batch, ncols, lda)]; returns the reference-style return code"""
    b, nc, lda = A.shape
    f = getattr(oracle(), f"oracle_potrf_batch_strided_{SUFFIX[_dt(A)]}")
    return f(uplo.encode(), n, _np_ptr(A), lda, nc * lda, b)


def oracle_pptrf(AP, n, uplo="L"):
    """in place on AP[(batch, strideAP)]: packed lower storage, oracle/kblas_oracle_impl.h oracle_pptrf_batch_strided"""
    b, stride = AP.shape
    f = getattr(oracle(), f"oracle_pptrf_batch_strided_{SUFFIX[_dt(AP)]}")
    return f(uplo.encode(), n, _np_ptr(AP), stride, b)


def oracle_trsm(side, uplo, trans, diag, m, n, alpha, A, B):
    b, nca, lda = A.shape
    _, ncb, ldb = B.shape
    f = getattr(oracle(), f"oracle_trsm_batch_strided_{SUFFIX[_dt(B)]}")
    return f(side.encode(), uplo.encode(), trans.encode(), diag.encode(), m, n, alpha, _np_ptr(A), lda, nca * lda,
             _np_ptr(B), ldb, ncb * ldb, b)


def oracle_potrs(side, uplo, m, n, A, B):
    b, nca, lda = A.shape
    _, ncb, ldb = B.shape
    f = getattr(oracle(), f"oracle_potrs_batch_strided_{SUFFIX[_dt(B)]}")
    return f(side.encode(), uplo.encode(), m, n, _np_ptr(A), lda, nca * lda, _np_ptr(B), ldb, ncb * ldb, b)


def oracle_posv(side, uplo, m, n, A, B):
    b, nca, lda = A.shape
    _, ncb, ldb = B.shape
    f = getattr(oracle(), f"oracle_posv_batch_strided_{SUFFIX[_dt(B)]}")
    return f(side.encode(), uplo.encode(), m, n, _np_ptr(A), lda, nca * lda, _np_ptr(B), ldb, ncb * ldb, b)


def oracle_gemm(transA, transB, m, n, k, alpha, A, B, beta, Cm):
    """C := alpha op(A) op(B) + beta C on (batch, ncols, ld) strided arrays"""
    b, nca, lda = A.shape
    _, ncb, ldb = B.shape
    _, ncc, ldc = Cm.shape
    f = getattr(oracle(), f"oracle_gemm_batch_strided_{SUFFIX[_dt(Cm)]}")
    return f(transA.encode(), transB.encode(), m, n, k, alpha, _np_ptr(A), lda, nca * lda, _np_ptr(B), ldb, ncb * ldb, beta,
             _np_ptr(Cm), ldc, ncc * ldc, b)


def oracle_syrk(uplo, trans, m, n, alpha, A, beta, Cm):
    b, nca, lda = A.shape
    _, ncc, ldc = Cm.shape
    f = getattr(oracle(), f"oracle_syrk_batch_strided_{SUFFIX[_dt(Cm)]}")
    return f(uplo.encode(), trans.encode(), m, n, alpha, _np_ptr(A), lda, nca * lda, beta, _np_ptr(Cm), ldc, ncc * ldc, b)


def oracle_inv(which, n, A, uplo="L", diag="N"):
    """in place on A[(batch, ncols, lda)]: which = 'trtri' | 'lauum' | 'potri' | 'poti' (lower triangle)"""
    b, nc, lda = A.shape
    f = getattr(oracle(), f"oracle_inv_batch_strided_{SUFFIX[_dt(A)]}")
    return f({"trtri": 0, "lauum": 1, "potri": 2, "poti": 3}[which], uplo.encode(), diag.encode(), n, _np_ptr(A), lda, nc * lda, b)


def oracle_posv_left(m, n, A, B):
    """side = 'L' extension (A of order m, A X = B): restated potrf + restated trsm L,L,N + trsm L,L,T"""
    b, nca, lda = A.shape
    _, ncb, ldb = B.shape
    f = getattr(oracle(), f"oracle_posv_left_batch_strided_{SUFFIX[_dt(B)]}")
    return f(m, n, _np_ptr(A), lda, nca * lda, _np_ptr(B), ldb, ncb * ldb, b)


# --------------------------------------------------------------------------------------------
# LAPACK (OpenBLAS bundled with scipy, LP64, symbols scipy_<name>_)
_lapack = None


def lapack_lib():
    global _lapack
    if _lapack is None:
        import scipy

        libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        if not libs:
            raise RuntimeError("scipy's bundled OpenBLAS not found")
        _lapack = C.CDLL(libs[0])
    return _lapack


# --------------------------------------------------------------------------------------------
# data generators -- the reference harness's distributions (testing/testing_helper.cu:353-402):
# uniform [0,1) entries, A made SPD by  a_ii += n  and mirroring the lower triangle.

def rand_spd_batch(batch, n, lda=None, dtype=np.float64, seed=1, extra_cols=0):
    """(batch, n + extra_cols, lda) array; leading n x n of each matrix SPD, padding filled with a sentinel"""
    lda = lda or n
    rng = np.random.default_rng(seed)
    A = np.full((batch, n + extra_cols, lda), -7.25, dtype=dtype)  # sentinel in padding
    M = rng.random((batch, n, n)).astype(dtype)
    M = np.tril(M) + np.transpose(np.tril(M, -1), (0, 2, 1))
    M[:, np.arange(n), np.arange(n)] += n
    A[:, :n, :n] = M  # symmetric, so [b, j, i] == [b, i, j]
    return A


def rand_batch(batch, rows, cols, ld=None, dtype=np.float64, seed=2):
    """(batch, cols, ld) column-major rows x cols matrices of uniform [0,1)"""
    ld = ld or rows
    rng = np.random.default_rng(seed)
    B = np.full((batch, cols, ld), -3.5, dtype=dtype)
    B[:, :, :rows] = rng.random((batch, cols, rows)).astype(dtype)
    return B


def hilbert_spd_batch(batch, n, dtype=np.float64, delta=1e-3):
    """Hilbert-like ill-conditioned SPD matrices (testing_helper.h:313-320) + delta*I"""
    i = np.arange(n)
    H = 1.0 / (i[:, None] + i[None, :] + 1.0)
    A = np.empty((batch, n, n), dtype=dtype)
    for b in range(batch):
        A[b] = (H * (1.0 + 0.01 * b) + delta * np.eye(n)).astype(dtype)
    return A


def as_mats(A, rows, cols):
    """(batch, ncols, ld) strided layout -> (batch, rows, cols) math-layout copy"""
    return np.transpose(A[:, :cols, :rows], (0, 2, 1)).copy()


def lower(M):
    return np.tril(M)


def pack_lower(A, n):
    """(batch, ncols, ld) strided layout -> (batch, n(n+1)/2): LAPACK packed lower, column by column (rows j..n-1 of
    column j) -- the layout of kblasx?pptrf_batch and of the large golden factors"""
    return np.concatenate([A[:, j, j:n] for j in range(n)], axis=1).copy()


def unpack_lower(P, n, dtype=None):
    """inverse of pack_lower into a zero-filled (batch, n, n) strided array (lda = n)"""
    out = np.zeros((P.shape[0], n, n), dtype=dtype or P.dtype)
    o = 0
    for j in range(n):
        out[:, j, j:n] = P[:, o:o + n - j]
        o += n - j
    return out


def sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def potrf_residual(A0, Lf, n):
    """max over the batch of ||A - L L^T||_F / ||A||_F, fp64 arithmetic"""
    A = as_mats(A0, n, n).astype(np.float64)
    L = np.tril(as_mats(Lf, n, n).astype(np.float64))
    R = A - L @ np.transpose(L, (0, 2, 1))
    num = np.sqrt((R * R).sum(axis=(1, 2)))
    den = np.sqrt((A * A).sum(axis=(1, 2)))
    return float((num / den).max())


# --------------------------------------------------------------------------------------------
# the unmodified reference GPU library (oracle/_ref/libkblas_ref.so, built by oracle/build_ref.sh)

def mangle(name, params):
    """Itanium mangling of a free function over (KBlasHandle*, char, int, ...) style parameters"""
    code = {"H": "P11KBlasHandle", "HH": "PP11KBlasHandle", "c": "c", "i": "i", "l": "l"}
    return f"_Z{len(name)}{name}" + "".join(code[p] for p in params)


class RefLib:
    """ctypes view of the reference library: C names for the kernels, mangled names for management"""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        H, i, l, c, P = C.c_void_p, C.c_int, C.c_long, C.c_char, C.c_void_p
        self._create = getattr(self.lib, mangle("kblasCreate", ["HH"]))
        self._create.argtypes = [C.POINTER(H)]
        self._destroy = getattr(self.lib, mangle("kblasDestroy", ["HH"]))
        self._destroy.argtypes = [C.POINTER(H)]
        self._alloc = getattr(self.lib, mangle("kblasAllocateWorkspace", ["H"]))
        self._alloc.argtypes = [H]
        self.h = H()
        assert self._create(C.byref(self.h)) == 1
        self.H, self.i, self.l, self.c, self.P = H, i, l, c, P

    def wsquery(self, name, sig, *args):
        f = getattr(self.lib, mangle(name, ["H"] + list(sig)))
        f.restype = None
        f.argtypes = [self.H] + [{"c": self.c, "i": self.i}[s] for s in sig]
        f(self.h, *args)

    def allocate(self):
        return self._alloc(self.h)

    def fn(self, name, argtypes):
        f = getattr(self.lib, name)
        f.restype = self.i
        f.argtypes = argtypes
        return f

    def close(self):
        if self.h:
            self._destroy(C.byref(self.h))
            self.h = self.H()


def have_ref():
    return os.path.exists(REF_SO)


# --------------------------------------------------------------------------------------------
# round-2 golden vectors (tests/golden/reference_gpu_r2.npz, made by tests/golden/make_golden_r2.py)

def load_golden(fname):
    path = os.path.join(GOLDEN_DIR, fname)
    if not os.path.exists(path):
        return None
    z = np.load(path)
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def golden_r2_inputs(cases, name):
    """materialise the inputs of one round-2 case: large inputs are stored as (seed, SHA-256) and regenerated here,
    factors shared between cases are stored once and referenced by name (L_from / A_from)"""
    c = cases[name]
    kind, p = name.split("_")[0], name.split("_")[1]
    dt = np.float64 if p == "D" else np.float32
    out = dict(c)
    if kind == "potrfbig":
        n = int(name.split("_n")[1])
        A = rand_spd_batch(1, n, dtype=dt, seed=int(c["seed"]))
        assert sha(A) == str(c["A_in_sha256"]), "numpy RNG stream changed: regenerate tests/golden/reference_gpu_r2.npz"
        out["A_in"] = A
    elif kind == "posvbig":
        n = int(name.split("_n")[1])
        A = rand_spd_batch(1, n, dtype=dt, seed=int(c["seed_A"]))
        assert sha(A) == str(c["A_in_sha256"]), "numpy RNG stream changed: regenerate tests/golden/reference_gpu_r2.npz"
        out["A_in"] = A
    if "L_from" in c:
        out["L_in"] = cases[str(c["L_from"])]["A_out"]
    if "A_from" in c:
        out["A_in"] = cases[str(c["A_from"])]["A_in"]
        out["A_out"] = cases[str(c["A_from"])]["A_out"]
    return out
