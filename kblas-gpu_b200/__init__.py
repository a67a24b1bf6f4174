"""kblas-gpu_b200 -- B200-native batched very-small-matrix Cholesky (potrf / trsm / potrs / posv).

Thin ctypes view of ``lib/libkblas-gpu.so`` (the C ABI declared in ``include/kblas_ffi.h``).
Names, argument order, argument meaning and return codes mirror the reference C API
(ecrc/kblas-gpu ``include/kblas.h:54-108``, ``include/kblas_batch.h:948-1053, 1486-1566,
2190-2278, 2896-2990``) so that parity tests read like the reference's own test programs:

    h = kb.Handle()                                   # kblasCreate
    h.potrf_batch_strided_wsquery(n, batch)           # kblas_potrf_batch_strided_wsquery
    h.allocate_workspace()                            # kblasAllocateWorkspace
    rc = h.potrf_batch_strided('L', n, A, lda, stride, batch, info)   # kblas{S,D}potrf_batch_strided

There is NO fallback: if the CUDA library is missing or does not load, importing this package
raises.  Matrices are device memory -- pass torch CUDA tensors (their ``data_ptr()`` is used) or
raw integer device addresses.  PyTorch is only plumbing here (allocation, streams).

This directory name contains a hyphen (it is the reference's repository name), so import it with
``importlib.import_module("kblas-gpu_b200")``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libkblas-gpu.so")

# return codes, include/kblas_defs.h (reference include/kblas_defs.h:36-49)
KBLAS_Success = 1
KBLAS_UnknownError = 0
KBLAS_NotSupported = -1
KBLAS_NotImplemented = -2
KBLAS_cuBLAS_Error = -3
KBLAS_WrongConfig = -4
KBLAS_CUDA_Error = -5
KBLAS_InsufficientWorkspace = -6
KBLAS_Error_Allocation = -7
KBLAS_Error_Deallocation = -8
KBLAS_Error_NotInitialized = -9
KBLAS_Error_WrongInput = -10

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C kblas-gpu_b200/csrc` "
        "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
    )
_lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)

_H = C.c_void_p          # kblasHandle_t
_P = C.c_void_p          # device pointer
_i, _l, _c = C.c_int, C.c_long, C.c_char


def _sig(name, res, *args):
    f = getattr(_lib, name)
    f.restype = res
    f.argtypes = list(args)
    return f


_sig("kblasCreate", _i, C.POINTER(_H))
_sig("kblasDestroy", _i, C.POINTER(_H))
_sig("kblasTimerTic", None, _H)
_sig("kblasTimerRecordEnd", None, _H)
_sig("kblasTimerToc", C.c_double, _H)
_sig("kblasCreateStreams", _i, _H, _i)
_sig("kblasGetStream", C.c_void_p, _H)
_sig("kblasSetStream", None, _H, C.c_void_p)
_sig("kblasGetCublasHandle", C.c_void_p, _H)
_sig("kblasEnableMagma", _i, _H)
_sig("kblasGetErrorString", C.c_char_p, _i)
_sig("kblasAllocateWorkspace", _i, _H)
_sig("kblasFreeWorkspace", _i, _H)
_sig("kblas_roundup", _i, _i, _i)
for _n in ("trsm", "posv"):
    _sig(f"kblas_{_n}_batch_wsquery", None, _H, _c, _i, _i, _i)
    _sig(f"kblas_{_n}_batch_strided_wsquery", None, _H, _c, _i, _i, _i)
_sig("kblas_gemm_batch_strided_wsquery", None, _H, _i)
_sig("kblas_syrk_batch_wsquery", None, _H, _i, _i)
for _n in ("trtri", "lauum", "potri", "poti"):
    _sig(f"kblas_{_n}_batch_wsquery", None, _H, _i, _i)
    _sig(f"kblas_{_n}_batch_strided_wsquery", None, _H, _i, _i)
_sig("kblas_potrf_batch_wsquery", None, _H, _i, _i)
_sig("kblas_potrf_batch_strided_wsquery", None, _H, _i, _i)
_sig("kblas_potrs_batch_wsquery", None, _H, _i, _i, _i)
_sig("kblas_potrs_batch_strided_wsquery", None, _H, _i, _i, _i)
_sig("kblas_iset_value_1", _i, _P, _i, _l, C.c_void_p)
_sig("kblas_iset_value_2", _i, _P, _i, _P, _i, _l, C.c_void_p)
_sig("kblas_iset_value_4", _i, _P, _i, _P, _i, _P, _i, _P, _i, _l, C.c_void_p)
_sig("kblas_iset_value_5", _i, _P, _i, _P, _i, _P, _i, _P, _i, _P, _i, _l, C.c_void_p)
_sig("kblasx_workspace_state", _i, _H, _i, C.POINTER(C.c_size_t))
_sig("kblasx_wsquery_bytes", _i, _i, _i, _c, _i, _i, _i, C.POINTER(C.c_size_t))
_sig("kblasx_potrf_smem_plan", _i, _i, C.POINTER(C.c_ubyte))
_sig("kblasx_launch_count", _l, _H)
_sig("kblasx_last_kernel", C.c_char_p, _H)
_sig("kblasx_version", C.c_char_p)
_sig("kblasx_reg_size", _i, _i)
_sig("kblasx_closest_reg_size", _i, _i)
for _p, _t in (("S", C.c_float), ("D", C.c_double)):
    _sig(f"kblasx{_p}potrf_batch_strided_host", _i, _H, _c, _i, _P, _P, _i, _l, _i, _P)
    _sig(f"kblasx{_p}pptrf_batch_strided_host", _i, _H, _c, _i, _P, _P, _l, _i, _P)
    _sig(f"kblasx{_p}pptrf_batch_strided", _i, _H, _c, _i, _P, _l, _i, _P)
    _sig(f"kblasx{_p}pptrf_batch", _i, _H, _c, _i, _P, _i, _P)
    _sig(f"kblasx{_p}tri_pack_batch_strided", _i, _H, _c, _i, _P, _i, _l, _P, _l, _i)
    _sig(f"kblasx{_p}tri_unpack_batch_strided", _i, _H, _c, _i, _P, _l, _P, _i, _l, _i)
    _sig(f"kblas{_p}potrf_batch", _i, _H, _c, _i, _P, _i, _i, _P)
    _sig(f"kblas{_p}trtri_batch", _i, _H, _c, _c, _i, _P, _i, _i, _P)
    _sig(f"kblas{_p}trtri_batch_strided", _i, _H, _c, _c, _i, _P, _i, _l, _i, _P)
    for _n in ("lauum", "potri", "poti"):
        _sig(f"kblas{_p}{_n}_batch", _i, _H, _c, _i, _P, _i, _i, _P)
        _sig(f"kblas{_p}{_n}_batch_strided", _i, _H, _c, _i, _P, _i, _l, _i, _P)
    _sig(f"kblas{_p}gemm_batch", _i, _H, _c, _c, _i, _i, _i, _t, _P, _i, _P, _i, _t, _P, _i, _i)
    _sig(f"kblas{_p}gemm_batch_strided", _i, _H, _c, _c, _i, _i, _i, _t, _P, _i, _l, _P, _i, _l, _t, _P, _i, _l, _i)
    _sig(f"kblas{_p}syrk_batch", _i, _H, _c, _c, _i, _i, _t, _P, _i, _t, _P, _i, _i)
    _sig(f"kblas{_p}syrk_batch_strided", _i, _H, _c, _c, _i, _i, _t, _P, _i, _l, _t, _P, _i, _l, _i)
    _sig(f"kblas{_p}potrf_batch_strided", _i, _H, _c, _i, _P, _i, _l, _i, _P)
    _sig(f"kblas{_p}trsm_batch", _i, _H, _c, _c, _c, _c, _i, _i, _t, _P, _i, _P, _i, _i)
    _sig(f"kblas{_p}trsm_batch_strided", _i, _H, _c, _c, _c, _c, _i, _i, _t, _P, _i, _l, _P, _i, _l, _i)
    _sig(f"kblas{_p}potrs_batch", _i, _H, _c, _c, _i, _i, _P, _i, _P, _i, _i)
    _sig(f"kblas{_p}potrs_batch_strided", _i, _H, _c, _c, _i, _i, _P, _i, _l, _P, _i, _l, _i)
    _sig(f"kblas{_p}posv_batch", _i, _H, _c, _c, _i, _i, _P, _i, _P, _i, _i, _P)
    _sig(f"kblas{_p}posv_batch_strided", _i, _H, _c, _c, _i, _i, _P, _i, _l, _P, _i, _l, _i, _P)
    _sig(f"kblas{_p}set_pointer_1", _i, _P, _P, _i, _l, _l, C.c_void_p)
    _sig(f"kblas{_p}set_pointer_2", _i, _P, _P, _i, _l, _P, _P, _i, _l, _l, C.c_void_p)
    _sig(f"kblas{_p}set_pointer_3", _i, _P, _P, _i, _l, _P, _P, _i, _l, _P, _P, _i, _l, _l, C.c_void_p)


def _ptr(x):
    """device address of a torch tensor / int / None"""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


def _hptr(x):
    """host address of a CPU torch tensor / numpy array / int / None"""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    return int(x)


def _prec(x, prec=None):
    """'S' or 'D' from an explicit flag or a torch dtype"""
    if prec is not None:
        return prec.upper()
    name = str(x.dtype).replace("torch.", "")
    if name == "float64":
        return "D"
    if name == "float32":
        return "S"
    raise TypeError(f"unsupported dtype {x.dtype}: this path implements s and d precision only")


def _ch(c):
    return c.encode() if isinstance(c, str) else c


def version() -> str:
    return _lib.kblasx_version().decode()


def error_string(code: int) -> str:
    """kblasGetErrorString (reference src/kblas_common.cu:171-202)"""
    return _lib.kblasGetErrorString(code).decode()


def roundup(x: int, y: int) -> int:
    return _lib.kblas_roundup(x, y)


def reg_size(n: int) -> bool:
    return bool(_lib.kblasx_reg_size(n))


def closest_reg_size(n: int) -> int:
    return _lib.kblasx_closest_reg_size(n)


def potrf_smem_plan(nblk: int):
    """(nslots, slot[I][K]) of the shared-memory resident Cholesky (csrc/kernels/potrf_smem.cuh); host logic"""
    out = (C.c_ubyte * 64)()
    ns = _lib.kblasx_potrf_smem_plan(nblk, out)
    if ns <= 0:
        raise ValueError(error_string(ns))
    return ns, [[out[i * 8 + k] for k in range(8)] for i in range(8)]


WS_OPS = {"trsm": 0, "potrf": 1, "potrs": 2, "posv": 3}


def wsquery_bytes(op: str, strided: bool, m: int, n: int, batch: int, side: str = "R"):
    """(h_data, h_ptrs, d_data, d_ptrs) bytes the reference's *_wsquery records; pure host code."""
    out = (C.c_size_t * 4)()
    rc = _lib.kblasx_wsquery_bytes(WS_OPS[op], int(strided), _ch(side), m, n, batch, out)
    if rc != KBLAS_Success:
        raise ValueError(error_string(rc))
    return tuple(out)


class Handle:
    """kblasHandle_t (reference struct KBlasHandle, src/kblas_struct.h:311-456).

    Bound to the CUDA device current at construction, stream 0 until ``set_stream``.
    Not thread-safe, one handle per device per host thread -- same contract as the reference.
    """

    def __init__(self):
        self._h = _H()
        rc = _lib.kblasCreate(C.byref(self._h))
        if rc != KBLAS_Success or not self._h:
            raise RuntimeError(f"kblasCreate failed: {error_string(rc)}")

    # -- life cycle -----------------------------------------------------------------
    def destroy(self):
        if self._h:
            _lib.kblasDestroy(C.byref(self._h))
            self._h = _H()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()

    # -- streams / timer ------------------------------------------------------------
    def set_stream(self, stream):
        """kblasSetStream; accepts a torch.cuda.Stream or a raw cudaStream_t value"""
        _lib.kblasSetStream(self._h, getattr(stream, "cuda_stream", stream))

    def get_stream(self) -> int:
        return _lib.kblasGetStream(self._h) or 0

    def create_streams(self, n: int) -> int:
        return _lib.kblasCreateStreams(self._h, n)

    def get_cublas_handle(self) -> int:
        return _lib.kblasGetCublasHandle(self._h) or 0

    def enable_magma(self) -> int:
        return _lib.kblasEnableMagma(self._h)

    def timer_tic(self):
        _lib.kblasTimerTic(self._h)

    def timer_record_end(self):
        _lib.kblasTimerRecordEnd(self._h)

    def timer_toc(self) -> float:
        return _lib.kblasTimerToc(self._h)

    # -- workspace ------------------------------------------------------------------
    def allocate_workspace(self) -> int:
        return _lib.kblasAllocateWorkspace(self._h)

    def free_workspace(self) -> int:
        return _lib.kblasFreeWorkspace(self._h)

    def workspace_state(self, which: str = "allocated"):
        out = (C.c_size_t * 4)()
        _lib.kblasx_workspace_state(self._h, {"requested": 0, "allocated": 1, "consumed": 2}[which], out)
        return tuple(out)

    def trsm_batch_wsquery(self, side, m, n, batch):
        _lib.kblas_trsm_batch_wsquery(self._h, _ch(side), m, n, batch)

    def trsm_batch_strided_wsquery(self, side, m, n, batch):
        _lib.kblas_trsm_batch_strided_wsquery(self._h, _ch(side), m, n, batch)

    def potrf_batch_wsquery(self, n, batch):
        _lib.kblas_potrf_batch_wsquery(self._h, n, batch)

    def potrf_batch_strided_wsquery(self, n, batch):
        _lib.kblas_potrf_batch_strided_wsquery(self._h, n, batch)

    def potrs_batch_wsquery(self, m, n, batch):
        _lib.kblas_potrs_batch_wsquery(self._h, m, n, batch)

    def potrs_batch_strided_wsquery(self, m, n, batch):
        _lib.kblas_potrs_batch_strided_wsquery(self._h, m, n, batch)

    def posv_batch_wsquery(self, side, m, n, batch):
        _lib.kblas_posv_batch_wsquery(self._h, _ch(side), m, n, batch)

    def posv_batch_strided_wsquery(self, side, m, n, batch):
        _lib.kblas_posv_batch_strided_wsquery(self._h, _ch(side), m, n, batch)

    # -- introspection ----------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return _lib.kblasx_launch_count(self._h)

    @property
    def last_kernel(self) -> str:
        return _lib.kblasx_last_kernel(self._h).decode()

    # -- compute: strided ---------------------------------------------------------------
    def potrf_batch_strided(self, uplo, n, A, lda, strideA, batch, info=None, prec=None):
        f = getattr(_lib, f"kblas{_prec(A, prec)}potrf_batch_strided")
        return f(self._h, _ch(uplo), n, _ptr(A), lda, strideA, batch, _ptr(info))

    def trsm_batch_strided(self, side, uplo, trans, diag, m, n, alpha, A, lda, strideA, B, ldb, strideB, batch,
                           prec=None):
        f = getattr(_lib, f"kblas{_prec(B, prec)}trsm_batch_strided")
        return f(self._h, _ch(side), _ch(uplo), _ch(trans), _ch(diag), m, n, alpha, _ptr(A), lda, strideA, _ptr(B),
                 ldb, strideB, batch)

    def trsm_batch_nonuniform(self, side, uplo, trans, diag, m, n, alpha, A_ptrs, lda, B_ptrs, ldb, batch, prec):
        """per-matrix sizes: m, n, lda, ldb are int32 DEVICE tensors, A_ptrs / B_ptrs int64 device tensors of addresses"""
        f = getattr(_lib, f"kblasx{prec}trsm_batch_nonuniform")
        f.argtypes = [C.c_void_p, C.c_char, C.c_char, C.c_char, C.c_char, C.c_void_p, C.c_void_p,
                      C.c_double if prec == "D" else C.c_float] + [C.c_void_p] * 4 + [C.c_int]
        f.restype = C.c_int
        return f(self._h, _ch(side), _ch(uplo), _ch(trans), _ch(diag), _ptr(m), _ptr(n), alpha, _ptr(A_ptrs), _ptr(lda),
                 _ptr(B_ptrs), _ptr(ldb), batch)

    def potrs_batch_strided(self, side, uplo, m, n, A, lda, strideA, B, ldb, strideB, batch, prec=None):
        f = getattr(_lib, f"kblas{_prec(B, prec)}potrs_batch_strided")
        return f(self._h, _ch(side), _ch(uplo), m, n, _ptr(A), lda, strideA, _ptr(B), ldb, strideB, batch)

    def posv_batch_strided(self, side, uplo, m, n, A, lda, strideA, B, ldb, strideB, batch, info=None, prec=None):
        f = getattr(_lib, f"kblas{_prec(B, prec)}posv_batch_strided")
        return f(self._h, _ch(side), _ch(uplo), m, n, _ptr(A), lda, strideA, _ptr(B), ldb, strideB, batch,
                 _ptr(info))

    # -- compute: matrices in HOST memory (pinned for full speed); no reference counterpart ----
    def potrf_batch_strided_host(self, uplo, n, A_in, A_out, lda, strideA, batch, info=None, prec=None):
        """chunked H2D / potrf / D2H pipeline with WHOLE-ARRAY copies by default (csrc/host_pipeline.cu; the
        lower-triangle-only transfer mode KBLAS_B200_HOSTCOPY=tri measured slower and is opt-in).
        A_in / A_out: CPU torch tensors, numpy arrays or raw host addresses holding at least
        (batch-1)*strideA + lda*(n-1) + n elements; A_out may be A_in (in place).  Out of place, A_out receives
        A_in's storage (padding included) with the lower triangles replaced by the factors."""
        f = getattr(_lib, f"kblasx{_prec(A_in, prec)}potrf_batch_strided_host")
        return f(self._h, _ch(uplo), n, _hptr(A_in), _hptr(A_out), lda, strideA, batch, _hptr(info))

    # -- the update steps of the path as public calls (reference kblas_batch.h:264-755) ----------------
    def gemm_batch_strided_wsquery(self, batch):
        _lib.kblas_gemm_batch_strided_wsquery(self._h, batch)

    def syrk_batch_wsquery(self, m, batch):
        _lib.kblas_syrk_batch_wsquery(self._h, m, batch)

    def gemm_batch_strided(self, transA, transB, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, Cm, ldc, strideC, batch,
                           prec=None):
        f = getattr(_lib, f"kblas{_prec(Cm, prec)}gemm_batch_strided")
        return f(self._h, _ch(transA), _ch(transB), m, n, k, alpha, _ptr(A), lda, strideA, _ptr(B), ldb, strideB, beta, _ptr(Cm), ldc,
                 strideC, batch)

    def gemm_batch(self, transA, transB, m, n, k, alpha, A_array, lda, B_array, ldb, beta, C_array, ldc, batch, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}gemm_batch")
        return f(self._h, _ch(transA), _ch(transB), m, n, k, alpha, _ptr(A_array), lda, _ptr(B_array), ldb, beta, _ptr(C_array), ldc, batch)

    def syrk_batch_strided(self, uplo, trans, m, n, alpha, A, lda, strideA, beta, B, ldb, strideB, batch, prec=None):
        f = getattr(_lib, f"kblas{_prec(B, prec)}syrk_batch_strided")
        return f(self._h, _ch(uplo), _ch(trans), m, n, alpha, _ptr(A), lda, strideA, beta, _ptr(B), ldb, strideB, batch)

    def syrk_batch(self, uplo, trans, m, n, alpha, A_array, lda, beta, B_array, ldb, batch, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}syrk_batch")
        return f(self._h, _ch(uplo), _ch(trans), m, n, alpha, _ptr(A_array), lda, beta, _ptr(B_array), ldb, batch)

    # -- the consumers of the factor (reference kblas_batch.h:1611-2729): in place on the lower triangle -----------
    def inv_batch_wsquery(self, which, n, batch, strided=True):
        getattr(_lib, f"kblas_{which}_batch{'_strided' if strided else ''}_wsquery")(self._h, n, batch)

    def inv_batch_strided(self, which, uplo, n, A, lda, strideA, batch, info=None, diag="N", prec=None):
        """which = 'trtri' (A := A^-1), 'lauum' (A := A^T A), 'potri' (A = L := (L L^T)^-1), 'poti' (SPD A := A^-1)"""
        f = getattr(_lib, f"kblas{_prec(A, prec)}{which}_batch_strided")
        if which == "trtri":
            return f(self._h, _ch(uplo), _ch(diag), n, _ptr(A), lda, strideA, batch, _ptr(info))
        return f(self._h, _ch(uplo), n, _ptr(A), lda, strideA, batch, _ptr(info))

    def inv_batch(self, which, uplo, n, A_array, lda, batch, info=None, diag="N", prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}{which}_batch")
        if which == "trtri":
            return f(self._h, _ch(uplo), _ch(diag), n, _ptr(A_array), lda, batch, _ptr(info))
        return f(self._h, _ch(uplo), n, _ptr(A_array), lda, batch, _ptr(info))

    # -- compute: packed lower-triangular layout (LAPACK ?pptrf storage); no reference counterpart ------
    def pptrf_batch_strided(self, uplo, n, AP, strideAP, batch, info=None, prec=None):
        """Cholesky of `batch` matrices stored packed-lower: AP[b*strideAP + j*n - j(j-1)/2 + (i-j)] = A_b(i,j), i >= j"""
        f = getattr(_lib, f"kblasx{_prec(AP, prec)}pptrf_batch_strided")
        return f(self._h, _ch(uplo), n, _ptr(AP), strideAP, batch, _ptr(info))

    def pptrf_batch(self, uplo, n, AP_array, batch, info=None, prec="D"):
        f = getattr(_lib, f"kblasx{prec.upper()}pptrf_batch")
        return f(self._h, _ch(uplo), n, _ptr(AP_array), batch, _ptr(info))

    def tri_pack_batch_strided(self, uplo, n, A, lda, strideA, AP, strideAP, batch, prec=None):
        f = getattr(_lib, f"kblasx{_prec(A, prec)}tri_pack_batch_strided")
        return f(self._h, _ch(uplo), n, _ptr(A), lda, strideA, _ptr(AP), strideAP, batch)

    def tri_unpack_batch_strided(self, uplo, n, AP, strideAP, A, lda, strideA, batch, prec=None):
        f = getattr(_lib, f"kblasx{_prec(A, prec)}tri_unpack_batch_strided")
        return f(self._h, _ch(uplo), n, _ptr(AP), strideAP, _ptr(A), lda, strideA, batch)

    def pptrf_batch_strided_host(self, uplo, n, AP_in, AP_out, strideAP, batch, info=None, prec=None):
        """packed matrices in HOST memory through the chunked H2D / pptrf / D2H pipeline (csrc/host_pipeline.cu)"""
        f = getattr(_lib, f"kblasx{_prec(AP_in, prec)}pptrf_batch_strided_host")
        return f(self._h, _ch(uplo), n, _hptr(AP_in), _hptr(AP_out), strideAP, batch, _hptr(info))

    # -- compute: pointer arrays (device arrays of device pointers) -----------------------
    def potrf_batch(self, uplo, n, A_array, lda, batch, info=None, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}potrf_batch")
        return f(self._h, _ch(uplo), n, _ptr(A_array), lda, batch, _ptr(info))

    def trsm_batch(self, side, uplo, trans, diag, m, n, alpha, A_array, lda, B_array, ldb, batch, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}trsm_batch")
        return f(self._h, _ch(side), _ch(uplo), _ch(trans), _ch(diag), m, n, alpha, _ptr(A_array), lda,
                 _ptr(B_array), ldb, batch)

    def potrs_batch(self, side, uplo, m, n, A_array, lda, B_array, ldb, batch, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}potrs_batch")
        return f(self._h, _ch(side), _ch(uplo), m, n, _ptr(A_array), lda, _ptr(B_array), ldb, batch)

    def posv_batch(self, side, uplo, m, n, A_array, lda, B_array, ldb, batch, info=None, prec="D"):
        f = getattr(_lib, f"kblas{prec.upper()}posv_batch")
        return f(self._h, _ch(side), _ch(uplo), m, n, _ptr(A_array), lda, _ptr(B_array), ldb, batch, _ptr(info))

    # -- helpers the reference's test programs use (src/Xhelper_funcs.ch:48-55) -----------
    def set_pointer_1(self, out_array, base, lda, batch_offset, batch, prec=None):
        f = getattr(_lib, f"kblas{_prec(base, prec)}set_pointer_1")
        return f(_ptr(out_array), _ptr(base), lda, batch_offset, batch, self.get_stream() or None)

    def set_pointer_2(self, out1, base1, ld1, off1, out2, base2, ld2, off2, batch, prec=None):
        f = getattr(_lib, f"kblas{_prec(base1, prec)}set_pointer_2")
        return f(_ptr(out1), _ptr(base1), ld1, off1, _ptr(out2), _ptr(base2), ld2, off2, batch, self.get_stream() or None)

    def set_pointer_3(self, out1, base1, ld1, off1, out2, base2, ld2, off2, out3, base3, ld3, off3, batch, prec=None):
        f = getattr(_lib, f"kblas{_prec(base1, prec)}set_pointer_3")
        return f(_ptr(out1), _ptr(base1), ld1, off1, _ptr(out2), _ptr(base2), ld2, off2, _ptr(out3), _ptr(base3), ld3,
                 off3, batch, self.get_stream() or None)

    def iset_value_1(self, out_array, value, batch):
        return _lib.kblas_iset_value_1(_ptr(out_array), value, batch, self.get_stream() or None)

    def iset_values(self, pairs, batch):
        """iset_value_{1,2,4,5} (reference src/kblas_common.cu:344-386): pairs = [(int32 device array, value), ...]"""
        f = getattr(_lib, f"kblas_iset_value_{len(pairs)}")
        flat = []
        for arr, val in pairs:
            flat += [_ptr(arr), int(val)]
        return f(*flat, batch, self.get_stream() or None)
