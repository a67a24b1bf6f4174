"""kblas{S,D}gemm_batch[_strided] / kblas{S,D}syrk_batch[_strided]: the update steps of the Cholesky
path as public calls, on the library's own DMMA / 3xTF32 fragment kernels (csrc/kernels/gemm_tile.cuh).
The reference routes these through cuBLAS batched GEMM (accumulation order unobservable), so parity with the oracle's
k-sequential fma restatement is tolerance-based: |C - C_oracle| <= 100 k eps max|C|."""
import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu
DT = {"D": np.float64, "S": np.float32}


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available()
    kb = U.kblas()
    h = kb.Handle()
    yield kb, h, torch
    h.destroy()


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("N", "T"), ("T", "N"), ("T", "T")])
@pytest.mark.parametrize("m,n,k", [(16, 16, 16), (32, 32, 32), (8, 8, 8), (33, 17, 5), (5, 70, 33), (64, 64, 64), (100, 3, 47), (1, 1, 1),
                                   (16, 32, 128)])
def test_gemm_batch_vs_oracle(env, p, ta, tb, m, n, k):
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    batch = 23
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cb = (k, n) if tb == "N" else (n, k)
    A = U.rand_batch(batch, ra, ca, ld=ra + 1, dtype=dt, seed=m + 3 * k)
    B = U.rand_batch(batch, rb, cb, ld=rb + 2, dtype=dt, seed=n + 5 * k)
    C0 = U.rand_batch(batch, m, n, ld=m + 3, dtype=dt, seed=7)
    h.gemm_batch_strided_wsquery(batch)
    h.allocate_workspace()
    for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.28, -0.5)):
        Co = C0.copy()
        assert U.oracle_gemm(ta, tb, m, n, k, alpha, A, B, beta, Co) == 1
        tol = 100 * k * U.EPS[dt] * max(1.0, np.abs(Co[:, :, :m]).max())
        Cin = C0.copy()
        if beta == 0.0:
            Cin[:, :, :m] = np.nan        # beta == 0: C must not be read (BLAS / cuBLAS semantics)
        dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, Cin))
        rc = h.gemm_batch_strided(ta, tb, m, n, k, alpha, dA, ra + 1, ca * (ra + 1), dB, rb + 2, cb * (rb + 2), beta, dC, m + 3, n * (m + 3), batch)
        torch.cuda.synchronize()
        assert rc == kb.KBLAS_Success
        got = dC.cpu().numpy()
        assert np.abs(got[:, :, :m] - Co[:, :, :m]).max() <= tol, (alpha, beta, h.last_kernel)
        assert np.array_equal(got[:, :, m:], C0[:, :, m:]), "ldc padding untouched"
        # pointer arrays, shuffled
        dC2 = torch.from_numpy(Cin).cuda()
        perm = torch.randperm(batch, device="cuda")
        pa = (dA.data_ptr() + perm * (ca * (ra + 1) * es)).contiguous()
        pb = (dB.data_ptr() + perm * (cb * (rb + 2) * es)).contiguous()
        pc = (dC2.data_ptr() + perm * (n * (m + 3) * es)).contiguous()
        assert h.gemm_batch(ta, tb, m, n, k, alpha, pa, ra + 1, pb, rb + 2, beta, pc, m + 3, batch, prec=p) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert np.array_equal(dC2.cpu().numpy()[:, :, :m], got[:, :, :m]), "pointer-array result == strided result"
    assert h.gemm_batch_strided(ta, tb, m, n, k, 1.0, dA, ra + 1, ca * (ra + 1), dB, rb + 2, cb * (rb + 2), 0.0, dC, m + 3, n * (m + 3), 0) \
        == kb.KBLAS_Error_WrongInput      # reference Xgemm_batch_core.cuh:181-182


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("trans", ["N", "T"])
@pytest.mark.parametrize("m,n", [(8, 8), (16, 16), (16, 32), (32, 16), (13, 29), (64, 32), (100, 7), (33, 64)])
def test_syrk_batch_vs_oracle(env, p, trans, m, n):
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    batch = 19
    ra, ca = (m, n) if trans == "N" else (n, m)
    A = U.rand_batch(batch, ra, ca, ld=ra + 1, dtype=dt, seed=m + 3 * n)
    C0 = U.rand_batch(batch, m, m, ld=m + 2, dtype=dt, seed=9)
    h.syrk_batch_wsquery(m, batch)
    h.allocate_workspace()
    for alpha, beta in ((-1.0, 1.0), (1.0, 0.0), (0.5, 2.0)):
        Co = C0.copy()
        assert U.oracle_syrk("L", trans, m, n, alpha, A, beta, Co) == 1
        tol = 100 * n * U.EPS[dt] * max(1.0, np.abs(Co[:, :, :m]).max())
        dA, dC = torch.from_numpy(A).cuda(), torch.from_numpy(C0).cuda()
        rc = h.syrk_batch_strided("L", trans, m, n, alpha, dA, ra + 1, ca * (ra + 1), beta, dC, m + 2, m * (m + 2), batch)
        torch.cuda.synchronize()
        assert rc == kb.KBLAS_Success
        got = dC.cpu().numpy()
        M, W, M0 = U.as_mats(got, m, m), U.as_mats(Co, m, m), U.as_mats(C0, m, m)
        assert np.abs(np.tril(M) - np.tril(W)).max() <= tol, (alpha, beta)
        assert np.array_equal(np.triu(M, 1), np.triu(M0, 1)), "strict upper triangle untouched (lower SYRK)"
        assert np.array_equal(got[:, :, m:], C0[:, :, m:])
        dC2 = torch.from_numpy(C0).cuda()
        perm = torch.randperm(batch, device="cuda")
        pa = (dA.data_ptr() + perm * (ca * (ra + 1) * es)).contiguous()
        pc = (dC2.data_ptr() + perm * (m * (m + 2) * es)).contiguous()
        assert h.syrk_batch("L", trans, m, n, alpha, pa, ra + 1, beta, pc, m + 2, batch, prec=p) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert np.array_equal(dC2.cpu().numpy(), got)
    assert h.syrk_batch_strided("U", trans, m, n, 1.0, dA, ra + 1, ca * (ra + 1), 0.0, dC, m + 2, m * (m + 2), batch) == kb.KBLAS_NotImplemented
    if m > 16:
        h2 = kb.Handle()      # workspace protocol skipped: the reference answers -6 for m > 16 (Xsyrk_batch_drivers.cuh:141-149)
        assert h2.syrk_batch_strided("L", trans, m, n, 1.0, dA, ra + 1, ca * (ra + 1), 0.0, dC, m + 2, m * (m + 2), batch) == kb.KBLAS_InsufficientWorkspace
        h2.destroy()


def test_syrk_is_the_cholesky_trailing_update(env):
    """potrf(A) == [potrf(A00); trsm; syrk; potrf(A11)] assembled from the public calls (the reference's own recursion,
    Xpotrf_batch_drivers.cuh:94-133), n = 64 split 32 + 32."""
    kb, h, torch = env
    n, h1, batch = 64, 32, 50
    A0 = U.rand_spd_batch(batch, n, seed=1)
    d1, d2 = torch.from_numpy(A0).cuda(), torch.from_numpy(A0).cuda()
    h.posv_batch_strided_wsquery("R", n, n, batch)
    h.syrk_batch_wsquery(n, batch)
    h.allocate_workspace()
    assert h.potrf_batch_strided("L", n, d1, n, n * n, batch, None) == 1
    s = n * n
    a00, a10, a11 = d2.data_ptr(), d2.data_ptr() + h1 * 8, d2.data_ptr() + (h1 + h1 * n) * 8
    assert h.potrf_batch_strided("L", h1, a00, n, s, batch, None, prec="D") == 1
    assert h.trsm_batch_strided("R", "L", "T", "N", n - h1, h1, 1.0, a00, n, s, a10, n, s, batch, prec="D") == 1
    assert h.syrk_batch_strided("L", "N", n - h1, h1, -1.0, a10, n, s, 1.0, a11, n, s, batch, prec="D") == 1
    assert h.potrf_batch_strided("L", n - h1, a11, n, s, batch, None, prec="D") == 1
    torch.cuda.synchronize()
    L1, L2 = np.tril(U.as_mats(d1.cpu().numpy(), n, n)), np.tril(U.as_mats(d2.cpu().numpy(), n, n))
    assert np.abs(L1 - L2).max() <= 100 * n * U.EPS[np.float64] * np.abs(A0).max()
