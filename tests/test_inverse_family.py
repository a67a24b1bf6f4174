"""kblas{S,D}{trtri,lauum,potri,poti}_batch[_strided]: the consumers of the Cholesky factor.
CPU: the oracle's definitions against numpy (inverse / Gram matrix).  GPU: the CUDA path against the oracle
(100 n eps scaled), plus the defining identities (L X = I, A A^-1 = I); strided and pointer array; n <= 32 in one launch,
trtri for any n through the reference's TRSM recursion; lauum / potri / poti for n > 32 through the blocked in-place LAUUM."""
import numpy as np
import pytest

from tests import _util as U

DT = {"D": np.float64, "S": np.float32}


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 8, 16, 24, 32, 50])
def test_inverse_oracle_matches_numpy(dt, n):
    A = U.rand_spd_batch(4, n, dtype=dt, seed=n)
    L = A.copy()
    U.oracle_potrf(L, n)
    Lm = np.tril(U.as_mats(L, n, n)).astype(np.float64)
    eps = U.EPS[dt]
    X = L.copy()
    assert U.oracle_inv("trtri", n, X) == 1
    assert np.abs(np.tril(U.as_mats(X, n, n)) - np.linalg.inv(Lm)).max() <= 100 * n * eps
    R = L.copy()
    assert U.oracle_inv("lauum", n, R) == 1
    assert np.abs(np.tril(U.as_mats(R, n, n)) - np.tril(np.transpose(Lm, (0, 2, 1)) @ Lm)).max() <= 100 * n * eps * np.abs(Lm).max() ** 2
    P = A.copy()
    assert U.oracle_inv("poti", n, P) == 1
    Ainv = np.linalg.inv(U.as_mats(A, n, n).astype(np.float64))
    assert np.abs(np.tril(U.as_mats(P, n, n)) - np.tril(Ainv)).max() <= 100 * n * eps * np.abs(Ainv).max()
    assert np.array_equal(np.triu(U.as_mats(P, n, n), 1), np.triu(U.as_mats(A, n, n), 1)), "upper triangle untouched"
    assert U.oracle_inv("trtri", n, X, uplo="U") == -2 and U.oracle_inv("trtri", n, X, diag="U") == -2


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available()
    kb = U.kblas()
    h = kb.Handle()
    yield kb, h, torch
    h.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("which", ["trtri", "lauum", "potri", "poti"])
@pytest.mark.parametrize("n", [1, 3, 8, 13, 16, 17, 24, 31, 32])
def test_inverse_family_small_vs_oracle(env, p, which, n):
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    eps = U.EPS[dt]
    batch, lda = 77, n + 2
    A0 = U.rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=n + 7, extra_cols=1)
    if which != "poti":      # the other three start from a lower factor
        U.oracle_potrf(A0, n)
    ref = A0.copy()
    assert U.oracle_inv(which, n, ref) == 1
    scale = max(1.0, np.abs(np.tril(U.as_mats(ref, n, n))).max())
    h.inv_batch_wsquery(which, n, batch, strided=True)
    h.inv_batch_wsquery(which, n, batch, strided=False)
    h.allocate_workspace()
    dA = torch.from_numpy(A0).cuda()
    info = torch.full((batch,), 77, dtype=torch.int32, device="cuda")
    assert h.inv_batch_strided(which, "L", n, dA, lda, (n + 1) * lda, batch, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    got = dA.cpu().numpy()
    M, W, M0 = U.as_mats(got, n, n), U.as_mats(ref, n, n), U.as_mats(A0, n, n)
    assert np.abs(np.tril(M) - np.tril(W)).max() <= 100 * n * eps * scale, h.last_kernel
    assert np.array_equal(np.triu(M, 1), np.triu(M0, 1)), "strict upper triangle untouched"
    assert np.array_equal(got[:, :, n:], A0[:, :, n:]) and np.array_equal(got[:, n:, :], A0[:, n:, :]), "padding untouched"
    assert (info.cpu().numpy() == 77).all()
    # pointer array, shuffled: same bits
    dA2 = torch.from_numpy(A0).cuda()
    perm = torch.randperm(batch, device="cuda")
    ptrs = (dA2.data_ptr() + perm * ((n + 1) * lda * es)).contiguous()
    assert h.inv_batch(which, "L", n, ptrs, lda, batch, None, prec=p) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert np.array_equal(dA2.cpu().numpy(), got)
    # return codes
    assert h.inv_batch_strided(which, "U", n, dA, lda, (n + 1) * lda, batch, None) == kb.KBLAS_NotImplemented
    if which == "trtri":
        assert h.inv_batch_strided(which, "L", n, dA, lda, (n + 1) * lda, batch, None, diag="U") == kb.KBLAS_NotImplemented


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [33, 48, 64, 100, 128, 200, 256])
def test_inverse_family_large_n(env, p, n):
    """n > 32: trtri through the TRSM recursion, lauum as one blocked in-place launch, potri = trtri + lauum, poti = potrf +
    potri -- against the oracle and against the defining identities (L X = I, X = L^T L, X A = I)."""
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    eps = U.EPS[dt]
    batch, lda = 6, n + 3
    A0 = U.rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=n, extra_cols=1)
    L0 = A0.copy()
    U.oracle_potrf(L0, n)
    h.inv_batch_wsquery("poti", n, batch)
    h.inv_batch_wsquery("poti", n, batch, strided=False)
    h.allocate_workspace()
    Lm = np.tril(U.as_mats(L0, n, n)).astype(np.float64)
    Am = U.as_mats(A0, n, n).astype(np.float64)
    Am = np.tril(Am) + np.transpose(np.tril(Am, -1), (0, 2, 1))
    I = np.eye(n)[None]
    for which in ("trtri", "lauum", "potri", "poti"):
        start = A0 if which == "poti" else L0
        dA = torch.from_numpy(start).cuda()
        assert h.inv_batch_strided(which, "L", n, dA, lda, (n + 1) * lda, batch, None) == kb.KBLAS_Success, which
        torch.cuda.synchronize()
        got = dA.cpu().numpy()
        X = np.tril(U.as_mats(got, n, n)).astype(np.float64)
        if which == "trtri":
            assert np.abs(Lm @ X - I).max() <= 100 * n * eps * np.abs(Lm).max() * np.abs(X).max()
        elif which == "lauum":
            W = np.tril(np.transpose(Lm, (0, 2, 1)) @ Lm)
            assert np.abs(X - W).max() <= 100 * n * eps * np.abs(W).max(), h.last_kernel
        else:
            Xs = X + np.transpose(np.tril(X, -1), (0, 2, 1))
            assert np.abs(Xs @ Am - I).max() <= 1000 * n * eps * np.abs(Am).max() * np.abs(Xs).max(), (which, h.last_kernel)
        ref = start.copy()
        assert U.oracle_inv(which, n, ref) == 1
        Wm = np.tril(U.as_mats(ref, n, n))
        assert np.abs(X - Wm).max() <= 100 * n * eps * max(1.0, np.abs(Wm).max()), (which, h.last_kernel)
        M, M0 = U.as_mats(got, n, n), U.as_mats(start, n, n)
        assert np.array_equal(np.triu(M, 1), np.triu(M0, 1)), "strict upper triangle untouched"
        assert np.array_equal(got[:, :, n:], start[:, :, n:]) and np.array_equal(got[:, n:, :], start[:, n:, :]), "padding untouched"
        # pointer array, shuffled: same bits
        dA2 = torch.from_numpy(start).cuda()
        perm = torch.randperm(batch, device="cuda")
        ptrs = (dA2.data_ptr() + perm * ((n + 1) * lda * es)).contiguous()
        assert h.inv_batch(which, "L", n, ptrs, lda, batch, None, prec=p) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert np.array_equal(dA2.cpu().numpy(), got), which
    h2 = kb.Handle()     # workspace protocol: the reference's recursion needs pointer workspace for the pointer-array form
    ptrs = torch.zeros(batch, dtype=torch.int64, device="cuda")
    assert h2.inv_batch("trtri", "L", n, ptrs, n, batch, None, prec=p) == kb.KBLAS_InsufficientWorkspace
    h2.destroy()
