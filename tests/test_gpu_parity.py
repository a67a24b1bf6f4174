"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle, the committed
golden vectors of the reference GPU library, and (when oracle/_ref was built) the reference library
run live on the same device buffers.

Tolerances are BASELINE.json's: per-matrix ||A - L L^T||_F / ||A||_F <= 10 n eps; element-wise
deviation from the reference factor <= 100 n eps ||A||; info flags bit-identical (the reference
never writes them); return codes identical.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu
DT = {"D": np.float64, "S": np.float32}
SENT = 77


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    kb = U.kblas()
    h = kb.Handle()
    yield kb, h, torch
    h.destroy()


def _dev(torch, a):
    return torch.from_numpy(a).cuda()


def _check_potrf(A0, A1, n, dt, Lref=None):
    eps = U.EPS[dt]
    assert U.potrf_residual(A0, A1, n) <= 10 * n * eps
    M0, M1 = U.as_mats(A0, n, n), U.as_mats(A1, n, n)
    assert np.array_equal(np.triu(M0, 1), np.triu(M1, 1)), "strict upper triangle must be bit-preserved"
    assert np.array_equal(A0[:, :, n:], A1[:, :, n:]), "rows beyond n (lda padding) must be untouched"
    assert np.array_equal(A0[:, n:, :], A1[:, n:, :]), "columns beyond n (stride padding) must be untouched"
    if Lref is not None:
        normA = np.abs(M0).max()
        assert np.abs(np.tril(M1) - np.tril(U.as_mats(Lref, n, n))).max() <= 100 * n * eps * normA


# =============================================================================================
# potrf
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 13, 16, 17, 23, 24, 25, 31, 32])
def test_potrf_strided_vs_oracle(env, p, n):
    kb, h, torch = env
    dt = DT[p]
    for batch, pad, extra in ((37, 0, 0), (1, 3, 2), (1000, 1, 0)):
        lda = n + pad
        A0 = U.rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=n + batch, extra_cols=extra)
        dA = _dev(torch, A0)
        info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
        h.potrf_batch_strided_wsquery(n, batch)
        h.allocate_workspace()
        rc = h.potrf_batch_strided("L", n, dA, lda, (n + extra) * lda, batch, info)
        torch.cuda.synchronize()
        assert rc == kb.KBLAS_Success
        Lo = A0.copy()
        U.oracle_potrf(Lo, n)
        _check_potrf(A0, dA.cpu().numpy(), n, dt, Lref=Lo)
        assert (info.cpu().numpy() == SENT).all(), "info must not be written (reference parity)"


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [8, 16, 24, 32, 20])
def test_potrf_pointer_array_shuffled(env, p, n):
    kb, h, torch = env
    dt = DT[p]
    batch = 531
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=3 * n)
    dA = _dev(torch, A0)
    perm = np.random.default_rng(n).permutation(batch)
    esz = A0.itemsize
    ptrs = torch.from_numpy((dA.data_ptr() + perm.astype(np.int64) * n * n * esz)).cuda()
    h.potrf_batch_wsquery(n, batch)
    h.allocate_workspace()
    rc = h.potrf_batch("L", n, ptrs, n, batch, None, prec=p)
    torch.cuda.synchronize()
    assert rc == kb.KBLAS_Success
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    _check_potrf(A0, dA.cpu().numpy(), n, dt, Lref=Lo)
    # pointer array built by the library helper (Xset_pointer_1), contiguous order
    dA2 = _dev(torch, A0)
    ptr2 = torch.zeros(batch, dtype=torch.int64, device="cuda")
    assert h.set_pointer_1(ptr2, dA2, n, n * n, batch) == kb.KBLAS_Success
    assert h.potrf_batch("L", n, ptr2, n, batch, None, prec=p) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert np.array_equal(dA2.cpu().numpy(), dA.cpu().numpy())


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [33, 40, 64, 100, 128, 200, 256])
def test_potrf_strided_large_n(env, p, n):
    kb, h, torch = env
    dt = DT[p]
    batch = 9
    A0 = U.rand_spd_batch(batch, n, lda=n + (n % 3), dtype=dt, seed=n)
    dA = _dev(torch, A0)
    h.potrf_batch_strided_wsquery(n, batch)
    h.allocate_workspace()
    rc = h.potrf_batch_strided("L", n, dA, A0.shape[2], n * A0.shape[2], batch, None)
    torch.cuda.synchronize()
    assert rc == kb.KBLAS_Success
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    _check_potrf(A0, dA.cpu().numpy(), n, dt, Lref=Lo)


@pytest.mark.parametrize("variant", [-1, 31, 32, 33, 34, 35, 36])
def test_dpotrf_large_n_kernel_variants(variant, monkeypatch):
    """fp64, 32 < n <= 256: the one-warp-per-matrix DMMA kernel (default) and the opt-in shared-memory resident kernel
    (31 / 32 / 33 = 2 / 4 / 8 warps per matrix): ragged n, padded lda, strided and pointer array, 16-byte aligned and
    element-aligned-only matrices (the cp.async loader's two paths), LAPACK-info mode."""
    import torch

    kb = U.kblas()
    if variant >= 0:
        monkeypatch.setenv("KBLAS_B200_VARIANT", str(variant))
    h = kb.Handle()
    dt = np.float64
    for n in (33, 40, 64, 65, 96, 100, 128, 129, 200, 224, 255, 256):
        batch = 7
        for lda, off in ((n, 0), (n + 3, 0), (n, 1)):
            A0 = U.rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=n + lda + off)
            Lo = A0.copy()
            U.oracle_potrf(Lo, n)
            dA = _slack_copy(torch, A0, off)
            h.potrf_batch_wsquery(n, batch)
            h.potrf_batch_strided_wsquery(n, batch)
            h.allocate_workspace()
            if off == 0:
                rc = h.potrf_batch_strided("L", n, dA, lda, n * lda, batch, None)
            else:
                perm = torch.randperm(batch, device="cuda")
                rc = h.potrf_batch("L", n, _ptrs(torch, dA, off, perm, n * lda, 8), lda, batch, None, prec="D")
            torch.cuda.synchronize()
            assert rc == kb.KBLAS_Success
            want = "potrf_panel_dmma" if variant < 0 else "potrf_smem" + {31: "<W=2,MB=8>", 32: "<W=4,MB=4>", 33: "<W=8>", 34: "<W=2,MB=4>", 35: "<W=4,MB=2>", 36: "<W=1>"}[variant]
            assert want in h.last_kernel, h.last_kernel
            got = dA[off:off + A0.size].cpu().numpy().reshape(A0.shape)
            _check_potrf(A0, got, n, dt, Lref=Lo)
            assert float(dA[:off].abs().sum()) == 0 and float(dA[off + A0.size:].abs().sum()) == 0
    h.destroy()
    # non-SPD: NaN from the bad column on (compat), LAPACK info on request
    monkeypatch.setenv("KBLAS_B200_INFO_MODE", "lapack")
    h = kb.Handle()
    n, batch = 100, 5
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=77)
    A0[1, 40, 40] = -3.0
    A0[3, 99, 99] = -1.0
    dA = _dev(torch, A0)
    info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    h.potrf_batch_strided_wsquery(n, batch)
    h.allocate_workspace()
    assert h.potrf_batch_strided("L", n, dA, n, n * n, batch, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert info.cpu().numpy().tolist() == [0, 41, 0, 100, 0]
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    assert np.array_equal(np.isfinite(np.tril(U.as_mats(dA.cpu().numpy(), n, n))), np.isfinite(np.tril(U.as_mats(Lo, n, n))))
    h.destroy()


def test_potrf_return_codes(env):
    kb, h, torch = env
    n, batch = 16, 4
    dA = _dev(torch, U.rand_spd_batch(batch, n))
    before = dA.clone()
    assert h.potrf_batch_strided("L", n, dA, n, n * n, 0, None) == kb.KBLAS_UnknownError  # empty grid in the reference
    assert h.potrf_batch_strided("U", n, dA, n, n * n, 0, None) == kb.KBLAS_UnknownError
    assert h.potrf_batch_strided("L", 0, dA, n, n * n, batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert torch.equal(dA, before)
    # workspace protocol: pointer-array potrf with n = 64 needs d_ptrs (reference Xpotrf_batch.cu:50-56)
    h2 = kb.Handle()
    ptrs = torch.zeros(batch, dtype=torch.int64, device="cuda")
    assert h2.potrf_batch("L", 64, ptrs, 64, batch, None) == kb.KBLAS_InsufficientWorkspace
    h2.potrf_batch_wsquery(64, batch)
    assert h2.workspace_state("requested") == (0, 0, 0, batch * 24)
    assert h2.allocate_workspace() == kb.KBLAS_Success
    assert h2.workspace_state("allocated") == (0, 0, 0, batch * 24)
    assert h2.workspace_state("requested") == (0, 0, 0, 0)
    h2.destroy()


@pytest.mark.parametrize("p", ["D", "S"])
def test_potrf_non_spd_matches_reference_behaviour(env, p):
    kb, h, torch = env
    dt = DT[p]
    n, batch = 32, 40
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=9)
    A0[3, 5, 5] = -3.0
    A0[17, 20, 20] = 0.0
    dA = _dev(torch, A0)
    info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    assert h.potrf_batch_strided("L", n, dA, n, n * n, batch, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    got = np.tril(U.as_mats(dA.cpu().numpy(), n, n))
    assert np.array_equal(np.isfinite(got), np.isfinite(np.tril(U.as_mats(Lo, n, n))))
    assert (info.cpu().numpy() == SENT).all()


@pytest.mark.parametrize("p", ["D", "S"])
def test_potrf_element_exact_stores_opt_out(p, monkeypatch):
    """KBLAS_B200_ELEMENT_EXACT_STORES=1 (read at kblasCreate): the n % 8 == 0 fast path, which re-writes the strict-upper
    elements sharing a sector with the diagonal with their own bits, is replaced by the element-exact kernel (ADVICE round 1:
    opt-out for callers that update the upper triangle concurrently).  Same factor, bit for bit."""
    import torch

    kb = U.kblas()
    dt = DT[p]
    n, batch = 32, 777
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=21)
    h0 = kb.Handle()
    d0 = _dev(torch, A0)
    assert h0.potrf_batch_strided("L", n, d0, n, n * n, batch, None) == kb.KBLAS_Success
    assert "EX=true" in h0.last_kernel
    monkeypatch.setenv("KBLAS_B200_ELEMENT_EXACT_STORES", "1")
    h1 = kb.Handle()
    d1 = _dev(torch, A0)
    assert h1.potrf_batch_strided("L", n, d1, n, n * n, batch, None) == kb.KBLAS_Success
    assert "EX=false" in h1.last_kernel
    torch.cuda.synchronize()
    assert torch.equal(d0, d1)
    h0.destroy()
    h1.destroy()


def test_potrf_lapack_info_mode_is_opt_in(env):
    kb, _, torch = env
    os.environ["KBLAS_B200_INFO_MODE"] = "lapack"
    try:
        h = kb.Handle()
    finally:
        del os.environ["KBLAS_B200_INFO_MODE"]
    n, batch = 32, 20
    A0 = U.rand_spd_batch(batch, n, seed=10)
    A0[3, 5, 5] = -3.0
    A0[17, 20, 20] = -1.0
    dA = _dev(torch, A0)
    info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    assert h.potrf_batch_strided("L", n, dA, n, n * n, batch, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    want = np.zeros(batch, dtype=np.int32)
    want[3], want[17] = 6, 21
    assert np.array_equal(info.cpu().numpy(), want)
    h.destroy()


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n,lda,extra", [(32, 32, 0), (32, 34, 1), (20, 20, 0), (8, 8, 0), (40, 40, 0), (16, 16, 2)])
@pytest.mark.parametrize("mode", ["tri", "full"])
def test_potrf_host_pipeline(env, p, n, lda, extra, mode, monkeypatch):
    """kblasx?potrf_batch_strided_host (host memory in, host memory out; chunked 3-stream pipeline, lower-triangle
    transfers) == H2D + kblas?potrf_batch_strided + D2H, bit for bit, in place and out of place."""
    kb, h, torch = env
    dt = DT[p]
    batch = 301
    A0 = U.rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=n + 5, extra_cols=extra)
    stride = (n + extra) * lda
    # device path
    dA = _dev(torch, A0)
    h.potrf_batch_strided_wsquery(n, batch)
    h.allocate_workspace()
    assert h.potrf_batch_strided("L", n, dA, lda, stride, batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    want = dA.cpu().numpy()
    # host path: ~6 matrices per chunk -> dozens of chunks through the 3 staging buffers
    monkeypatch.setenv("KBLAS_B200_HOSTCHUNK_MB", str(6.5 * stride * np.dtype(dt).itemsize / (1 << 20)))
    monkeypatch.setenv("KBLAS_B200_HOSTCOPY", mode)
    inplace = A0.copy()
    assert h.potrf_batch_strided_host("L", n, inplace, inplace, lda, stride, batch) == kb.KBLAS_Success
    assert np.array_equal(inplace, want)
    out = np.full_like(A0, 9.5)
    src = A0.copy()
    assert h.potrf_batch_strided_host("L", n, src, out, lda, stride, batch) == kb.KBLAS_Success
    assert np.array_equal(src, A0), "A_in is read-only"
    M, W = U.as_mats(out, n, n), U.as_mats(want, n, n)
    assert np.array_equal(np.tril(M), np.tril(W))
    if mode == "tri" and n > 8 and extra == 0:
        # above the diagonal 8 x 8 blocks nothing is written
        i, j = np.indices((n, n))
        untouched = (j // 8) > (i // 8)          # as_mats gives [b, row, col]
        assert (M[:, untouched] == 9.5).all()
    assert h.potrf_batch_strided_host("U", n, src, out, lda, stride, batch) == kb.KBLAS_NotImplemented
    if mode == "full":
        # a caller with stride > lda*n owns nothing behind the last matrix -- the last chunk must stop at
        # the last element of the last matrix (host buffers of exactly the minimal strided size, guard words behind them)
        need = (batch - 1) * stride + lda * (n - 1) + n
        buf = np.full(need + 64, 123.0, dtype=dt)
        buf[:need] = A0.flatten()[:need]
        assert h.potrf_batch_strided_host("L", n, buf, buf, lda, stride, batch) == kb.KBLAS_Success
        assert np.array_equal(buf[:need], want.flatten()[:need])
        assert (buf[need:] == 123.0).all(), "wrote past the last element of the last matrix"


# =============================================================================================
# trsm / potrs / posv
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("side,trans", [("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")])
@pytest.mark.parametrize("m,n", [(8, 8), (16, 16), (32, 32), (13, 7), (7, 13), (32, 20), (20, 32), (1, 1), (24, 24), (32, 100), (100, 32),
                                 (32, 16), (16, 32), (70, 16), (16, 70), (24, 50), (50, 24)])
def test_trsm_strided_vs_oracle(env, p, side, trans, m, n):
    kb, h, torch = env
    dt = DT[p]
    k = m if side == "L" else n
    if k > 32:
        pytest.skip("covered by test_trsm_large_k")
    batch, alpha = 67, 0.28
    A = U.rand_spd_batch(batch, k, lda=k + 1, dtype=dt, seed=k)
    B0 = U.rand_batch(batch, m, n, ld=m + 2, dtype=dt, seed=m * 100 + n)
    dA, dB = _dev(torch, A), _dev(torch, B0)
    h.trsm_batch_strided_wsquery(side, m, n, batch)
    h.allocate_workspace()
    rc = h.trsm_batch_strided(side, "L", trans, "N", m, n, alpha, dA, k + 1, k * (k + 1), dB, m + 2, n * (m + 2), batch)
    torch.cuda.synchronize()
    assert rc == kb.KBLAS_Success
    Bo = B0.copy()
    U.oracle_trsm(side, "L", trans, "N", m, n, alpha, A, Bo)
    got = dB.cpu().numpy()
    assert np.abs(got - Bo).max() <= 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :m]).max())
    assert np.array_equal(got[:, :, m:], B0[:, :, m:])
    assert np.array_equal(dA.cpu().numpy(), A), "A is read-only"


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("side,trans", [("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")])
@pytest.mark.parametrize("k,other", [(33, 16), (64, 16), (100, 40), (128, 16), (256, 16)])
def test_trsm_large_k(env, p, side, trans, k, other):
    kb, h, torch = env
    dt = DT[p]
    m, n = (k, other) if side == "L" else (other, k)
    batch, alpha = 5, 0.28
    A = U.rand_spd_batch(batch, k, dtype=dt, seed=k)
    B0 = U.rand_batch(batch, m, n, dtype=dt, seed=m * 100 + n)
    dA, dB = _dev(torch, A), _dev(torch, B0)
    h.trsm_batch_strided_wsquery(side, m, n, batch)
    h.allocate_workspace()
    rc = h.trsm_batch_strided(side, "L", trans, "N", m, n, alpha, dA, k, k * k, dB, m, n * m, batch)
    torch.cuda.synchronize()
    assert rc == kb.KBLAS_Success
    Bo = B0.copy()
    U.oracle_trsm(side, "L", trans, "N", m, n, alpha, A, Bo)
    assert np.abs(dB.cpu().numpy() - Bo).max() <= 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo).max())


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("op", ["N", "T", "potrs"])
@pytest.mark.parametrize("k,vec,pad", [(32, 32, 0), (32, 100, 1), (24, 24, 0), (24, 40, 2), (16, 32, 1), (16, 17, 0), (8, 33, 1),
                                         (20, 32, 1), (28, 9, 0), (12, 64, 0), (4, 5, 1)])
def test_left_side_16_byte_kernel(p, op, k, vec, pad, monkeypatch):
    """kernels/trsm_left_vec.cuh (strided side L, every global access 16 bytes wide): forced on with variant 40 wherever
    its alignment conditions hold, for every factor order / ragged slab / padded leading dimension it claims, against the
    oracle; plus the fallback when the operands are NOT 16-byte aligned (same numbers from the element-wise kernels)."""
    import torch
    monkeypatch.setenv("KBLAS_B200_VARIANT", "40")
    kb = U.kblas()
    h = kb.Handle()
    try:
        dt = DT[p]
        vw = 16 // np.dtype(dt).itemsize
        batch, alpha = 37, 0.28
        lda, ldb = k + pad * vw, k + 2 * pad * vw
        A = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=k + 7)
        assert U.oracle_potrf(A, k) == 1                # a real Cholesky factor in the lower triangle
        B0 = U.rand_batch(batch, k, vec, ld=ldb, dtype=dt, seed=k * 100 + vec)
        Bo = B0.copy()
        if op == "potrs":                               # (L L^T) X = B: the restated trsm L,L,N then L,L,T
            U.oracle_trsm("L", "L", "N", "N", k, vec, 1.0, A, Bo)
            U.oracle_trsm("L", "L", "T", "N", k, vec, 1.0, A, Bo)
        else:
            U.oracle_trsm("L", "L", op, "N", k, vec, alpha, A, Bo)
        tol = 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :k]).max())

        def run(dA, dB, a_off, b_off):
            Av, Bv = dA[a_off:], dB[b_off:]
            if op == "potrs":
                return h.potrs_batch_strided("L", "L", k, vec, Av, lda, k * lda, Bv, ldb, vec * ldb, batch)
            return h.trsm_batch_strided("L", "L", op, "N", k, vec, alpha, Av, lda, k * lda, Bv, ldb, vec * ldb, batch)

        # aligned operands: the 16-byte kernel
        dA, dB = _dev(torch, A).flatten(), _dev(torch, B0).flatten()
        assert run(dA, dB, 0, 0) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert h.last_kernel.startswith("tri_left_vec"), h.last_kernel
        got = dB.cpu().numpy().reshape(B0.shape)
        assert np.abs(got[:, :, :k] - Bo[:, :, :k]).max() <= tol
        assert np.array_equal(got[:, :, k:], B0[:, :, k:]), "ldb padding untouched"
        assert np.array_equal(dA.cpu().numpy().reshape(A.shape), A), "the factor is read-only"
        # B one element past a 16-byte boundary: not eligible, same result from the element-wise kernels
        dB1 = _slack_copy(torch, B0, 1)
        assert run(dA, dB1, 0, 1) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert not h.last_kernel.startswith("tri_left_vec"), h.last_kernel
        got1 = dB1.cpu().numpy()[1:1 + B0.size].reshape(B0.shape)
        assert np.abs(got1[:, :, :k] - Bo[:, :, :k]).max() <= tol
    finally:
        h.destroy()


@pytest.mark.parametrize("k,vec,pad", [(32, 32, 0), (32, 100, 1), (28, 45, 1), (26, 32, 0)])
def test_right_side_one_vector_kernel(env, k, vec, pad):
    """kernels/trsm_left_vec.cuh, tri_right_vec: dtrsm R,L,N (X L = alpha B) with a 16-byte aligned strided factor of order
    25..32 and at least 32 rows -- picked by default; against the oracle, and against the two-vector kernel it replaces
    (B one element off a 16-byte boundary changes nothing for side R, an odd lda makes the factor ineligible)."""
    kb, h, torch = env
    dt = np.float64
    batch, alpha = 41, 0.28
    lda, ldb = k + 2 * pad, vec + 3 * pad
    A = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=k + 11)
    assert U.oracle_potrf(A, k) == 1
    B0 = U.rand_batch(batch, vec, k, ld=ldb, dtype=dt, seed=k * 100 + vec)
    Bo = B0.copy()
    U.oracle_trsm("R", "L", "N", "N", vec, k, alpha, A, Bo)
    tol = 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :vec]).max())
    dA, dB = _dev(torch, A), _dev(torch, B0)
    assert h.trsm_batch_strided("R", "L", "N", "N", vec, k, alpha, dA, lda, k * lda, dB, ldb, k * ldb, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert h.last_kernel.startswith("tri_right_vec"), h.last_kernel
    got = dB.cpu().numpy()
    assert np.abs(got[:, :, :vec] - Bo[:, :, :vec]).max() <= tol
    assert np.array_equal(got[:, :, vec:], B0[:, :, vec:]), "ldb padding untouched"
    assert np.array_equal(dA.cpu().numpy(), A), "the factor is read-only"
    # factor one element off a 16-byte boundary: the two-vector / element-wise kernels, same numbers
    dA1 = _slack_copy(torch, A, 1)
    dB.copy_(_dev(torch, B0))
    assert h.trsm_batch_strided("R", "L", "N", "N", vec, k, alpha, dA1[1:], lda, k * lda, dB, ldb, k * ldb, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert not h.last_kernel.startswith("tri_right_vec"), h.last_kernel
    assert np.abs(dB.cpu().numpy()[:, :, :vec] - Bo[:, :, :vec]).max() <= tol


def _tri_batch(batch, k, lda, dt, upper, seed):
    """well-conditioned triangular matrices (batch, k, lda) column-major: diagonal in [1, 2), off-diagonal 0.2 * U(-1, 1) in the
    stored triangle, NaN in the other triangle and in the padding (they must never be referenced)"""
    rng = np.random.default_rng(seed)
    M = 0.2 * (2 * rng.random((batch, k, k)) - 1)
    M = np.triu(M, 1) if upper else np.tril(M, -1)               # [b, row, col]
    M = M + np.eye(k)[None] * (1 + rng.random((batch, k, 1)))
    A = np.full((batch, k, lda), np.nan, dtype=dt)
    keep = np.triu(np.ones((k, k), bool)) if upper else np.tril(np.ones((k, k), bool))
    cm = np.transpose(M, (0, 2, 1)).astype(dt)                    # [b, col, row]
    A[:, :, :k] = np.where(np.transpose(keep)[None], cm, np.nan)
    return A, M.astype(dt).astype(np.float64)


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("k,other", [(5, 7), (16, 16), (32, 32), (24, 40), (40, 16), (100, 9), (64, 33)])
def test_upper_and_unit_extensions(env, p, k, other):
    """uplo = Upper and diag = Unit are KBLAS_NotImplemented in the reference (Xtrsm_batch_drivers.cuh:64-67,
    Xpotrf_batch_drivers.cuh:38-41, Xpotrs_batch_drivers.cuh:40-43) and implemented here: all 16 trsm
    variants against a float64 numpy solve, with NaN in the triangle (and, for Unit, on the diagonal) that must not be
    referenced; potrf Upper bit-identical to the transposed Lower factor; potrs / posv Upper against the Lower solution."""
    kb, h, torch = env
    dt = DT[p]
    eps = U.EPS[dt]
    batch, alpha = 19, 0.28
    lda = k + 1
    h.posv_batch_strided_wsquery("R", max(k, other), max(k, other), batch)
    h.posv_batch_strided_wsquery("L", max(k, other), max(k, other), batch)
    h.allocate_workspace()
    for uplo in ("U", "L"):
        A, M = _tri_batch(batch, k, lda, dt, uplo == "U", seed=k + (uplo == "U"))
        for diag in ("N", "U"):
            if uplo == "L" and diag == "N":
                continue  # the reference's own case: covered everywhere else
            Ad, Md = A.copy(), M.copy()
            if diag == "U":
                Ad[:, np.arange(k), np.arange(k)] = np.nan
                Md[:, np.arange(k), np.arange(k)] = 1.0
            dA = _dev(torch, Ad)
            for side in ("L", "R"):
                m, n = (k, other) if side == "L" else (other, k)
                B0 = U.rand_batch(batch, m, n, ld=m + 2, dtype=dt, seed=m * 100 + n)
                Bm = U.as_mats(B0, m, n).astype(np.float64)
                for trans in ("N", "T"):
                    Op = Md if trans == "N" else np.transpose(Md, (0, 2, 1))
                    want = np.linalg.solve(Op, alpha * Bm) if side == "L" else np.transpose(
                        np.linalg.solve(np.transpose(Op, (0, 2, 1)), np.transpose(alpha * Bm, (0, 2, 1))), (0, 2, 1))
                    dB = _dev(torch, B0)
                    rc = h.trsm_batch_strided(side, uplo, trans, diag, m, n, alpha, dA, lda, k * lda, dB, m + 2, n * (m + 2), batch)
                    torch.cuda.synchronize()
                    assert rc == kb.KBLAS_Success
                    got = dB.cpu().numpy()
                    X = U.as_mats(got, m, n).astype(np.float64)
                    assert np.isfinite(X).all(), (uplo, diag, side, trans, h.last_kernel)
                    assert np.abs(X - want).max() <= 100 * k * eps * max(1.0, np.abs(want).max()), (uplo, diag, side, trans, h.last_kernel)
                    assert np.array_equal(got[:, :, m:], B0[:, :, m:]), "ldb padding untouched"
            assert np.array_equal(dA.cpu().numpy(), Ad, equal_nan=True), "A is read-only"
    # ---- potrf / potrs / posv with uplo = Upper --------------------------------------------------------------------
    A0 = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=k + 3)
    sym = U.as_mats(A0, k, k)
    sym = np.tril(sym) + np.transpose(np.tril(sym, -1), (0, 2, 1))
    Afull = A0.copy()
    Afull[:, :, :k] = np.transpose(sym, (0, 2, 1))       # both triangles stored
    dL, dU = _dev(torch, Afull), _dev(torch, Afull)
    assert h.potrf_batch_strided("L", k, dL, lda, k * lda, batch, None) == kb.KBLAS_Success
    assert h.potrf_batch_strided("U", k, dU, lda, k * lda, batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    Lm, Um = U.as_mats(dL.cpu().numpy(), k, k), U.as_mats(dU.cpu().numpy(), k, k)
    assert np.array_equal(np.triu(Um), np.transpose(np.tril(Lm), (0, 2, 1))), "U = L^T, bit for bit"
    assert np.array_equal(np.tril(Um, -1), np.tril(sym, -1)), "strictly lower triangle untouched"
    assert np.array_equal(dU.cpu().numpy()[:, :, k:], Afull[:, :, k:]), "padding untouched"
    for side in ("R", "L"):
        m, n = (other, k) if side == "R" else (k, other)
        B0 = U.rand_batch(batch, m, n, ld=m + 1, dtype=dt, seed=7 * m + n)
        dB1, dB2, dB3 = _dev(torch, B0), _dev(torch, B0), _dev(torch, B0)
        assert h.potrs_batch_strided(side, "L", m, n, dL, lda, k * lda, dB1, m + 1, n * (m + 1), batch) == kb.KBLAS_Success
        assert h.potrs_batch_strided(side, "U", m, n, dU, lda, k * lda, dB2, m + 1, n * (m + 1), batch) == kb.KBLAS_Success
        dA3 = _dev(torch, Afull)
        assert h.posv_batch_strided(side, "U", m, n, dA3, lda, k * lda, dB3, m + 1, n * (m + 1), batch, None) == kb.KBLAS_Success
        torch.cuda.synchronize()
        X1, X2, X3 = (U.as_mats(d.cpu().numpy(), m, n).astype(np.float64) for d in (dB1, dB2, dB3))
        tol = 100 * k * eps * max(1.0, np.abs(X1).max())
        assert np.abs(X2 - X1).max() <= tol and np.abs(X3 - X1).max() <= tol, (side, h.last_kernel)
        assert torch.equal(dA3, dU), "posv Upper leaves the same factor as potrf Upper"
    # pointer-array form of the Upper factorisation: same bits
    es = np.dtype(dt).itemsize
    dP = _dev(torch, Afull)
    perm = torch.randperm(batch, device="cuda")
    ptrs = (dP.data_ptr() + perm * (k * lda * es)).contiguous()
    h.potrf_batch_wsquery(k, batch)
    h.allocate_workspace()
    assert h.potrf_batch("U", k, ptrs, lda, batch, None, prec=p) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert torch.equal(dP, dU)


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("side,uplo,trans,diag", [("L", "L", "N", "N"), ("L", "L", "T", "N"), ("R", "L", "N", "N"), ("R", "L", "T", "N"),
                                                   ("L", "U", "N", "U"), ("R", "U", "T", "N")])
def test_trsm_nonuniform_batch(env, p, side, uplo, trans, diag):
    """kblas_trsm_batch with per-matrix m[b], n[b], lda[b], ldb[b] (device arrays): MAGMA-only in the reference
    (Xtrsm_batch_drivers.cuh:277-367), native here.  Random sizes 0..70 per matrix, every matrix
    checked against a float64 numpy solve; storage around each matrix must stay untouched."""
    kb, h, torch = env
    dt = DT[p]
    eps = U.EPS[dt]
    rng = np.random.default_rng(11)
    batch, alpha = 37, 0.28
    ms = rng.integers(0, 71, batch).astype(np.int32)
    ns = rng.integers(0, 71, batch).astype(np.int32)
    ms[0], ns[0] = 64, 33
    ms[1], ns[1] = 1, 1
    ks = ms if side == "L" else ns
    ldas = (np.maximum(ks, 1) + rng.integers(0, 3, batch)).astype(np.int32)
    ldbs = (np.maximum(ms, 1) + rng.integers(0, 3, batch)).astype(np.int32)
    a_off = np.concatenate([[0], np.cumsum(ldas.astype(np.int64) * np.maximum(ks, 1) + 5)])
    b_off = np.concatenate([[0], np.cumsum(ldbs.astype(np.int64) * np.maximum(ns, 1) + 5)])
    Abuf = np.full(a_off[-1], np.nan, dtype=dt)
    Bbuf = np.full(b_off[-1], -3.5, dtype=dt)
    mats, rhs = [], []
    for b in range(batch):
        k, m, n = int(ks[b]), int(ms[b]), int(ns[b])
        if k > 0:
            Ab, M = _tri_batch(1, k, int(ldas[b]), dt, uplo == "U", seed=100 + b)
            if diag == "U":
                Ab[0, np.arange(k), np.arange(k)] = np.nan
                M[0, np.arange(k), np.arange(k)] = 1.0
            Abuf[a_off[b]:a_off[b] + k * ldas[b]] = Ab.flatten()
            mats.append(M[0])
        else:
            mats.append(None)
        Bb = rng.random((max(n, 0), int(ldbs[b]))).astype(dt)
        Bb[:, m:] = -3.5
        Bbuf[b_off[b]:b_off[b] + n * ldbs[b]] = Bb.flatten()
        rhs.append(Bb[:, :m].T.astype(np.float64).copy())
    dA, dB = torch.from_numpy(Abuf).cuda(), torch.from_numpy(Bbuf).cuda()
    es = np.dtype(dt).itemsize
    pa = torch.from_numpy(dA.data_ptr() + a_off[:-1] * es).cuda()
    pb = torch.from_numpy(dB.data_ptr() + b_off[:-1] * es).cuda()
    dm, dn = torch.from_numpy(ms).cuda(), torch.from_numpy(ns).cuda()
    dlda, dldb = torch.from_numpy(ldas).cuda(), torch.from_numpy(ldbs).cuda()
    rc = h.trsm_batch_nonuniform(side, uplo, trans, diag, dm, dn, alpha, pa, dlda, pb, dldb, batch, p)
    torch.cuda.synchronize()
    assert rc == kb.KBLAS_Success
    assert h.last_kernel == "tri_nonuniform"
    got = dB.cpu().numpy()
    assert np.array_equal(dA.cpu().numpy(), Abuf, equal_nan=True), "A is read-only"
    for b in range(batch):
        k, m, n = int(ks[b]), int(ms[b]), int(ns[b])
        blk = got[b_off[b]:b_off[b] + n * ldbs[b]].reshape(n, ldbs[b]) if n > 0 else np.zeros((0, ldbs[b]), dt)
        assert (got[b_off[b] + n * ldbs[b]:b_off[b + 1]] == -3.5).all(), "gap after the matrix untouched"
        if n > 0:
            assert (blk[:, m:] == -3.5).all(), "ldb padding untouched"
        if m <= 0 or n <= 0:
            continue
        X = blk[:, :m].T.astype(np.float64)
        Op = mats[b] if trans == "N" else mats[b].T
        want = np.linalg.solve(Op, alpha * rhs[b]) if side == "L" else np.linalg.solve(Op.T, alpha * rhs[b].T).T
        assert np.abs(X - want).max() <= 100 * max(k, 1) * eps * max(1.0, np.abs(want).max()), (b, m, n)


@pytest.mark.parametrize("p,k", [("D", 24), ("S", 24), ("S", 32)])
@pytest.mark.parametrize("vec,pad", [(32, 0), (40, 1), (17, 2)])
def test_potrs_left_two_vector_16_byte_kernel(env, p, k, vec, pad):
    """fused side-L potrs on the two-vector kernel with 16-byte staging (tri_dual16: k = 24, and k = 32 in fp32, aligned strided
    operands, more than 16 right-hand sides) -- the default for those shapes; against the oracle's trsm L,L,N + L,L,T."""
    kb, h, torch = env
    dt = DT[p]
    vw = 16 // np.dtype(dt).itemsize
    batch = 43
    lda, ldb = k + pad * vw, k + 2 * pad * vw
    A = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=k + 21)
    assert U.oracle_potrf(A, k) == 1
    B0 = U.rand_batch(batch, k, vec, ld=ldb, dtype=dt, seed=k * 10 + vec)
    Bo = B0.copy()
    U.oracle_trsm("L", "L", "N", "N", k, vec, 1.0, A, Bo)
    U.oracle_trsm("L", "L", "T", "N", k, vec, 1.0, A, Bo)
    dA, dB = _dev(torch, A), _dev(torch, B0)
    h.potrs_batch_strided_wsquery(k, vec, batch)
    h.allocate_workspace()
    assert h.potrs_batch_strided("L", "L", k, vec, dA, lda, k * lda, dB, ldb, vec * ldb, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert h.last_kernel.startswith("tri_dual16"), h.last_kernel
    got = dB.cpu().numpy()
    assert np.abs(got[:, :, :k] - Bo[:, :, :k]).max() <= 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :k]).max())
    assert np.array_equal(got[:, :, k:], B0[:, :, k:]), "ldb padding untouched"
    assert np.array_equal(dA.cpu().numpy(), A), "the factor is read-only"


def _slack_copy(torch, a, off):
    """device copy of numpy array `a` with `off` elements of slack in front (and 4 behind): every matrix then starts
    `off` elements past a 16-byte boundary -- with off = 1 the pointers are element-aligned but NOT 16-byte aligned"""
    d = torch.zeros(a.size + off + 4, dtype=getattr(torch, a.dtype.name), device="cuda")
    d[off:off + a.size] = torch.from_numpy(a).cuda().flatten()
    return d


def _ptrs(torch, d, off, perm, elems_per_matrix, itemsize):
    return (d.data_ptr() + (off + perm * elems_per_matrix) * itemsize).contiguous()


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("side,trans", [("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")])
@pytest.mark.parametrize("k", [5, 8, 13, 16, 24, 32, 40, 48])
def test_trsm_pointer_array_default_kernels(env, p, side, trans, k):
    """kblas?trsm_batch (pointer array) on the DEFAULT kernels -- no variant override -- with shuffled pointer arrays,
    16-byte aligned and element-aligned-only (off by one element) matrices, lda == k and padded, few / many vectors.
    Reference: Xtrsm_batch.cu:42-112."""
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    batch, alpha = 131, -1.7
    seen = set()
    for vec in (3, 16, 33):
        for off, pad in ((0, 0), (1, 0), (1, 1)):
            m, n = (k, vec) if side == "L" else (vec, k)
            lda, ldb = k + pad, m + pad
            A = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=k + 3)
            B0 = U.rand_batch(batch, m, n, ld=ldb, dtype=dt, seed=m * 10 + n)
            Bo = B0.copy()
            U.oracle_trsm(side, "L", trans, "N", m, n, alpha, A, Bo)
            dA, dB = _slack_copy(torch, A, off), _slack_copy(torch, B0, off)
            perm = torch.randperm(batch, device="cuda")
            pa, pb = _ptrs(torch, dA, off, perm, k * lda, es), _ptrs(torch, dB, off, perm, n * ldb, es)
            h.trsm_batch_wsquery(side, m, n, batch)
            h.allocate_workspace()
            rc = h.trsm_batch(side, "L", trans, "N", m, n, alpha, pa, lda, pb, ldb, batch, prec=p)
            torch.cuda.synchronize()
            assert rc == kb.KBLAS_Success
            seen.add(h.last_kernel.split("<")[0])
            got = dB[off:off + B0.size].cpu().numpy().reshape(B0.shape)
            tol = 100 * k * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :m]).max())
            assert np.abs(got[:, :, :m] - Bo[:, :, :m]).max() <= tol, (vec, off, pad, h.last_kernel)
            assert np.array_equal(got[:, :, m:], B0[:, :, m:]), "ldb padding untouched"
            assert np.array_equal(dA[off:off + A.size].cpu().numpy().reshape(A.shape), A), "A is read-only"
            assert float(dB[:off].abs().sum()) == 0 and float(dB[off + B0.size:].abs().sum()) == 0, "slack untouched"
    want = {"tri_blocked"} if k > 32 else ({"tri_dual"} if k in (16, 24, 32) and not (p == "S" and side == "L" and k == 32) else set())
    assert want <= seen, (seen, want)
    assert not any("bcast" in s for s in seen)


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [5, 8, 16, 24, 32, 48])
def test_potrs_posv_pointer_array_default_kernels(env, p, n):
    """kblas?potrs_batch and kblas?posv_batch, pointer-array form, n <= 32 and the first blocked size, shuffled and
    off-by-one-element pointers.  Reference: Xpotrs_batch.cu:42-101, Xposv_batch.cu:42-107."""
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    batch = 97
    for m in (3, 16, 40):
        for off in (0, 1):
            A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n + 1)
            B0 = U.rand_batch(batch, m, n, dtype=dt, seed=n + 2 + m)
            Ao, Bo = A0.copy(), B0.copy()
            U.oracle_posv("R", "L", m, n, Ao, Bo)
            tol = 100 * n * U.EPS[dt] * max(1.0, np.abs(Bo).max())
            perm = torch.randperm(batch, device="cuda")
            # potrs from the oracle's factor
            dL, dB = _slack_copy(torch, Ao, off), _slack_copy(torch, B0, off)
            pa, pb = _ptrs(torch, dL, off, perm, n * n, es), _ptrs(torch, dB, off, perm, m * n, es)
            h.posv_batch_wsquery("R", m, n, batch)
            h.potrs_batch_wsquery(m, n, batch)
            h.allocate_workspace()
            assert h.potrs_batch("R", "L", m, n, pa, n, pb, m, batch, prec=p) == kb.KBLAS_Success
            torch.cuda.synchronize()
            got = dB[off:off + B0.size].cpu().numpy().reshape(B0.shape)
            assert np.abs(got - Bo).max() <= tol, ("potrs", m, off, h.last_kernel)
            assert np.array_equal(dL[off:off + Ao.size].cpu().numpy().reshape(Ao.shape), Ao), "factor is read-only"
            # posv from A
            dA, dB2 = _slack_copy(torch, A0, off), _slack_copy(torch, B0, off)
            pa, pb = _ptrs(torch, dA, off, perm, n * n, es), _ptrs(torch, dB2, off, perm, m * n, es)
            info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
            assert h.posv_batch("R", "L", m, n, pa, n, pb, m, batch, info, prec=p) == kb.KBLAS_Success
            torch.cuda.synchronize()
            _check_potrf(A0, dA[off:off + A0.size].cpu().numpy().reshape(A0.shape), n, dt, Lref=Ao)
            got = dB2[off:off + B0.size].cpu().numpy().reshape(B0.shape)
            assert np.abs(got - Bo).max() <= tol, ("posv", m, off, h.last_kernel)
            assert (info.cpu().numpy() == SENT).all()


def test_pointer_and_value_helpers(env):
    """Xset_pointer_{1,2,3} / iset_value_{1,2,4,5} (reference Xhelper_funcs.cu:74-105, kblas_common.cu:344-386)"""
    kb, h, torch = env
    batch = 1000
    bases = [torch.zeros(16, dtype=torch.float64, device="cuda") for _ in range(3)]
    outs = [torch.zeros(batch, dtype=torch.int64, device="cuda") for _ in range(3)]
    offs = [64, 7, 1024]
    i = torch.arange(batch, dtype=torch.int64, device="cuda")
    assert h.set_pointer_1(outs[0], bases[0], 8, offs[0], batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert torch.equal(outs[0], bases[0].data_ptr() + i * offs[0] * 8)
    for o in outs:
        o.zero_()
    assert h.set_pointer_2(outs[0], bases[0], 8, offs[0], outs[1], bases[1], 8, offs[1], batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    for q in range(2):
        assert torch.equal(outs[q], bases[q].data_ptr() + i * offs[q] * 8)
    for o in outs:
        o.zero_()
    assert h.set_pointer_3(outs[0], bases[0], 8, offs[0], outs[1], bases[1], 8, offs[1], outs[2], bases[2], 8, offs[2], batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    for q in range(3):
        assert torch.equal(outs[q], bases[q].data_ptr() + i * offs[q] * 8)
    # fp32 flavour: element size 4
    f32 = torch.zeros(16, dtype=torch.float32, device="cuda")
    assert h.set_pointer_1(outs[0], f32, 8, 5, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert torch.equal(outs[0], f32.data_ptr() + i * 5 * 4)
    for count in (1, 2, 4, 5):
        arrs = [torch.zeros(batch, dtype=torch.int32, device="cuda") for _ in range(count)]
        assert h.iset_values([(a, 11 + q) for q, a in enumerate(arrs)], batch) == kb.KBLAS_Success
        torch.cuda.synchronize()
        for q, a in enumerate(arrs):
            assert bool((a == 11 + q).all())


@pytest.mark.parametrize("p", ["D", "S"])
def test_trsm_alpha_zero_is_blas_semantics(env, p):
    """alpha == 0: B := 0 for every variant and size (BLAS semantics).  Documented deviation (SURVEY Appendix A,
    DESIGN.md §1): the reference's recursion multiplies by -1/alpha for side R / trans T and k > 16
    (Xtrsm_batch_drivers.cuh:154-163) and returns Inf/NaN there; for k <= 16 it also returns zeros."""
    kb, h, torch = env
    dt = DT[p]
    for side, trans in (("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")):
        for k in (8, 16, 32, 64):
            m, n, batch = k, k, 33
            A = U.rand_spd_batch(batch, k, dtype=dt, seed=k)
            dA, dB = _dev(torch, A), _dev(torch, U.rand_batch(batch, m, n, dtype=dt, seed=3))
            h.trsm_batch_strided_wsquery(side, m, n, batch)
            h.allocate_workspace()
            assert h.trsm_batch_strided(side, "L", trans, "N", m, n, 0.0, dA, k, k * k, dB, m, m * n, batch) == kb.KBLAS_Success
            torch.cuda.synchronize()
            assert bool((dB == 0).all()), (side, trans, k)


def test_trsm_potrs_random_shapes(env):
    """seeded sweep over ragged shapes / leading dimensions / both precisions: every dispatch branch of the k <= 32
    solves (register, broadcast, packed, dual, one-vector kernels) and the blocked kernel for a few k > 32"""
    kb, h, torch = env
    rng = np.random.default_rng(2026)
    seen = set()
    for it in range(120):
        p = "DS"[it % 2]
        dt = DT[p]
        side, trans = [("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")][rng.integers(4)]
        k = int(rng.choice([1, 2, 5, 8, 9, 12, 16, 17, 23, 24, 29, 32, 33, 48, 64]))
        vec = int(rng.choice([1, 3, 8, 15, 16, 17, 32, 33, 47]))
        m, n = (k, vec) if side == "L" else (vec, k)
        lda, ldb = k + int(rng.integers(0, 3)), m + int(rng.integers(0, 3))
        batch, alpha = int(rng.integers(1, 40)), float(rng.choice([1.0, 0.28, -2.5]))
        A = U.rand_spd_batch(batch, k, lda=lda, dtype=dt, seed=it)
        B0 = U.rand_batch(batch, m, n, ld=ldb, dtype=dt, seed=1000 + it)
        # factor on the CPU (LAPACK-grade lower factor), solve on the GPU
        Lf = A.copy()
        U.oracle_potrf(Lf, k)
        Bo = B0.copy()
        U.oracle_trsm(side, "L", trans, "N", m, n, alpha, Lf, Bo)
        dA, dB = _dev(torch, Lf), _dev(torch, B0)
        h.trsm_batch_strided_wsquery(side, m, n, batch)
        h.allocate_workspace()
        rc = h.trsm_batch_strided(side, "L", trans, "N", m, n, alpha, dA, lda, k * lda, dB, ldb, n * ldb, batch)
        torch.cuda.synchronize()
        assert rc == kb.KBLAS_Success, (it, side, trans, m, n)
        seen.add(h.last_kernel.split("<")[0])
        got = dB.cpu().numpy()
        scale = max(1.0, np.abs(Bo[:, :, :m]).max())
        assert np.abs(got[:, :, :m] - Bo[:, :, :m]).max() <= 100 * k * U.EPS[dt] * scale, (it, p, side, trans, m, n, lda, ldb, h.last_kernel)
        assert np.array_equal(got[:, :, m:], B0[:, :, m:]), "ldb padding untouched"
        if side == "R":
            # potrs on the same factor
            Bp = B0.copy()
            U.oracle_potrs("R", "L", m, n, Lf, Bp)
            dB2 = _dev(torch, B0)
            h.potrs_batch_strided_wsquery(m, n, batch)
            h.allocate_workspace()
            assert h.potrs_batch_strided("R", "L", m, n, dA, lda, k * lda, dB2, ldb, n * ldb, batch) == kb.KBLAS_Success
            torch.cuda.synchronize()
            g2 = dB2.cpu().numpy()
            assert np.abs(g2[:, :, :m] - Bp[:, :, :m]).max() <= 100 * k * U.EPS[dt] * max(1.0, np.abs(Bp[:, :, :m]).max()), (it, p, m, n, h.last_kernel)
            seen.add(h.last_kernel.split("<")[0])
    assert {"tri_reg", "tri_dual", "tri_small", "tri_blocked"} <= seen, seen


def test_trsm_potrs_posv_return_codes(env):
    kb, h, torch = env
    dA, dB = _dev(torch, U.rand_spd_batch(2, 8)), _dev(torch, U.rand_batch(2, 8, 8))
    a = (dA, 8, 64, dB, 8, 64, 2)
    assert h.trsm_batch_strided("X", "L", "N", "N", 8, 8, 1.0, *a) == kb.KBLAS_NotImplemented
    assert h.potrs_batch_strided("X", "L", 8, 8, *a) == kb.KBLAS_NotImplemented
    assert h.posv_batch_strided("X", "L", 8, 8, *a, None) == kb.KBLAS_NotImplemented
    # side L, uplo U and diag U are extensions here (the reference: KBLAS_NotImplemented, golden posv_?_left rc = -2): see
    # test_potrs_posv_left_side_extension and test_upper_and_unit_extensions


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("m,n", [(8, 8), (16, 16), (24, 24), (32, 32), (16, 32), (5, 13), (40, 32), (3, 1), (32, 16), (70, 16), (50, 24), (100, 8)])
def test_potrs_and_posv_strided_vs_oracle(env, p, m, n):
    kb, h, torch = env
    dt = DT[p]
    batch = 45
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n + 1)
    B0 = U.rand_batch(batch, m, n, dtype=dt, seed=n + 2)
    Ao, Bo = A0.copy(), B0.copy()
    U.oracle_posv("R", "L", m, n, Ao, Bo)
    tol = 100 * n * U.EPS[dt] * max(1.0, np.abs(Bo).max())
    # posv
    dA, dB = _dev(torch, A0), _dev(torch, B0)
    info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    h.posv_batch_strided_wsquery("R", m, n, batch)
    h.allocate_workspace()
    assert h.posv_batch_strided("R", "L", m, n, dA, n, n * n, dB, m, m * n, batch, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    _check_potrf(A0, dA.cpu().numpy(), n, dt, Lref=Ao)
    assert np.abs(dB.cpu().numpy() - Bo).max() <= tol
    assert (info.cpu().numpy() == SENT).all()
    # potrs from the oracle's factor
    dL, dB2 = _dev(torch, Ao), _dev(torch, B0)
    assert h.potrs_batch_strided("R", "L", m, n, dL, n, n * n, dB2, m, m * n, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert np.abs(dB2.cpu().numpy() - Bo).max() <= tol
    assert np.array_equal(dL.cpu().numpy(), Ao)


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("m,n", [(8, 8), (16, 3), (24, 24), (32, 32), (32, 100), (5, 13), (13, 40), (48, 16), (64, 33), (100, 7)])
def test_potrs_posv_left_side_extension(env, p, m, n):
    """side = 'L': (L L^T) X = B with A of order m, B m x n -- KBLAS_NotImplemented in the reference
    (Xpotrs_batch_drivers.cuh:40-43, Xposv_batch_drivers.cuh:41-44), implemented here.  Checked against
    the oracle's composition of the restated reference TRSMs, against LAPACK-style residuals, strided and pointer array."""
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    batch = 45
    A0 = U.rand_spd_batch(batch, m, lda=m + 1, dtype=dt, seed=m + 1)
    B0 = U.rand_batch(batch, m, n, ld=m + 2, dtype=dt, seed=n + 2)
    Ao, Bo = A0.copy(), B0.copy()
    assert U.oracle_posv("L", "L", m, n, Ao.copy(), Bo.copy()) == -2      # the reference: not implemented
    assert U.oracle_posv_left(m, n, Ao, Bo) == 1
    tol = 100 * m * U.EPS[dt] * max(1.0, np.abs(Bo[:, :, :m]).max())
    # the oracle's X really solves A X = B
    Am, Xm, Bm = U.as_mats(A0, m, m).astype(np.float64), U.as_mats(Bo, m, n).astype(np.float64), U.as_mats(B0, m, n).astype(np.float64)
    assert np.abs(Am @ Xm - Bm).max() <= 100 * m * U.EPS[dt] * np.abs(Am).max() * max(1.0, np.abs(Xm).max())
    # posv, strided
    dA, dB = _dev(torch, A0), _dev(torch, B0)
    h.posv_batch_strided_wsquery("L", m, n, batch)
    h.posv_batch_wsquery("L", m, n, batch)
    h.allocate_workspace()
    assert h.posv_batch_strided("L", "L", m, n, dA, m + 1, m * (m + 1), dB, m + 2, n * (m + 2), batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    _check_potrf(A0, dA.cpu().numpy(), m, dt, Lref=Ao)
    got = dB.cpu().numpy()
    assert np.abs(got[:, :, :m] - Bo[:, :, :m]).max() <= tol, h.last_kernel
    assert np.array_equal(got[:, :, m:], B0[:, :, m:]), "ldb padding untouched"
    # potrs from the oracle's factor, pointer array (shuffled)
    dL, dB2 = _dev(torch, Ao), _dev(torch, B0)
    perm = torch.randperm(batch, device="cuda")
    pa = (dL.data_ptr() + perm * (m * (m + 1) * es)).contiguous()
    pb = (dB2.data_ptr() + perm * (n * (m + 2) * es)).contiguous()
    assert h.potrs_batch("L", "L", m, n, pa, m + 1, pb, m + 2, batch, prec=p) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert np.abs(dB2.cpu().numpy()[:, :, :m] - Bo[:, :, :m]).max() <= tol, h.last_kernel
    assert np.array_equal(dL.cpu().numpy(), Ao), "factor is read-only"


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [64, 128, 256])
def test_posv_pointer_array_large_n(env, p, n):
    """BASELINE config 4: pointer-array posv, n = 64/128/256, 16 right-hand-side rows"""
    kb, h, torch = env
    dt = DT[p]
    m, batch = 16, 12
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n + 5)
    B0 = U.rand_batch(batch, m, n, dtype=dt, seed=n + 6)
    Ao, Bo = A0.copy(), B0.copy()
    U.oracle_posv("R", "L", m, n, Ao, Bo)
    dA, dB = _dev(torch, A0), _dev(torch, B0)
    perm = np.random.default_rng(n).permutation(batch).astype(np.int64)
    pa = torch.from_numpy(dA.data_ptr() + perm * n * n * A0.itemsize).cuda()
    pb = torch.from_numpy(dB.data_ptr() + perm * m * n * A0.itemsize).cuda()
    h.posv_batch_wsquery("R", m, n, batch)
    h.allocate_workspace()
    assert h.posv_batch("R", "L", m, n, pa, n, pb, m, batch, None, prec=p) == kb.KBLAS_Success
    torch.cuda.synchronize()
    _check_potrf(A0, dA.cpu().numpy(), n, dt, Lref=Ao)
    assert np.abs(dB.cpu().numpy() - Bo).max() <= 100 * n * U.EPS[dt] * max(1.0, np.abs(Bo).max())


# =============================================================================================
# golden vectors of the reference GPU library
def _gold():
    path = os.path.join(U.GOLDEN_DIR, "reference_gpu.npz")
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet")
    z = np.load(path)
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def _num(name, tag):
    for part in name.split("_"):
        if part.startswith(tag) and part[len(tag):].isdigit():
            return int(part[len(tag):])
    raise KeyError(name)


def test_golden_vectors_through_the_c_abi(env):
    kb, h, torch = env
    seen = 0
    for name, c in _gold().items():
        kind, p = name.split("_")[0], name.split("_")[1]
        if "A_in" not in c and "L_in" not in c:
            continue
        dt = DT[p]
        if kind == "potrf":
            n = _num(name, "n")
            lda = c["A_in"].shape[2]
            batch = c["A_in"].shape[0]
            dA = _dev(torch, c["A_in"])
            info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
            h.potrf_batch_strided_wsquery(n, batch)
            h.allocate_workspace()
            rc = h.potrf_batch_strided("L", n, dA, lda, n * lda, batch, info)
            torch.cuda.synchronize()
            assert rc == int(c["rc"]), name
            got = dA.cpu().numpy()
            if "nonspd" in name:
                assert np.array_equal(np.isfinite(np.tril(U.as_mats(got, n, n))), np.isfinite(np.tril(U.as_mats(c["A_out"], n, n)))), name
            else:
                _check_potrf(c["A_in"], got, n, dt, Lref=c["A_out"])
            assert np.array_equal(info.cpu().numpy(), c["info"]), name
        else:
            m, n = _num(name, "m"), _num(name, "n")
            batch = c["B_in"].shape[0]
            dB = _dev(torch, c["B_in"])
            if kind == "trsm":
                side, trans = name.split("_")[2]
                k = m if side == "L" else n
                dA = _dev(torch, c["A_in"])
                h.trsm_batch_strided_wsquery(side, m, n, batch)
                h.allocate_workspace()
                rc = h.trsm_batch_strided(side, "L", trans, "N", m, n, float(c["alpha"]), dA, k, k * k, dB, m, m * n, batch)
            elif kind == "potrs":
                k = n
                if n == 1:
                    continue  # documented deviation: the reference refuses n == 1, we solve it
                dA = _dev(torch, c["L_in"])
                h.potrs_batch_strided_wsquery(m, n, batch)
                h.allocate_workspace()
                rc = h.potrs_batch_strided("R", "L", m, n, dA, n, n * n, dB, m, m * n, batch)
            else:
                k = n
                dA = _dev(torch, c["A_in"])
                h.posv_batch_strided_wsquery("R", m, n, batch)
                h.allocate_workspace()
                rc = h.posv_batch_strided("R", "L", m, n, dA, n, n * n, dB, m, m * n, batch, None)
            torch.cuda.synchronize()
            assert rc == int(c["rc"]) == 1, name
            ref = c["B_out"]
            assert np.abs(dB.cpu().numpy() - ref).max() <= 100 * k * U.EPS[dt] * max(1.0, np.abs(ref).max()), name
        seen += 1
    assert seen >= 80


def test_golden_r2_vectors_through_the_c_abi(env):
    """reference_gpu_r2.npz: config-4 sizes (n = 128 / 256) and the POINTER-ARRAY entry points of all four routines,
    outputs of the unmodified reference library; pointer arrays rebuilt from the stored permutation."""
    kb, h, torch = env
    cases = U.load_golden("reference_gpu_r2.npz")
    if cases is None:
        pytest.skip("tests/golden/reference_gpu_r2.npz not generated yet")
    seen = 0
    for name in cases:
        kind, p = name.split("_")[0], name.split("_")[1]
        if kind == "trsmalpha0":
            continue        # pinned in tests/test_oracle.py; our (different, documented) behaviour in test_trsm_alpha_zero_*
        dt = DT[p]
        es = np.dtype(dt).itemsize
        eps = U.EPS[dt]
        c = U.golden_r2_inputs(cases, name)

        def ptr_array(d, elems):
            return torch.from_numpy(d.data_ptr() + c["perm"].astype(np.int64) * elems * es).cuda()

        if kind in ("potrfbig", "potrfptr"):
            n = _num(name, "n")
            batch = c["A_in"].shape[0]
            dA = _dev(torch, c["A_in"])
            info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
            if kind == "potrfbig":
                h.potrf_batch_strided_wsquery(n, batch)
                h.allocate_workspace()
                rc = h.potrf_batch_strided("L", n, dA, n, n * n, batch, info)
            else:
                h.potrf_batch_wsquery(n, batch)
                h.allocate_workspace()
                rc = h.potrf_batch("L", n, ptr_array(dA, n * n), n, batch, info, prec=p)
            torch.cuda.synchronize()
            assert rc == int(c["rc"]) == 1, name
            got = dA.cpu().numpy()
            if kind == "potrfbig":
                assert U.potrf_residual(c["A_in"], got, n) <= 10 * n * eps
                assert np.abs(U.pack_lower(got, n) - c["L_out_packed"]).max() <= 100 * n * eps * np.abs(c["A_in"]).max(), name
                assert np.array_equal(np.triu(U.as_mats(got, n, n), 1), np.triu(U.as_mats(c["A_in"], n, n), 1))
            else:
                _check_potrf(c["A_in"], got, n, dt, Lref=c["A_out"])
            assert np.array_equal(info.cpu().numpy(), c["info"]), name
        else:
            m, n = _num(name, "m"), _num(name, "n")
            batch = c["B_in"].shape[0]
            dB = _dev(torch, c["B_in"])
            if kind == "posvbig":
                dA = _dev(torch, c["A_in"])
                h.posv_batch_strided_wsquery("R", m, n, batch)
                h.allocate_workspace()
                rc = h.posv_batch_strided("R", "L", m, n, dA, n, n * n, dB, m, m * n, batch, None)
                k = n
            elif kind == "trsmptr":
                side, trans = name.split("_")[2]
                k = m if side == "L" else n
                dA = _dev(torch, c["L_in"])
                h.trsm_batch_wsquery(side, m, n, batch)
                h.allocate_workspace()
                rc = h.trsm_batch(side, "L", trans, "N", m, n, float(c["alpha"]), ptr_array(dA, k * k), k, ptr_array(dB, m * n), m,
                                  batch, prec=p)
            elif kind == "potrsptr":
                k = n
                dA = _dev(torch, c["L_in"])
                h.potrs_batch_wsquery(m, n, batch)
                h.allocate_workspace()
                rc = h.potrs_batch("R", "L", m, n, ptr_array(dA, n * n), n, ptr_array(dB, m * n), m, batch, prec=p)
            else:
                assert kind == "posvptr", name
                k = n
                dA = _dev(torch, c["A_in"])
                h.posv_batch_wsquery("R", m, n, batch)
                h.allocate_workspace()
                rc = h.posv_batch("R", "L", m, n, ptr_array(dA, n * n), n, ptr_array(dB, m * n), m, batch, None, prec=p)
            torch.cuda.synchronize()
            assert rc == int(c["rc"]) == 1, name
            ref = c["B_out"]
            assert np.abs(dB.cpu().numpy() - ref).max() <= 100 * k * eps * max(1.0, np.abs(ref).max()), name
            if kind == "posvptr":
                _check_potrf(c["A_in"], dA.cpu().numpy(), n, dt, Lref=c["A_out"])
            if kind == "posvbig":
                assert np.abs(U.pack_lower(dA.cpu().numpy(), n) - c["L_out_packed"]).max() <= 100 * n * eps * np.abs(c["A_in"]).max()
        seen += 1
    assert seen >= 50, seen


# =============================================================================================
# live A/B against the unmodified reference library on identical device buffers
@pytest.mark.parametrize("p", ["D", "S"])
def test_live_against_reference_library(env, p):
    if not U.have_ref():
        pytest.skip("oracle/_ref/libkblas_ref.so not built (needs /root/reference at build time)")
    kb, h, torch = env
    dt = DT[p]
    ct = C.c_double if p == "D" else C.c_float
    ref = U.RefLib()
    H, i, l, c, P = ref.H, ref.i, ref.l, ref.c, ref.P
    r_potrf = ref.fn(f"kblas{p}potrf_batch_strided", [H, c, i, P, i, l, i, P])
    r_potrs = ref.fn(f"kblas{p}potrs_batch_strided", [H, c, c, i, i, P, i, l, P, i, l, i])
    r_trsm = ref.fn(f"kblas{p}trsm_batch_strided", [H, c, c, c, c, i, i, ct, P, i, l, P, i, l, i])
    batch = 4099
    for n in (8, 16, 24, 32):
        m = n
        A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n)
        B0 = U.rand_batch(batch, m, n, dtype=dt, seed=n + 1)
        mine_A, ref_A = _dev(torch, A0), _dev(torch, A0)
        mine_B, ref_B = _dev(torch, B0), _dev(torch, B0)
        mine_T, ref_T = _dev(torch, B0), _dev(torch, B0)
        info_m = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
        info_r = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
        ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
        ref.wsquery("kblas_trsm_batch_strided_wsquery", "ciii", b"L", m, n, batch)
        ref.allocate()
        h.posv_batch_strided_wsquery("R", m, n, batch)
        h.allocate_workspace()
        rc_r = r_potrf(ref.h, b"L", n, ref_A.data_ptr(), n, n * n, batch, info_r.data_ptr())
        rc_m = h.potrf_batch_strided("L", n, mine_A, n, n * n, batch, info_m)
        assert rc_r == rc_m == 1
        rc_r = r_potrs(ref.h, b"R", b"L", m, n, ref_A.data_ptr(), n, n * n, ref_B.data_ptr(), m, m * n, batch)
        rc_m = h.potrs_batch_strided("R", "L", m, n, ref_A, n, n * n, mine_B, m, m * n, batch)
        assert rc_r == rc_m == 1
        rc_r = r_trsm(ref.h, b"L", b"L", b"N", b"N", m, n, 0.28, ref_A.data_ptr(), n, n * n, ref_T.data_ptr(), m, m * n, batch)
        rc_m = h.trsm_batch_strided("L", "L", "N", "N", m, n, 0.28, ref_A, n, n * n, mine_T, m, m * n, batch)
        assert rc_r == rc_m == 1
        torch.cuda.synchronize()
        _check_potrf(A0, mine_A.cpu().numpy(), n, dt, Lref=ref_A.cpu().numpy())
        assert torch.equal(info_m, info_r)
        for mine, theirs in ((mine_B, ref_B), (mine_T, ref_T)):
            t = theirs.cpu().numpy()
            assert np.abs(mine.cpu().numpy() - t).max() <= 100 * n * U.EPS[dt] * max(1.0, np.abs(t).max())
    ref.close()


@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [64, 128, 256])
def test_live_large_n_against_reference_library(env, p, n):
    """BASELINE config 4 sizes, live against the unmodified reference library on identical device buffers: potrf
    (strided + pointer array) and posv with m = 16 right-hand-side rows (pointer array, shuffled).  The reference goes
    through cuBLAS batched GEMM here, so equality is tolerance-based: |L - L_ref| <= 100 n eps ||A||."""
    if not U.have_ref():
        pytest.skip("oracle/_ref/libkblas_ref.so not built (needs /root/reference at build time)")
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    ref = U.RefLib()
    H, i, l, c, P = ref.H, ref.i, ref.l, ref.c, ref.P
    r_potrf = ref.fn(f"kblas{p}potrf_batch_strided", [H, c, i, P, i, l, i, P])
    r_potrf_p = ref.fn(f"kblas{p}potrf_batch", [H, c, i, P, i, i, P])
    r_posv_p = ref.fn(f"kblas{p}posv_batch", [H, c, c, i, i, P, i, P, i, i, P])
    m, batch = 16, 61
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n)
    B0 = U.rand_batch(batch, m, n, dtype=dt, seed=n + 1)
    ref.wsquery("kblas_posv_batch_wsquery", "ciii", b"R", m, n, batch)
    ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
    ref.allocate()
    h.posv_batch_wsquery("R", m, n, batch)
    h.posv_batch_strided_wsquery("R", m, n, batch)
    h.allocate_workspace()
    perm = np.random.default_rng(n).permutation(batch).astype(np.int64)
    # strided potrf
    mine_A, ref_A = _dev(torch, A0), _dev(torch, A0)
    assert r_potrf(ref.h, b"L", n, ref_A.data_ptr(), n, n * n, batch, None) == 1
    assert h.potrf_batch_strided("L", n, mine_A, n, n * n, batch, None) == 1
    torch.cuda.synchronize()
    _check_potrf(A0, mine_A.cpu().numpy(), n, dt, Lref=ref_A.cpu().numpy())
    # pointer-array potrf
    mine_A, ref_A = _dev(torch, A0), _dev(torch, A0)
    pm = torch.from_numpy(mine_A.data_ptr() + perm * n * n * es).cuda()
    pr = torch.from_numpy(ref_A.data_ptr() + perm * n * n * es).cuda()
    assert r_potrf_p(ref.h, b"L", n, pr.data_ptr(), n, batch, None) == 1
    assert h.potrf_batch("L", n, pm, n, batch, None, prec=p) == 1
    torch.cuda.synchronize()
    _check_potrf(A0, mine_A.cpu().numpy(), n, dt, Lref=ref_A.cpu().numpy())
    # pointer-array posv
    mine_A, ref_A = _dev(torch, A0), _dev(torch, A0)
    mine_B, ref_B = _dev(torch, B0), _dev(torch, B0)
    pm = torch.from_numpy(mine_A.data_ptr() + perm * n * n * es).cuda()
    pr = torch.from_numpy(ref_A.data_ptr() + perm * n * n * es).cuda()
    pmb = torch.from_numpy(mine_B.data_ptr() + perm * m * n * es).cuda()
    prb = torch.from_numpy(ref_B.data_ptr() + perm * m * n * es).cuda()
    info_m = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    info_r = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
    assert r_posv_p(ref.h, b"R", b"L", m, n, pr.data_ptr(), n, prb.data_ptr(), m, batch, info_r.data_ptr()) == 1
    assert h.posv_batch("R", "L", m, n, pm, n, pmb, m, batch, info_m, prec=p) == 1
    torch.cuda.synchronize()
    _check_potrf(A0, mine_A.cpu().numpy(), n, dt, Lref=ref_A.cpu().numpy())
    t = ref_B.cpu().numpy()
    assert np.abs(mine_B.cpu().numpy() - t).max() <= 100 * n * U.EPS[dt] * max(1.0, np.abs(t).max())
    assert torch.equal(info_m, info_r)
    ref.close()


# =============================================================================================
# BASELINE.json full sizes through size-independent properties
@pytest.mark.parametrize("p,n,batch", [("D", 32, 1 << 20), ("D", 8, 1 << 20), ("D", 16, 1 << 20), ("D", 24, 1 << 20), ("S", 32, 1 << 20)])
def test_full_size_potrf_potrs_properties(env, p, n, batch):
    """config 2: strided potrf + potrs, batch = 1M.  Checked by residual / solve-residual on slices
    spread over the batch (first, middle, last CTAs) and by finiteness + untouched upper over all of it."""
    kb, h, torch = env
    tdt = torch.float64 if p == "D" else torch.float32
    eps = U.EPS[DT[p]]
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand((batch, n, n), generator=g, device="cuda", dtype=tdt)
    A = torch.tril(A) + torch.tril(A, -1).transpose(1, 2)
    A.diagonal(dim1=1, dim2=2).add_(n)
    A0 = A.clone()
    B = torch.rand((batch, n, n), generator=g, device="cuda", dtype=tdt)   # m = n right-hand-side rows
    B0 = B.clone()
    h.posv_batch_strided_wsquery("R", n, n, batch)
    h.allocate_workspace()
    assert h.potrf_batch_strided("L", n, A, n, n * n, batch, None) == kb.KBLAS_Success
    assert h.potrs_batch_strided("R", "L", n, n, A, n, n * n, B, n, n * n, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert bool(torch.isfinite(A).all()) and bool(torch.isfinite(B).all())
    # memory layout [b, col, row]: the strict upper triangle of the matrix is torch's strict LOWER of A[b]
    assert torch.equal(torch.tril(A, -1), torch.tril(A0, -1)), "strict upper triangle modified"
    for lo in (0, batch // 2 - 2048, batch - 4096):
        sl = slice(lo, lo + 4096)
        Lm = torch.triu(A[sl]).transpose(1, 2).double()          # math-layout lower factor
        Am = A0[sl].transpose(1, 2).double()
        R = Am - Lm @ Lm.transpose(1, 2)
        res = (R.flatten(1).norm(dim=1) / Am.flatten(1).norm(dim=1)).max().item()
        assert res <= 10 * n * eps, (lo, res)
        Xm, Bm = B[sl].transpose(1, 2).double(), B0[sl].transpose(1, 2).double()
        R2 = Xm @ Am - Bm                                          # X A = B
        res2 = (R2.flatten(1).norm(dim=1) / (Am.flatten(1).norm(dim=1) * Xm.flatten(1).norm(dim=1))).max().item()
        assert res2 <= 10 * n * eps, (lo, res2)


@pytest.mark.parametrize("p", ["D", "S"])
def test_full_size_config3_trsm_left(env, p):
    """config 3: strided trsm side L, uplo L, n = 32, nrhs = 32, batch = 1M, fp32 and fp64; residual
    ||L X - alpha B|| / (||L|| ||X||) on slices, finiteness over the whole batch."""
    kb, h, torch = env
    n, batch, alpha = 32, 1 << 20, 0.28
    tdt = torch.float64 if p == "D" else torch.float32
    eps = U.EPS[DT[p]]
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand((batch, n, n), generator=g, device="cuda", dtype=tdt)
    A.diagonal(dim1=1, dim2=2).add_(n)            # memory [b, col, row]: lower factor = torch.triu of A[b]
    B = torch.rand((batch, n, n), generator=g, device="cuda", dtype=tdt)
    B0 = B.clone()
    h.trsm_batch_strided_wsquery("L", n, n, batch)
    h.allocate_workspace()
    for trans in ("N", "T"):
        B.copy_(B0)
        assert h.trsm_batch_strided("L", "L", trans, "N", n, n, alpha, A, n, n * n, B, n, n * n, batch) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert bool(torch.isfinite(B).all())
        for lo in (0, batch // 2 - 2048, batch - 4096):
            sl = slice(lo, lo + 4096)
            Lm = torch.triu(A[sl]).transpose(1, 2).double()
            if trans == "T":
                Lm = Lm.transpose(1, 2)
            Xm, Bm = B[sl].transpose(1, 2).double(), B0[sl].transpose(1, 2).double()
            R = Lm @ Xm - alpha * Bm
            res = (R.flatten(1).norm(dim=1) / (Lm.flatten(1).norm(dim=1) * Xm.flatten(1).norm(dim=1))).max().item()
            assert res <= 10 * n * eps, (trans, lo, res)


@pytest.mark.parametrize("n", [64, 128, 256])
def test_full_size_config4_posv_pointer_array(env, n):
    """config 4: pointer-array dposv, n = 64 / 128 / 256, 16 right-hand-side rows, batch = 64K (shuffled pointers);
    factor residual and solve residual X A = B on slices, finiteness + untouched strict upper over all of it."""
    kb, h, torch = env
    m, batch = 16, 1 << 16
    eps = U.EPS[np.float64]
    g = torch.Generator(device="cuda").manual_seed(n)
    A = torch.empty((batch, n, n), device="cuda", dtype=torch.float64)
    for lo in range(0, batch, 1 << 13):
        a = torch.rand((1 << 13, n, n), generator=g, device="cuda", dtype=torch.float64)
        a = torch.tril(a) + torch.tril(a, -1).transpose(1, 2)
        a.diagonal(dim1=1, dim2=2).add_(n)
        A[lo:lo + (1 << 13)] = a
    A0 = A.clone()
    B = torch.rand((batch, n, m), generator=g, device="cuda", dtype=torch.float64)
    B0 = B.clone()
    perm = torch.randperm(batch, device="cuda")
    pa = (A.data_ptr() + perm * (n * n * 8)).contiguous()
    pb = (B.data_ptr() + perm * (m * n * 8)).contiguous()
    h.posv_batch_wsquery("R", m, n, batch)
    h.allocate_workspace()
    assert h.posv_batch("R", "L", m, n, pa, n, pb, m, batch, None, prec="D") == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert bool(torch.isfinite(B).all())
    for lo in range(0, batch, 1 << 12):      # chunked: the n = 256 batch is 32 GiB, twice with the pristine copy
        sl = slice(lo, lo + (1 << 12))
        assert bool(torch.isfinite(A[sl]).all())
        assert torch.equal(torch.tril(A[sl], -1), torch.tril(A0[sl], -1)), "strict upper triangle modified"
    for lo in (0, batch // 2 - 256, batch - 512):
        sl = slice(lo, lo + 512)
        Lm = torch.triu(A[sl]).transpose(1, 2)
        Am = A0[sl].transpose(1, 2)
        R = Am - Lm @ Lm.transpose(1, 2)
        res = (R.flatten(1).norm(dim=1) / Am.flatten(1).norm(dim=1)).max().item()
        assert res <= 10 * n * eps, (lo, res)
        Xm, Bm = B[sl].transpose(1, 2), B0[sl].transpose(1, 2)
        R2 = Xm @ Am - Bm
        res2 = (R2.flatten(1).norm(dim=1) / (Am.flatten(1).norm(dim=1) * Xm.flatten(1).norm(dim=1))).max().item()
        assert res2 <= 10 * n * eps, (lo, res2)


def test_multi_stream_and_timer(env):
    kb, h, torch = env
    s = torch.cuda.Stream()
    h.set_stream(s)
    assert h.get_stream() == s.cuda_stream
    n, batch = 32, 5000
    A0 = U.rand_spd_batch(batch, n, seed=4)
    dA = _dev(torch, A0)
    s.wait_stream(torch.cuda.current_stream())
    h.timer_tic()
    assert h.potrf_batch_strided("L", n, dA, n, n * n, batch, None) == kb.KBLAS_Success
    h.timer_record_end()
    sec = h.timer_toc()
    assert 0 < sec < 1.0
    s.synchronize()
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    _check_potrf(A0, dA.cpu().numpy(), n, np.float64, Lref=Lo)
    h.set_stream(0)


def test_two_devices_one_process(env):
    """one handle per device from ONE process (what the reference harness does with --ngpu, test_Xpotrf_batch.cpp:97-160):
    kernels that need a raised dynamic shared-memory limit must work on every device, not only the first."""
    kb, _, torch = env
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, m, batch = 32, 32, 257
    A0 = U.rand_spd_batch(batch, n, seed=3)
    B0 = U.rand_batch(batch, m, n, seed=4)
    Ao, Bo = A0.copy(), B0.copy()
    U.oracle_posv("R", "L", m, n, Ao, Bo)
    try:
        for dev in (0, 1):
            torch.cuda.set_device(dev)
            h = kb.Handle()
            dA, dB = torch.from_numpy(A0).cuda(dev), torch.from_numpy(B0).cuda(dev)
            h.posv_batch_strided_wsquery("R", m, n, batch)
            h.allocate_workspace()
            assert h.posv_batch_strided("R", "L", m, n, dA, n, n * n, dB, m, m * n, batch, None) == kb.KBLAS_Success
            torch.cuda.synchronize(dev)
            assert "tri_dual" in h.last_kernel
            assert np.abs(dB.cpu().numpy() - Bo).max() <= 100 * n * U.EPS[np.float64] * max(1.0, np.abs(Bo).max())
            # n = 256 potrf: the tensor-path panel kernel also raises its shared-memory limit
            P = U.rand_spd_batch(9, 256, seed=5)
            dP = torch.from_numpy(P).cuda(dev)
            h.potrf_batch_strided_wsquery(256, 9)
            h.allocate_workspace()
            assert h.potrf_batch_strided("L", 256, dP, 256, 256 * 256, 9, None) == kb.KBLAS_Success
            torch.cuda.synchronize(dev)
            assert U.potrf_residual(P, dP.cpu().numpy(), 256) <= 10 * 256 * U.EPS[np.float64]
            h.destroy()
    finally:
        torch.cuda.set_device(0)
