"""Packed lower-triangular batch layout (kblasx?pptrf_batch, tri_pack / tri_unpack): SURVEY.md §8(f)4.

The reference has no packed routine, so the anchor is   pptrf(pack(A)) == pack(potrf(A))   with potrf the pinned path:
* CPU: the packed oracle against the (golden-pinned) potrf oracle and against LAPACK's own packed Cholesky ?pptrf;
* GPU: the CUDA path against the packed oracle (100 n eps ||A||), residual (10 n eps), BIT-IDENTICAL to
  kblas?potrf_batch_strided for n % 8 == 0, every data-movement variant (plain / TMA in / TMA in+out), unaligned
  strides, pointer arrays, ragged n, LAPACK-info mode, pack/unpack round trips, and the full 2^20 batch.
"""
import os

import numpy as np
import pytest

from tests import _util as U

DT = {"D": np.float64, "S": np.float32}
SENT = 77


# ---------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 5, 8, 13, 16, 24, 31, 32, 40, 64])
def test_packed_oracle_matches_potrf_oracle_and_lapack_pptrf(dt, n):
    from scipy.linalg import lapack

    A = U.rand_spd_batch(6, n, dtype=dt, seed=n)
    P = U.pack_lower(A, n)
    assert P.shape == (6, n * (n + 1) // 2)
    assert U.oracle_pptrf(P, n) == 1
    L = A.copy()
    U.oracle_potrf(L, n)
    assert np.array_equal(P, U.pack_lower(L, n)), "packed oracle == pack(potrf oracle), bit for bit"
    pptrf = lapack.dpptrf if dt == np.float64 else lapack.spptrf
    eps = U.EPS[dt]
    for b in range(6):
        ref, info = pptrf(n, U.pack_lower(A, n)[b], lower=1)
        assert info == 0
        assert np.abs(ref - P[b]).max() <= 100 * n * eps * np.abs(A[b]).max()
    assert U.oracle_pptrf(P, n, uplo="U") == -2


def test_pack_unpack_helpers_roundtrip():
    A = U.rand_spd_batch(3, 7, lda=9, extra_cols=1)
    P = U.pack_lower(A, 7)
    B = U.unpack_lower(P, 7)
    assert np.array_equal(np.tril(U.as_mats(B, 7, 7)), np.tril(U.as_mats(A, 7, 7)))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    kb = U.kblas()
    h = kb.Handle()
    yield kb, h, torch
    h.destroy()


def _check(A0, P_out, n, dt, Pref):
    eps = U.EPS[dt]
    Lf = U.unpack_lower(P_out, n)
    assert U.potrf_residual(A0[:, :n, :n].copy(), Lf, n) <= 10 * n * eps
    assert np.abs(P_out - Pref).max() <= 100 * n * eps * np.abs(A0).max()


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 13, 16, 17, 23, 24, 25, 31, 32])
def test_pptrf_strided_vs_oracle(env, p, n):
    kb, h, torch = env
    dt = DT[p]
    sz = n * (n + 1) // 2
    al = 16 // np.dtype(dt).itemsize
    for batch, pad in ((37, 0), (1, 3), (1000, 0), (131, (-sz) % al + al)):
        A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n + batch)
        P0 = np.full((batch, sz + pad), -7.25, dtype=dt)
        P0[:, :sz] = U.pack_lower(A0, n)
        dP = torch.from_numpy(P0).cuda()
        info = torch.full((batch,), SENT, dtype=torch.int32, device="cuda")
        rc = h.pptrf_batch_strided("L", n, dP, sz + pad, batch, info)
        torch.cuda.synchronize()
        assert rc == kb.KBLAS_Success
        got = dP.cpu().numpy()
        Pref = P0[:, :sz].copy()
        U.oracle_pptrf(Pref, n)
        _check(A0, got[:, :sz], n, dt, Pref)
        assert np.array_equal(got[:, sz:], P0[:, sz:]), "stride padding untouched"
        assert (info.cpu().numpy() == SENT).all(), "info must not be written (potrf_batch parity)"
        # same bits as the drop-in routine on full storage (same arithmetic, n % 8 == 0 and the generic path alike)
        dA = torch.from_numpy(A0).cuda()
        h.potrf_batch_strided_wsquery(n, batch)
        h.allocate_workspace()
        assert h.potrf_batch_strided("L", n, dA, n, n * n, batch, None) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert np.array_equal(got[:, :sz], U.pack_lower(dA.cpu().numpy(), n)), (batch, pad, h.last_kernel)


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("variant", [20, 21, 22, 23])
@pytest.mark.parametrize("n", [8, 16, 24, 32])
def test_pptrf_every_data_movement_variant(p, variant, n, monkeypatch):
    """plain loads/stores, TMA bulk loads, TMA bulk loads + bulk stores, with / without the per-batch CTA barrier:
    identical results, batch sizes around the persistent-grid boundaries (tail warp-batches, several rounds per CTA)"""
    import torch

    kb = U.kblas()
    monkeypatch.setenv("KBLAS_B200_VARIANT", str(variant))
    h = kb.Handle()
    dt = DT[p]
    sz = n * (n + 1) // 2
    want_tag = {20: "<ldg,stg>", 21: "<tma-in,stg>", 22: "lockstep", 23: "<tma-in,tma-out>"}[variant]
    for batch in (1, 3, 4, 5, 31, 148 * 32 + 1, 40001):
        A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=batch % 97 + n)
        P0 = U.pack_lower(A0, n)
        dP = torch.from_numpy(P0).cuda()
        assert h.pptrf_batch_strided("L", n, dP, sz, batch, None) == kb.KBLAS_Success
        torch.cuda.synchronize()
        assert want_tag in h.last_kernel, h.last_kernel
        Pref = P0.copy()
        U.oracle_pptrf(Pref, n)
        _check(A0, dP.cpu().numpy(), n, dt, Pref)
    h.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n", [8, 20, 32])
def test_pptrf_pointer_array_and_info_mode(env, p, n, monkeypatch):
    kb, h, torch = env
    dt = DT[p]
    es = np.dtype(dt).itemsize
    sz = n * (n + 1) // 2
    batch = 531
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=3 * n)
    P0 = U.pack_lower(A0, n)
    for off in (0, 1):     # off = 1: matrices start one element past a 16-byte boundary
        d = torch.zeros(P0.size + off + 4, dtype=getattr(torch, np.dtype(dt).name), device="cuda")
        d[off:off + P0.size] = torch.from_numpy(P0).cuda().flatten()
        perm = torch.randperm(batch, device="cuda")
        ptrs = (d.data_ptr() + (off + perm * sz) * es).contiguous()
        assert h.pptrf_batch("L", n, ptrs, batch, None, prec=p) == kb.KBLAS_Success
        torch.cuda.synchronize()
        Pref = P0.copy()
        U.oracle_pptrf(Pref, n)
        _check(A0, d[off:off + P0.size].cpu().numpy().reshape(P0.shape), n, dt, Pref)
        assert float(d[:off].abs().sum()) == 0 and float(d[off + P0.size:].abs().sum()) == 0
    # LAPACK info semantics are opt-in, as for potrf
    monkeypatch.setenv("KBLAS_B200_INFO_MODE", "lapack")
    h2 = kb.Handle()
    A1 = A0[:20].copy()
    A1[3, 5 % n, 5 % n] = -3.0
    A1[17, n - 1, n - 1] = -1.0
    dP = torch.from_numpy(U.pack_lower(A1, n)).cuda()
    info = torch.full((20,), SENT, dtype=torch.int32, device="cuda")
    assert h2.pptrf_batch_strided("L", n, dP, sz, 20, info) == kb.KBLAS_Success
    torch.cuda.synchronize()
    want = np.zeros(20, dtype=np.int32)
    want[3], want[17] = 5 % n + 1, n
    assert np.array_equal(info.cpu().numpy(), want)
    h2.destroy()


@pytest.mark.gpu
def test_pptrf_return_codes_and_pack_unpack(env):
    kb, h, torch = env
    n, batch = 16, 9
    sz = n * (n + 1) // 2
    A0 = U.rand_spd_batch(batch, n, lda=n + 3, extra_cols=2, seed=5)
    dA = torch.from_numpy(A0).cuda()
    dP = torch.zeros((batch, sz + 1), dtype=torch.float64, device="cuda")
    assert h.tri_pack_batch_strided("L", n, dA, n + 3, (n + 2) * (n + 3), dP, sz + 1, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert np.array_equal(dP.cpu().numpy()[:, :sz], U.pack_lower(A0, n))
    before = dP.clone()
    assert h.pptrf_batch_strided("U", n, dP, sz + 1, batch, None) == kb.KBLAS_NotImplemented
    assert h.pptrf_batch_strided("L", 33, dP, 33 * 17, batch, None) == kb.KBLAS_NotImplemented
    assert h.pptrf_batch_strided("L", n, dP, sz + 1, 0, None) == kb.KBLAS_UnknownError
    assert h.pptrf_batch_strided("L", n, dP, sz - 1, batch, None) == kb.KBLAS_Error_WrongInput
    assert h.pptrf_batch_strided("L", 0, dP, sz + 1, batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert torch.equal(dP, before)
    # factor packed, unpack into a sentinel-filled full array: only the lower triangle is written
    assert h.pptrf_batch_strided("L", n, dP, sz + 1, batch, None) == kb.KBLAS_Success
    out = torch.full_like(dA, 9.5)
    assert h.tri_unpack_batch_strided("L", n, dP, sz + 1, out, n + 3, (n + 2) * (n + 3), batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    O = out.cpu().numpy()
    Lo = A0.copy()
    U.oracle_potrf(Lo, n)
    M, W = U.as_mats(O, n, n), U.as_mats(Lo, n, n)
    assert np.abs(np.tril(M) - np.tril(W)).max() <= 100 * n * U.EPS[np.float64] * np.abs(A0).max()
    i, j = np.indices((n, n))
    assert (M[:, j > i] == 9.5).all() and (O[:, :, n:] == 9.5).all() and (O[:, n:, :] == 9.5).all()


@pytest.mark.gpu
@pytest.mark.parametrize("p,n", [("D", 32), ("D", 16), ("D", 8), ("S", 32), ("D", 24)])
def test_full_size_pptrf_properties(env, p, n):
    """2^20 packed matrices: pack on the device, factor, compare with kblas?potrf_batch_strided on full storage (bit
    for bit, whole batch), residual on slices."""
    kb, h, torch = env
    batch = 1 << 20
    tdt = torch.float64 if p == "D" else torch.float32
    eps = U.EPS[DT[p]]
    sz = n * (n + 1) // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand((batch, n, n), generator=g, device="cuda", dtype=tdt)
    A = torch.tril(A) + torch.tril(A, -1).transpose(1, 2)
    A.diagonal(dim1=1, dim2=2).add_(n)
    A0 = A.clone()
    P = torch.empty((batch, sz), device="cuda", dtype=tdt)
    assert h.tri_pack_batch_strided("L", n, A, n, n * n, P, sz, batch) == kb.KBLAS_Success
    assert h.pptrf_batch_strided("L", n, P, sz, batch, None) == kb.KBLAS_Success
    h.potrf_batch_strided_wsquery(n, batch)
    h.allocate_workspace()
    assert h.potrf_batch_strided("L", n, A, n, n * n, batch, None) == kb.KBLAS_Success
    P2 = torch.empty_like(P)
    assert h.tri_pack_batch_strided("L", n, A, n, n * n, P2, sz, batch) == kb.KBLAS_Success
    torch.cuda.synchronize()
    assert bool(torch.isfinite(P).all())
    assert torch.equal(P, P2), "packed and full-storage factors differ"
    for lo in (0, batch // 2 - 2048, batch - 4096):
        sl = slice(lo, lo + 4096)
        Lm = torch.triu(A[sl]).transpose(1, 2).double()
        Am = A0[sl].transpose(1, 2).double()
        R = Am - Lm @ Lm.transpose(1, 2)
        assert (R.flatten(1).norm(dim=1) / Am.flatten(1).norm(dim=1)).max().item() <= 10 * n * eps


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["D", "S"])
@pytest.mark.parametrize("n,pad", [(32, 0), (8, 0), (20, 3), (16, 2)])
def test_pptrf_host_pipeline(env, p, n, pad, monkeypatch):
    """kblasx?pptrf_batch_strided_host: packed matrices in host memory through the chunked 3-stream pipeline ==
    H2D + kblasx?pptrf_batch_strided + D2H, bit for bit, in place and out of place; the last chunk stops at the last
    element of the last matrix (minimal strided allocation)."""
    kb, h, torch = env
    dt = DT[p]
    batch = 301
    sz = n * (n + 1) // 2
    A0 = U.rand_spd_batch(batch, n, dtype=dt, seed=n + 5)
    P0 = np.full((batch, sz + pad), -7.25, dtype=dt)
    P0[:, :sz] = U.pack_lower(A0, n)
    dP = torch.from_numpy(P0).cuda()
    assert h.pptrf_batch_strided("L", n, dP, sz + pad, batch, None) == kb.KBLAS_Success
    torch.cuda.synchronize()
    want = dP.cpu().numpy()
    monkeypatch.setenv("KBLAS_B200_HOSTCHUNK_MB", str(6.5 * (sz + pad) * np.dtype(dt).itemsize / (1 << 20)))
    flat = P0.flatten()[: (batch - 1) * (sz + pad) + sz].copy()     # minimal allocation: nothing behind the last matrix
    assert h.pptrf_batch_strided_host("L", n, flat, flat, sz + pad, batch) == kb.KBLAS_Success
    assert np.array_equal(flat, want.flatten()[: flat.size])
    src, out = P0.copy(), np.full_like(P0, 9.5)
    assert h.pptrf_batch_strided_host("L", n, src, out, sz + pad, batch) == kb.KBLAS_Success
    assert np.array_equal(src, P0), "AP_in is read-only"
    assert np.array_equal(out[:, :sz], want[:, :sz])
    assert h.pptrf_batch_strided_host("U", n, src, out, sz + pad, batch) == kb.KBLAS_NotImplemented
