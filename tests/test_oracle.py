"""CPU: the oracle (oracle/kblas_oracle.c) against LAPACK/numpy and against the committed golden
vectors produced by the unmodified reference GPU library (tests/golden/reference_gpu.npz)."""
import os

import numpy as np
import pytest

from tests import _util as U

DTYPES = [np.float64, np.float32]


def _tol(dt, n, c):
    return c * n * U.EPS[dt]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 48, 64, 100])
def test_potrf_matches_lapack(dt, n):
    A0 = U.rand_spd_batch(6, n, lda=n + 2, dtype=dt, seed=n)
    A = A0.copy()
    assert U.oracle_potrf(A, n) == 1
    # residual bar of BASELINE.json: ||A - L L^T|| / ||A|| <= 10 n eps
    assert U.potrf_residual(A0, A, n) <= _tol(dt, n, 10)
    # element-wise against LAPACK's factor: <= 100 n eps ||A||
    L = np.tril(U.as_mats(A, n, n))
    Lref = np.linalg.cholesky(U.as_mats(A0, n, n).astype(np.float64))
    normA = np.abs(U.as_mats(A0, n, n)).max()
    assert np.abs(L - Lref).max() <= _tol(dt, n, 100) * normA
    # strict upper triangle and the padding rows/cols are untouched
    Up0, Up1 = np.triu(U.as_mats(A0, n, n), 1), np.triu(U.as_mats(A, n, n), 1)
    assert np.array_equal(Up0, Up1)
    assert np.array_equal(A0[:, :, n:], A[:, :, n:])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("side,trans", [("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")])
@pytest.mark.parametrize("m,n", [(1, 1), (8, 8), (16, 16), (13, 7), (7, 13), (32, 32), (20, 33), (33, 20), (64, 16)])
def test_trsm_matches_numpy(dt, side, trans, m, n):
    k = m if side == "L" else n
    A = U.rand_spd_batch(4, k, dtype=dt, seed=k)
    B0 = U.rand_batch(4, m, n, ld=m + 1, dtype=dt, seed=m * 100 + n)
    B = B0.copy()
    alpha = 0.28
    assert U.oracle_trsm(side, "L", trans, "N", m, n, alpha, A, B) == 1
    Lm = np.tril(U.as_mats(A, k, k)).astype(np.float64)
    op = Lm if trans == "N" else np.transpose(Lm, (0, 2, 1))
    Bm = U.as_mats(B0, m, n).astype(np.float64) * alpha
    X = np.linalg.solve(op, Bm) if side == "L" else np.transpose(
        np.linalg.solve(np.transpose(op, (0, 2, 1)), np.transpose(Bm, (0, 2, 1))), (0, 2, 1))
    got = U.as_mats(B, m, n)
    assert np.abs(got - X).max() <= _tol(dt, k, 100) * max(1.0, np.abs(X).max())
    assert np.array_equal(B0[:, :, m:], B[:, :, m:])  # padding rows untouched


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("m,n", [(3, 1), (8, 8), (16, 16), (5, 13), (32, 32), (16, 24), (16, 64), (16, 100)])
def test_potrs_posv_match_numpy(dt, m, n):
    A0 = U.rand_spd_batch(3, n, dtype=dt, seed=n + 1)
    B0 = U.rand_batch(3, m, n, dtype=dt, seed=n + 2)
    Am, Bm = U.as_mats(A0, n, n).astype(np.float64), U.as_mats(B0, m, n).astype(np.float64)
    X = np.transpose(np.linalg.solve(Am, np.transpose(Bm, (0, 2, 1))), (0, 2, 1))  # X A = B, A symmetric
    # posv = potrf + potrs
    A, B = A0.copy(), B0.copy()
    assert U.oracle_posv("R", "L", m, n, A, B) == 1
    assert np.abs(U.as_mats(B, m, n) - X).max() <= _tol(dt, n, 100) * max(1.0, np.abs(X).max())
    # potrs given the factor
    A2, B2 = A0.copy(), B0.copy()
    U.oracle_potrf(A2, n)
    assert np.array_equal(A2, A)
    assert U.oracle_potrs("R", "L", m, n, A2, B2) == 1
    assert np.array_equal(B2, B)


def test_unsupported_variants_return_not_implemented():
    A = U.rand_spd_batch(1, 8)
    B = U.rand_batch(1, 8, 8)
    assert U.oracle_potrf(A.copy(), 8, uplo="U") == -2          # Xpotrf_batch_drivers.cuh:38-41
    assert U.oracle_trsm("L", "U", "N", "N", 8, 8, 1.0, A, B.copy()) == -2   # Xtrsm_batch_drivers.cuh:64-67
    assert U.oracle_trsm("L", "L", "N", "U", 8, 8, 1.0, A, B.copy()) == -2
    assert U.oracle_potrs("L", "L", 8, 8, A, B.copy()) == -2     # Xpotrs_batch_drivers.cuh:40-43
    assert U.oracle_posv("L", "L", 8, 8, A.copy(), B.copy()) == -2


@pytest.mark.parametrize("dt", DTYPES)
def test_non_spd_propagates_nan_like_the_reference(dt):
    # sqrt(negative) = NaN from that column on (Xpotrf_batch_kernels.cuh:121-129); earlier columns fine
    n = 16
    A = U.rand_spd_batch(3, n, dtype=dt, seed=5)
    A[1, 5, 5] = -3.0
    good = A.copy()
    U.oracle_potrf(A, n)
    U.oracle_potrf(good[[0, 2]], n)
    M = U.as_mats(A, n, n)
    assert np.isfinite(np.tril(M[0])).all() and np.isfinite(np.tril(M[2])).all()
    assert np.isfinite(np.tril(M[1])[:, :5]).all()
    assert np.isnan(M[1][5, 5]) and np.isnan(np.tril(M[1])[5:, 5]).all()


# ---- pinning against the reference itself --------------------------------------------------------
GOLD = os.path.join(U.GOLDEN_DIR, "reference_gpu.npz")


def _gold():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/reference_gpu.npz not generated yet (tests/golden/make_golden.py)")
    z = np.load(GOLD)
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def _n_of(name, tag="n"):
    for part in name.split("_"):
        if part.startswith(tag) and part[len(tag):].isdigit():
            return int(part[len(tag):])
    raise KeyError(name)


def test_golden_potrf_bit_exact_up_to_32_and_close_above():
    """n <= 32: no cuBLAS in the reference path -> the restatement must reproduce every bit.
    n > 32: the trailing updates go through cuBLAS GEMM -> BASELINE.json's tolerances."""
    cases = _gold()
    seen = 0
    for name, c in cases.items():
        if not name.startswith("potrf_") or "A_in" not in c or "nonspd" in name:
            continue
        dt = np.float64 if name.split("_")[1] == "D" else np.float32
        n = _n_of(name)
        A = c["A_in"].copy()
        assert U.oracle_potrf(A, n) == int(c["rc"]) == 1
        ref = c["A_out"]
        if n <= 32:
            assert np.array_equal(A, ref), name
        else:
            normA = np.abs(c["A_in"]).max()
            assert np.abs(np.tril(U.as_mats(A, n, n)) - np.tril(U.as_mats(ref, n, n))).max() <= 100 * n * U.EPS[dt] * normA, name
            assert np.array_equal(np.triu(U.as_mats(A, n, n), 1), np.triu(U.as_mats(ref, n, n), 1))
        assert (c["info"] == 77).all()  # the reference never writes info
        seen += 1
    assert seen >= 20


def test_golden_nonspd_finite_masks_and_return_codes():
    cases = _gold()
    for p, dt in (("D", np.float64), ("S", np.float32)):
        c = cases[f"potrf_{p}_nonspd_n16"]
        A = c["A_in"].copy()
        assert U.oracle_potrf(A, 16) == int(c["rc"])
        assert np.array_equal(np.isfinite(np.tril(U.as_mats(A, 16, 16))), np.isfinite(np.tril(U.as_mats(c["A_out"], 16, 16))))
        assert (c["info"] == 77).all()
        assert int(cases[f"potrf_{p}_upper"]["rc"]) == -2
        assert int(cases[f"posv_{p}_left"]["rc"]) == -2


def test_golden_trsm_potrs_posv():
    cases = _gold()
    seen = 0
    for name, c in cases.items():
        kind = name.split("_")[0]
        if kind not in ("trsm", "potrs", "posv") or "B_in" not in c:
            continue
        dt = np.float64 if name.split("_")[1] == "D" else np.float32
        m, n = _n_of(name, "m"), _n_of(name, "n")
        B = c["B_in"].copy()
        if kind == "trsm":
            side, trans = name.split("_")[2]
            k = m if side == "L" else n
            rc = U.oracle_trsm(side, "L", trans, "N", m, n, float(c["alpha"]), c["A_in"], B)
        elif kind == "potrs":
            k = n
            if n == 1:
                assert int(c["rc"]) == -2  # reference: n1 = 0 -> TRSM says NotImplemented
                continue
            rc = U.oracle_potrs("R", "L", m, n, c["L_in"], B)
        else:
            k = n
            A = c["A_in"].copy()
            rc = U.oracle_posv("R", "L", m, n, A, B)
        assert rc == int(c["rc"]) == 1, name
        ref = c["B_out"]
        scale = max(1.0, np.abs(ref[:, :, :m]).max())
        if k <= 16 and kind == "trsm":
            assert np.array_equal(B, ref), name       # pure register kernels: bit-exact
        else:
            assert np.abs(B - ref).max() <= 100 * k * U.EPS[dt] * scale, name
        seen += 1
    assert seen >= 60


# ---- round-2 goldens: config-4 sizes, pointer-array entry points, alpha == 0 --------------------------------
def _num(name, tag):
    for part in name.split("_"):
        if part.startswith(tag) and part[len(tag):].isdigit():
            return int(part[len(tag):])
    raise KeyError(name)


def test_golden_r2_oracle_matches_the_reference_library():
    """oracle vs reference_gpu_r2.npz (unmodified reference library on a B200): potrf n = 128 / 256, posv n = 256, the
    POINTER-ARRAY entry points (same arithmetic as the strided ones: bit-exact where the reference has no cuBLAS call),
    and trsm with alpha == 0 (the reference returns zeros except side R / trans T above k = 16, where its recursion
    multiplies by -1/alpha, Xtrsm_batch_drivers.cuh:154-163 -> non-finite; documented deviation: we return zeros)."""
    cases = U.load_golden("reference_gpu_r2.npz")
    if cases is None:
        pytest.skip("tests/golden/reference_gpu_r2.npz not generated yet (tests/golden/make_golden_r2.py)")
    seen = 0
    for name in cases:
        kind, p = name.split("_")[0], name.split("_")[1]
        dt = np.float64 if p == "D" else np.float32
        eps = U.EPS[dt]
        c = U.golden_r2_inputs(cases, name)
        if kind == "potrfbig":
            n = _num(name, "n")
            A = c["A_in"].copy()
            assert U.oracle_potrf(A, n) == int(c["rc"]) == 1
            assert np.abs(U.pack_lower(A, n) - c["L_out_packed"]).max() <= 100 * n * eps * np.abs(c["A_in"]).max(), name
            assert (c["info"] == 77).all()
        elif kind == "posvbig":
            m, n = _num(name, "m"), _num(name, "n")
            A, B = c["A_in"].copy(), c["B_in"].copy()
            assert U.oracle_posv("R", "L", m, n, A, B) == int(c["rc"]) == 1
            assert np.abs(U.pack_lower(A, n) - c["L_out_packed"]).max() <= 100 * n * eps * np.abs(c["A_in"]).max(), name
            assert np.abs(B - c["B_out"]).max() <= 100 * n * eps * max(1.0, np.abs(c["B_out"]).max()), name
        elif kind == "potrfptr":
            n = _num(name, "n")
            A = c["A_in"].copy()
            assert U.oracle_potrf(A, n) == int(c["rc"]) == 1
            if n <= 32:
                assert np.array_equal(A, c["A_out"]), name
            else:
                assert np.abs(np.tril(U.as_mats(A, n, n)) - np.tril(U.as_mats(c["A_out"], n, n))).max() <= 100 * n * eps * np.abs(c["A_in"]).max()
        elif kind in ("potrsptr", "posvptr", "trsmptr"):
            m, n = _num(name, "m"), _num(name, "n")
            B = c["B_in"].copy()
            if kind == "trsmptr":
                side, trans = name.split("_")[2]
                k = m if side == "L" else n
                rc = U.oracle_trsm(side, "L", trans, "N", m, n, float(c["alpha"]), c["L_in"], B)
            elif kind == "potrsptr":
                k = n
                rc = U.oracle_potrs("R", "L", m, n, c["L_in"], B)
            else:
                k = n
                rc = U.oracle_posv("R", "L", m, n, c["A_in"].copy(), B)
            assert rc == int(c["rc"]) == 1, name
            assert np.abs(B - c["B_out"]).max() <= 100 * k * eps * max(1.0, np.abs(c["B_out"]).max()), name
        elif kind == "trsmalpha0":
            side, trans = name.split("_")[2]
            k = _num(name, "m")
            finite = bool(np.isfinite(c["B_out"]).all())
            assert finite == (not (side == "R" and trans == "T" and k > 16)), name
            if finite:
                assert (c["B_out"] == 0).all(), name
        else:
            raise AssertionError(name)
        seen += 1
    assert seen >= 60, seen
