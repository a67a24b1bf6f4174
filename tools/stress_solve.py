#!/usr/bin/env python
"""Randomised stress of every solve dispatch branch: trsm (all side / uplo / trans / diag), potrs (both sides, both uplo) and
posv on random k, number of right-hand sides, leading dimensions, batch strides, base-pointer alignment and precision, strided
and pointer-array, each checked against a float64 numpy solve.  GPU only; `python tools/stress_solve.py [cases] [seed]`."""
import importlib, os, sys, collections
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
kb = importlib.import_module("kblas-gpu_b200")

def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    h = kb.Handle()
    h.posv_batch_strided_wsquery("R", 300, 300, 64); h.posv_batch_wsquery("R", 300, 300, 64)
    h.posv_batch_strided_wsquery("L", 300, 300, 64); h.posv_batch_wsquery("L", 300, 300, 64)
    h.trsm_batch_strided_wsquery("L", 300, 300, 64); h.trsm_batch_wsquery("L", 300, 300, 64)
    h.trsm_batch_strided_wsquery("R", 300, 300, 64); h.trsm_batch_wsquery("R", 300, 300, 64)
    assert h.allocate_workspace() == kb.KBLAS_Success
    seen = collections.Counter()
    worst = 0.0
    for it in range(cases):
        p = "D" if rng.random() < 0.5 else "S"
        dt = np.float64 if p == "D" else np.float32
        tdt = torch.float64 if p == "D" else torch.float32
        eps = np.finfo(dt).eps
        es = np.dtype(dt).itemsize
        vw = 16 // es
        k = int(rng.choice([rng.integers(1, 33), rng.choice([8, 16, 24, 32]), rng.integers(33, 140), rng.choice([64, 128])], p=[0.35, 0.35, 0.2, 0.1]))
        vec = int(rng.choice([rng.integers(1, 70), rng.choice([8, 16, 32, 64]), k]))
        side = "L" if rng.random() < 0.5 else "R"
        m, n = (k, vec) if side == "L" else (vec, k)
        aligned = rng.random() < 0.6
        lda = k + (int(rng.integers(0, 3)) * vw if aligned else int(rng.integers(0, 4)))
        ldb = m + (int(rng.integers(0, 3)) * vw if aligned else int(rng.integers(0, 4)))
        if aligned:
            lda = (lda + vw - 1) // vw * vw
            ldb = (ldb + vw - 1) // vw * vw
        batch = int(rng.integers(1, 40))
        sa = k * lda + (int(rng.integers(0, 2)) * vw if aligned else int(rng.integers(0, 3)))
        sb = n * ldb + (int(rng.integers(0, 2)) * vw if aligned else int(rng.integers(0, 3)))
        offa = 0 if aligned else int(rng.integers(0, 2))
        offb = 0 if aligned else int(rng.integers(0, 2))
        op = rng.choice(["trsm", "potrs", "posv"], p=[0.6, 0.25, 0.15])
        uplo = "U" if rng.random() < 0.2 else "L"
        # well-conditioned triangular factor T (stored triangle only; NaN elsewhere for trsm)
        M = 0.2 * (2 * rng.random((batch, k, k)) - 1)
        M = np.triu(M, 1) if uplo == "U" else np.tril(M, -1)
        M = M + np.eye(k)[None] * (1 + rng.random((batch, k, 1)))
        M = M.astype(dt).astype(np.float64)
        Bm = rng.random((batch, m, n)).astype(dt).astype(np.float64)
        Abuf = np.full(offa + batch * sa + 8, np.nan, dtype=dt)
        Bbuf = np.full(offb + batch * sb + 8, -3.5, dtype=dt)
        if op == "posv":
            full = (np.transpose(M, (0, 2, 1)) @ M) if uplo == "U" else (M @ np.transpose(M, (0, 2, 1)))
            stored = full
        else:
            stored = M
        for b in range(batch):
            blk = np.full((k, lda), np.nan, dtype=dt)
            blk[:, :k] = stored[b].T.astype(dt)
            if op != "posv":
                keep = np.triu(np.ones((k, k), bool)) if uplo == "U" else np.tril(np.ones((k, k), bool))
                blk[:, :k] = np.where(keep.T, blk[:, :k], np.nan)
            Abuf[offa + b * sa: offa + b * sa + k * lda] = blk.flatten()
            bb = np.full((n, ldb), -3.5, dtype=dt)
            bb[:, :m] = Bm[b].T.astype(dt)
            Bbuf[offb + b * sb: offb + b * sb + n * ldb] = bb.flatten()
        if op == "posv":
            Mf = np.linalg.cholesky(stored)            # float64 factor of the stored (rounded) matrices
            Mf = np.transpose(Mf, (0, 2, 1)) if uplo == "U" else Mf
        else:
            Mf = M
        dA, dB = torch.from_numpy(Abuf).cuda(), torch.from_numpy(Bbuf).cuda()
        ptr = rng.random() < 0.3
        alpha = 0.28
        trans, diag = ("T" if rng.random() < 0.5 else "N"), ("U" if (op == "trsm" and rng.random() < 0.15) else "N")
        if diag == "U":
            for b in range(batch):
                idx = offa + b * sa + np.arange(k) * (lda + 1)
                Abuf[idx] = np.nan
            dA = torch.from_numpy(Abuf).cuda()
            Mf = Mf.copy(); Mf[:, np.arange(k), np.arange(k)] = 1.0
        Av, Bv = dA[offa:], dB[offb:]
        if ptr:
            pa = (Av.data_ptr() + torch.arange(batch, device="cuda") * (sa * es)).contiguous()
            pb = (Bv.data_ptr() + torch.arange(batch, device="cuda") * (sb * es)).contiguous()
        if op == "trsm":
            Op = Mf if trans == "N" else np.transpose(Mf, (0, 2, 1))
            want = np.linalg.solve(Op, alpha * Bm) if side == "L" else np.transpose(np.linalg.solve(np.transpose(Op, (0, 2, 1)), np.transpose(alpha * Bm, (0, 2, 1))), (0, 2, 1))
            rc = (h.trsm_batch(side, uplo, trans, diag, m, n, alpha, pa, lda, pb, ldb, batch, prec=p) if ptr else
                  h.trsm_batch_strided(side, uplo, trans, diag, m, n, alpha, Av, lda, sa, Bv, ldb, sb, batch))
        else:
            Afull = (np.transpose(Mf, (0, 2, 1)) @ Mf) if uplo == "U" else (Mf @ np.transpose(Mf, (0, 2, 1)))
            want = np.linalg.solve(Afull, Bm) if side == "L" else np.transpose(np.linalg.solve(np.transpose(Afull, (0, 2, 1)), np.transpose(Bm, (0, 2, 1))), (0, 2, 1))
            if op == "potrs":
                rc = (h.potrs_batch(side, uplo, m, n, pa, lda, pb, ldb, batch, prec=p) if ptr else
                      h.potrs_batch_strided(side, uplo, m, n, Av, lda, sa, Bv, ldb, sb, batch))
            else:
                rc = (h.posv_batch(side, uplo, m, n, pa, lda, pb, ldb, batch, None, prec=p) if ptr else
                      h.posv_batch_strided(side, uplo, m, n, Av, lda, sa, Bv, ldb, sb, batch, None))
        torch.cuda.synchronize()
        tag = (op, side, uplo, trans, diag, p, k, vec, lda, ldb, sa, sb, offa, offb, batch, "ptr" if ptr else "strided", h.last_kernel)
        assert rc == kb.KBLAS_Success, (rc, tag)
        got = dB.cpu().numpy()
        err = 0.0
        for b in range(batch):
            blk = got[offb + b * sb: offb + b * sb + n * ldb].reshape(n, ldb)
            X = blk[:, :m].T.astype(np.float64)
            assert np.isfinite(X).all(), tag
            err = max(err, np.abs(X - want[b]).max() / max(1.0, np.abs(want[b]).max()))
            assert (blk[:, m:] == -3.5).all(), ("padding", tag)
            assert (got[offb + b * sb + n * ldb: offb + (b + 1) * sb] == -3.5).all(), ("gap", tag)
        tol = (2000 if op != "trsm" else 100) * k * eps
        assert err <= tol, (err, tol, tag)
        worst = max(worst, err / tol)
        if op == "trsm":
            assert np.array_equal(dA.cpu().numpy(), Abuf, equal_nan=True), ("A modified", tag)
        seen[h.last_kernel.split("<")[0]] += 1
    print("ok", cases, "cases; worst err/tol", round(worst, 3), dict(seen))

if __name__ == "__main__":
    main()
