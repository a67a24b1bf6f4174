#!/usr/bin/env python
"""Randomised stress of the factorisation dispatch: potrf on random n (1..300), leading dimension, batch stride, base-pointer
alignment, uplo, precision, strided and pointer-array; every matrix against a float64 numpy Cholesky, the other triangle and
all padding must keep their bits.  GPU only; `python tools/stress_potrf.py [cases] [seed]`."""
import importlib, os, sys, collections
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
kb = importlib.import_module("kblas-gpu_b200")

def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
    h = kb.Handle()
    h.potrf_batch_wsquery(300, 64); h.potrf_batch_strided_wsquery(300, 64)
    assert h.allocate_workspace() == kb.KBLAS_Success
    seen = collections.Counter(); worst = 0.0
    for it in range(cases):
        p = "D" if rng.random() < 0.5 else "S"
        dt = np.float64 if p == "D" else np.float32
        eps = np.finfo(dt).eps; es = np.dtype(dt).itemsize
        n = int(rng.choice([rng.integers(1, 33), rng.choice([8, 16, 24, 32]), rng.integers(33, 130), rng.choice([64, 128, 256]), rng.integers(130, 300)], p=[0.3, 0.25, 0.25, 0.1, 0.1]))
        lda = n + int(rng.choice([0, 0, 1, 2, 3, 4]))
        batch = int(rng.integers(1, 24 if n <= 64 else 6))
        stride = n * lda + int(rng.choice([0, 0, 1, 2, 8]))
        off = int(rng.choice([0, 0, 1]))
        uplo = "U" if rng.random() < 0.25 else "L"
        G = rng.random((batch, n, n)) - 0.5
        S = G @ np.transpose(G, (0, 2, 1)) + n * np.eye(n)[None]
        S = S.astype(dt).astype(np.float64); S = (S + np.transpose(S, (0, 2, 1))) / 2
        S = S.astype(dt)
        buf = np.full(off + batch * stride + 8, -7.25, dtype=dt)
        for b in range(batch):
            blk = np.full((n, lda), -7.25, dtype=dt); blk[:, :n] = S[b].T
            buf[off + b * stride: off + b * stride + n * lda] = blk.flatten()
        d = torch.from_numpy(buf).cuda(); Av = d[off:]
        ptr = rng.random() < 0.3
        if ptr:
            perm = torch.randperm(batch, device="cuda")
            pa = (Av.data_ptr() + perm * (stride * es)).contiguous()
            rc = h.potrf_batch(uplo, n, pa, lda, batch, None, prec=p)
        else:
            rc = h.potrf_batch_strided(uplo, n, Av, lda, stride, batch, None)
        torch.cuda.synchronize()
        tag = (p, n, lda, stride, off, uplo, batch, "ptr" if ptr else "strided", h.last_kernel)
        assert rc == kb.KBLAS_Success, (rc, tag)
        got = d.cpu().numpy()
        for b in range(batch):
            blk = got[off + b * stride: off + b * stride + n * lda].reshape(n, lda)
            M = blk[:, :n].T            # [row, col]
            L = np.linalg.cholesky(S[b].astype(np.float64))
            if uplo == "L":
                F, other, orig = np.tril(M), np.triu(M, 1), np.triu(S[b], 1)
                ref = L
            else:
                F, other, orig = np.triu(M), np.tril(M, -1), np.tril(S[b], -1)
                ref = L.T
            assert np.array_equal(other, orig), ("other triangle modified", tag)
            assert (blk[:, n:] == -7.25).all(), ("lda padding", tag)
            assert (got[off + b * stride + n * lda: off + (b + 1) * stride] == -7.25).all(), ("gap", tag)
            err = np.abs(F.astype(np.float64) - ref).max() / np.abs(ref).max()
            assert np.isfinite(err) and err <= 100 * n * eps, (err, tag)
            worst = max(worst, err / (100 * n * eps))
        seen[h.last_kernel.split("<")[0]] += 1
    print("ok", cases, "cases; worst err/tol", round(worst, 4), dict(seen))

if __name__ == "__main__":
    main()
