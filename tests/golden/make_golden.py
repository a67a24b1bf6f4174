#!/usr/bin/env python
"""Generate golden input/output vectors from the UNMODIFIED reference GPU library.

Runs on a GPU box only (the reference path is CUDA):
    make -C oracle ref          # here, where /root/reference exists -> oracle/_ref/libkblas_ref.so
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.

Each case stores the exact inputs (so the fixtures do not depend on any RNG's stream staying
stable), the reference's outputs, its return code and the info array it was handed (pre-set to
the sentinel 77: the reference never writes it, SURVEY.md §0 finding 1).
tests/test_oracle.py pins oracle/kblas_oracle.c against these files; tests/test_gpu_parity.py
compares the CUDA kernels with them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests._util import RefLib, rand_batch, rand_spd_batch  # noqa: E402

INFO_SENTINEL = 77


def main(outdir):
    import ctypes as C

    import torch

    os.makedirs(outdir, exist_ok=True)
    ref = RefLib()
    H, i, l, c, P = ref.H, ref.i, ref.l, ref.c, ref.P
    dev = torch.device("cuda:0")
    cases = {}

    def dev_of(a):
        return torch.from_numpy(a).to(dev)

    for dt, p, ct in ((np.float64, "D", C.c_double), (np.float32, "S", C.c_float)):
        potrf = ref.fn(f"kblas{p}potrf_batch_strided", [H, c, i, P, i, l, i, P])
        trsm = ref.fn(f"kblas{p}trsm_batch_strided", [H, c, c, c, c, i, i, ct, P, i, l, P, i, l, i])
        potrs = ref.fn(f"kblas{p}potrs_batch_strided", [H, c, c, i, i, P, i, l, P, i, l, i])
        posv = ref.fn(f"kblas{p}posv_batch_strided", [H, c, c, i, i, P, i, l, P, i, l, i, P])

        # ---- potrf -------------------------------------------------------------------------
        for n, pad in ((1, 0), (2, 0), (5, 3), (8, 0), (11, 0), (16, 0), (17, 2), (24, 0), (31, 1), (32, 0),
                       (40, 0), (64, 0), (100, 0)):
            batch, lda = 5, n + pad
            A = rand_spd_batch(batch, n, lda=lda, dtype=dt, seed=100 + n)
            dA = dev_of(A)
            info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
            ref.wsquery("kblas_potrf_batch_strided_wsquery", "ii", n, batch)
            ref.allocate()
            rc = potrf(ref.h, b"L", n, dA.data_ptr(), lda, n * lda, batch, info.data_ptr())
            torch.cuda.synchronize()
            cases[f"potrf_{p}_n{n}_lda{lda}"] = dict(A_in=A, A_out=dA.cpu().numpy(), rc=rc, info=info.cpu().numpy())

        # ---- potrf on non-SPD input: NaN propagation, info untouched --------------------------
        n, batch = 16, 4
        A = rand_spd_batch(batch, n, dtype=dt, seed=7)
        A[1, 5, 5] = -3.0   # negative pivot region
        A[2, 9, 9] = 0.0
        dA = dev_of(A)
        info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
        rc = potrf(ref.h, b"L", n, dA.data_ptr(), n, n * n, batch, info.data_ptr())
        torch.cuda.synchronize()
        cases[f"potrf_{p}_nonspd_n{n}"] = dict(A_in=A, A_out=dA.cpu().numpy(), rc=rc, info=info.cpu().numpy())

        # ---- upper: not implemented -----------------------------------------------------------
        rc = potrf(ref.h, b"U", n, dA.data_ptr(), n, n * n, batch, info.data_ptr())
        cases[f"potrf_{p}_upper"] = dict(rc=rc)

        # ---- trsm -----------------------------------------------------------------------------
        for side in "LR":
            for trans in "NT":
                for m, n in ((8, 8), (16, 16), (13, 7), (7, 13), (32, 32), (20, 32), (32, 20), (48, 16), (16, 48)):
                    k = m if side == "L" else n
                    batch = 3
                    A = rand_spd_batch(batch, k, dtype=dt, seed=200 + k)
                    B = rand_batch(batch, m, n, dtype=dt, seed=300 + m * 64 + n)
                    dA, dB = dev_of(A), dev_of(B)
                    ref.wsquery("kblas_trsm_batch_strided_wsquery", "ciii", side.encode(), m, n, batch)
                    ref.allocate()
                    rc = trsm(ref.h, side.encode(), b"L", trans.encode(), b"N", m, n, 0.28, dA.data_ptr(), k, k * k,
                              dB.data_ptr(), m, m * n, batch)
                    torch.cuda.synchronize()
                    cases[f"trsm_{p}_{side}{trans}_m{m}_n{n}"] = dict(A_in=A, B_in=B, B_out=dB.cpu().numpy(), rc=rc,
                                                                     alpha=0.28)

        # ---- potrs (given the reference's own factor) -------------------------------------------
        for m, n in ((8, 8), (16, 16), (32, 32), (16, 24), (5, 13), (16, 64), (3, 1)):
            batch = 3
            A = rand_spd_batch(batch, n, dtype=dt, seed=400 + n)
            B = rand_batch(batch, m, n, dtype=dt, seed=500 + m * 64 + n)
            dA, dB = dev_of(A), dev_of(B)
            info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
            ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
            ref.allocate()
            rc0 = potrf(ref.h, b"L", n, dA.data_ptr(), n, n * n, batch, info.data_ptr())
            torch.cuda.synchronize()
            L = dA.cpu().numpy()
            rc = potrs(ref.h, b"R", b"L", m, n, dA.data_ptr(), n, n * n, dB.data_ptr(), m, m * n, batch)
            torch.cuda.synchronize()
            cases[f"potrs_{p}_m{m}_n{n}"] = dict(L_in=L, B_in=B, B_out=dB.cpu().numpy(), rc=rc, rc_potrf=rc0)

        # ---- posv -----------------------------------------------------------------------------------
        for m, n in ((16, 32), (8, 20), (16, 64), (16, 128)):
            batch = 2
            A = rand_spd_batch(batch, n, dtype=dt, seed=600 + n)
            B = rand_batch(batch, m, n, dtype=dt, seed=700 + m * 64 + n)
            dA, dB = dev_of(A), dev_of(B)
            info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
            ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
            ref.allocate()
            rc = posv(ref.h, b"R", b"L", m, n, dA.data_ptr(), n, n * n, dB.data_ptr(), m, m * n, batch,
                      info.data_ptr())
            torch.cuda.synchronize()
            cases[f"posv_{p}_m{m}_n{n}"] = dict(A_in=A, B_in=B, A_out=dA.cpu().numpy(), B_out=dB.cpu().numpy(), rc=rc,
                                                info=info.cpu().numpy())
        rc = posv(ref.h, b"L", b"L", m, n, dA.data_ptr(), n, n * n, dB.data_ptr(), m, m * n, batch, info.data_ptr())
        cases[f"posv_{p}_left"] = dict(rc=rc)

    flat = {}
    for name, d in cases.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = np.asarray(v)
    path = os.path.join(outdir, "reference_gpu.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")
    for name, d in sorted(cases.items()):
        print(f"  {name}: rc={int(d['rc'])}")
    ref.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
