#!/usr/bin/env python
"""Round-2 golden vectors from the UNMODIFIED reference GPU library (oracle/_ref/libkblas_ref.so): the sizes and the
API forms reference_gpu.npz does not hold -- n = 128 / 256 (potrf, posv: BASELINE config 4 sizes), the POINTER-ARRAY
entry points (potrf / trsm / potrs / posv with shuffled device pointer arrays), and trsm with alpha == 0 (SURVEY
Appendix A: the reference's recursion divides by alpha for side R / trans T above k = 16).

    gpurun -- 'python tests/golden/make_golden_r2.py gpurun_out/golden'     -> reference_gpu_r2.npz
Same storage conventions as make_golden.py; pointer-array cases also store the permutation that built the arrays.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests._util import RefLib, as_mats, pack_lower, rand_batch, rand_spd_batch, sha  # noqa: E402

INFO_SENTINEL = 77


def main(outdir):
    import torch

    os.makedirs(outdir, exist_ok=True)
    ref = RefLib()
    H, i, l, c, P = ref.H, ref.i, ref.l, ref.c, ref.P
    dev = torch.device("cuda:0")
    cases = {}

    def dev_of(a):
        return torch.from_numpy(a).to(dev)

    def ptrs(d, perm, elems, es):
        return torch.from_numpy(d.data_ptr() + perm.astype(np.int64) * elems * es).to(dev)

    for dt, p, ct in ((np.float64, "D", C.c_double), (np.float32, "S", C.c_float)):
        es = np.dtype(dt).itemsize
        potrf_s = ref.fn(f"kblas{p}potrf_batch_strided", [H, c, i, P, i, l, i, P])
        posv_s = ref.fn(f"kblas{p}posv_batch_strided", [H, c, c, i, i, P, i, l, P, i, l, i, P])
        trsm_s = ref.fn(f"kblas{p}trsm_batch_strided", [H, c, c, c, c, i, i, ct, P, i, l, P, i, l, i])
        potrf_p = ref.fn(f"kblas{p}potrf_batch", [H, c, i, P, i, i, P])
        trsm_p = ref.fn(f"kblas{p}trsm_batch", [H, c, c, c, c, i, i, ct, P, i, P, i, i])
        potrs_p = ref.fn(f"kblas{p}potrs_batch", [H, c, c, i, i, P, i, P, i, i])
        posv_p = ref.fn(f"kblas{p}posv_batch", [H, c, c, i, i, P, i, P, i, i, P])

        # ---- strided potrf / posv at the config-4 sizes ----------------------------------------
        # (large: the input is stored as its generator seed + a SHA-256 of the bytes, the factor as its packed
        #  lower triangle -- tests regenerate the input with tests/_util.rand_spd_batch and verify the hash)
        for n in (128, 256):
            batch = 1
            A = rand_spd_batch(batch, n, dtype=dt, seed=900 + n)
            dA = dev_of(A)
            info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
            ref.wsquery("kblas_potrf_batch_strided_wsquery", "ii", n, batch)
            ref.allocate()
            rc = potrf_s(ref.h, b"L", n, dA.data_ptr(), n, n * n, batch, info.data_ptr())
            torch.cuda.synchronize()
            out = dA.cpu().numpy()
            assert np.array_equal(np.triu(as_mats(out, n, n), 1), np.triu(as_mats(A, n, n), 1))
            cases[f"potrfbig_{p}_n{n}"] = dict(seed=900 + n, A_in_sha256=sha(A), L_out_packed=pack_lower(out, n), rc=rc,
                                               info=info.cpu().numpy())
        m, n, batch = 16, 256, 1
        A = rand_spd_batch(batch, n, dtype=dt, seed=950)
        B = rand_batch(batch, m, n, dtype=dt, seed=951)
        dA, dB = dev_of(A), dev_of(B)
        info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
        ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
        ref.allocate()
        rc = posv_s(ref.h, b"R", b"L", m, n, dA.data_ptr(), n, n * n, dB.data_ptr(), m, m * n, batch, info.data_ptr())
        torch.cuda.synchronize()
        cases[f"posvbig_{p}_m{m}_n{n}"] = dict(seed_A=950, seed_B=951, A_in_sha256=sha(A), B_in=B,
                                               L_out_packed=pack_lower(dA.cpu().numpy(), n), B_out=dB.cpu().numpy(), rc=rc,
                                               info=info.cpu().numpy())

        # ---- pointer-array entry points -----------------------------------------------------------
        for n in (8, 16, 32, 64):
            batch, m = 3, 16
            perm = np.random.default_rng(n).permutation(batch)
            A = rand_spd_batch(batch, n, dtype=dt, seed=1000 + n)
            B = rand_batch(batch, m, n, dtype=dt, seed=1100 + n)
            # potrf
            dA = dev_of(A)
            info = torch.full((batch,), INFO_SENTINEL, dtype=torch.int32, device=dev)
            pa = ptrs(dA, perm, n * n, es)
            ref.wsquery("kblas_posv_batch_wsquery", "ciii", b"R", m, n, batch)
            ref.wsquery("kblas_trsm_batch_wsquery", "ciii", b"L", n, m, batch)
            ref.allocate()
            rc = potrf_p(ref.h, b"L", n, pa.data_ptr(), n, batch, info.data_ptr())
            torch.cuda.synchronize()
            L = dA.cpu().numpy()
            key = f"potrfptr_{p}_n{n}"
            cases[key] = dict(A_in=A, A_out=L, rc=rc, info=info.cpu().numpy(), perm=perm)
            # potrs with that factor (stored once: L_from names the case whose A_out is the factor)
            dB = dev_of(B)
            pb = ptrs(dB, perm, m * n, es)
            rc = potrs_p(ref.h, b"R", b"L", m, n, pa.data_ptr(), n, pb.data_ptr(), m, batch)
            torch.cuda.synchronize()
            cases[f"potrsptr_{p}_m{m}_n{n}"] = dict(L_from=key, B_in=B, B_out=dB.cpu().numpy(), rc=rc, perm=perm)
            # posv from A (A_from names the case whose A_in is the input; its factor equals that case's A_out)
            dA2, dB2 = dev_of(A), dev_of(B)
            pa2, pb2 = ptrs(dA2, perm, n * n, es), ptrs(dB2, perm, m * n, es)
            rc = posv_p(ref.h, b"R", b"L", m, n, pa2.data_ptr(), n, pb2.data_ptr(), m, batch, info.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(dA2.cpu().numpy(), L), "posv's factor differs from potrf's"
            cases[f"posvptr_{p}_m{m}_n{n}"] = dict(A_from=key, B_in=B, B_out=dB2.cpu().numpy(), rc=rc,
                                                   info=info.cpu().numpy(), perm=perm)
            # trsm, four variants, on the factor
            for side in "LR":
                for trans in "NT":
                    mm, nn = (n, m) if side == "L" else (m, n)
                    Bt = rand_batch(batch, mm, nn, dtype=dt, seed=1200 + n)
                    dBt = dev_of(Bt)
                    pbt = ptrs(dBt, perm, mm * nn, es)
                    rc = trsm_p(ref.h, side.encode(), b"L", trans.encode(), b"N", mm, nn, 0.28, pa.data_ptr(), n,
                                pbt.data_ptr(), mm, batch)
                    torch.cuda.synchronize()
                    cases[f"trsmptr_{p}_{side}{trans}_m{mm}_n{nn}"] = dict(L_from=key, B_in=Bt, B_out=dBt.cpu().numpy(), rc=rc,
                                                                          alpha=0.28, perm=perm)

        # ---- alpha == 0 (documents what the reference does; strided) ---------------------------------
        for side in "LR":
            for trans in "NT":
                for k in (16, 32):
                    batch = 2
                    A = rand_spd_batch(batch, k, dtype=dt, seed=1300 + k)
                    Bt = rand_batch(batch, k, k, dtype=dt, seed=1400 + k)
                    dA, dBt = dev_of(A), dev_of(Bt)
                    ref.wsquery("kblas_trsm_batch_strided_wsquery", "ciii", side.encode(), k, k, batch)
                    ref.allocate()
                    rc = trsm_s(ref.h, side.encode(), b"L", trans.encode(), b"N", k, k, 0.0, dA.data_ptr(), k, k * k,
                                dBt.data_ptr(), k, k * k, batch)
                    torch.cuda.synchronize()
                    cases[f"trsmalpha0_{p}_{side}{trans}_m{k}_n{k}"] = dict(A_in=A, B_in=Bt, B_out=dBt.cpu().numpy(), rc=rc,
                                                                           alpha=0.0)

    flat = {}
    for name, d in cases.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = np.asarray(v)
    path = os.path.join(outdir, "reference_gpu_r2.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")
    for name, d in sorted(cases.items()):
        extra = ""
        if name.startswith("trsmalpha0"):
            out = d["B_out"]
            extra = f" finite={bool(np.isfinite(out).all())} zeros={bool((out == 0).all())}"
        print(f"  {name}: rc={int(d['rc'])}{extra}")
    ref.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
