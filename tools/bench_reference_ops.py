#!/usr/bin/env python
"""Time the UNMODIFIED reference library (oracle/_ref/libkblas_ref.so) on the BASELINE configurations
with the same buffers / events as tools/bench_variants.py.  Developer tool, GPU only."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tests._util import RefLib  # noqa: E402
from tools.bench_variants import timeit, PEAK  # noqa: E402


def main():
    ref = RefLib()
    H, i, l, c, P = ref.H, ref.i, ref.l, ref.c, ref.P
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for prec, dt, es, ct in (("D", torch.float64, 8, C.c_double), ("S", torch.float32, 4, C.c_float)):
        potrf = ref.fn(f"kblas{prec}potrf_batch_strided", [H, c, i, P, i, l, i, P])
        potrs = ref.fn(f"kblas{prec}potrs_batch_strided", [H, c, c, i, i, P, i, l, P, i, l, i])
        trsm = ref.fn(f"kblas{prec}trsm_batch_strided", [H, c, c, c, c, i, i, ct, P, i, l, P, i, l, i])
        posvp = ref.fn(f"kblas{prec}posv_batch", [H, c, c, i, i, P, i, P, i, i, P])
        potrfp = ref.fn(f"kblas{prec}potrf_batch", [H, c, i, P, i, i, P])
        if which in ("all", "small"):
            batch = 1 << 20
            for n in (32, 24, 16, 8):
                m = n
                Pm = bench.make_spd(torch, batch, n, dt, 1)
                A = torch.empty_like(Pm)
                ref.wsquery("kblas_posv_batch_strided_wsquery", "ciii", b"R", m, n, batch)
                ref.wsquery("kblas_trsm_batch_strided_wsquery", "ciii", b"L", m, n, batch)
                ref.allocate()
                best, _ = timeit(lambda: potrf(ref.h, b"L", n, A.data_ptr(), n, n * n, batch, None), lambda: A.copy_(Pm))
                algo = n * (n + 1) * es
                print(json.dumps({"impl": "reference", "op": f"{prec}potrf", "n": n, "ms_best": best, "Mmat_s": batch / best / 1e3,
                                  "frac": batch * algo / best / 1e6 / PEAK}), flush=True)
                if n in (32, 24, 16, 8):
                    L = Pm.clone()
                    potrf(ref.h, b"L", n, L.data_ptr(), n, n * n, batch, None)
                    B0 = torch.rand((batch, n, m), device="cuda", dtype=dt)
                    B = torch.empty_like(B0)
                    algo_t = (n * (n + 1) // 2 + 2 * m * n) * es
                    ops = [("potrs_R", lambda: potrs(ref.h, b"R", b"L", m, n, L.data_ptr(), n, n * n, B.data_ptr(), m, m * n, batch))]
                    for s_, t_ in (("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")):
                        ops.append((f"trsm_{s_}L{t_}", (lambda s_=s_, t_=t_: trsm(ref.h, s_.encode(), b"L", t_.encode(), b"N", m, n, 0.28,
                                                                                 L.data_ptr(), n, n * n, B.data_ptr(), m, m * n, batch))))
                    for name, fn in ops:
                        best, _ = timeit(fn, lambda: B.copy_(B0))
                        print(json.dumps({"impl": "reference", "op": f"{prec}{name}", "n": n, "ms_best": best, "Mprob_s": batch / best / 1e3,
                                          "frac": batch * algo_t / best / 1e6 / PEAK}), flush=True)
                    del L, B0, B
                del Pm, A
        if which in ("all", "large"):
            m = 16
            for n in (64, 128, 256):
                b = (1 << 16) if n < 256 or es == 4 else (1 << 15)
                Pm = bench.make_spd(torch, b, n, dt, 1)
                A = torch.empty_like(Pm)
                B0 = torch.rand((b, n, m), device="cuda", dtype=dt)
                B = torch.empty_like(B0)
                perm = torch.randperm(b, device="cuda")
                pa = (A.data_ptr() + perm * (n * n * es)).contiguous()
                pb = (B.data_ptr() + perm * (m * n * es)).contiguous()
                ref.wsquery("kblas_posv_batch_wsquery", "ciii", b"R", m, n, b)
                ref.allocate()

                def restore():
                    A.copy_(Pm)
                    B.copy_(B0)
                flops = n ** 3 / 3 + n ** 2 / 2 + n / 6 + 2 * m * n * n
                for name, fn in (("posv_ptr", lambda: posvp(ref.h, b"R", b"L", m, n, pa.data_ptr(), n, pb.data_ptr(), m, b, None)),
                                 ("potrf_ptr", lambda: potrfp(ref.h, b"L", n, pa.data_ptr(), n, b, None))):
                    best, _ = timeit(fn, restore, reps=3)
                    fl = flops if name == "posv_ptr" else n ** 3 / 3 + n ** 2 / 2 + n / 6
                    print(json.dumps({"impl": "reference", "op": f"{prec}{name}", "n": n, "m": m, "batch": b, "ms_best": best,
                                      "Mprob_s": b / best / 1e3, "TFLOPs": b * fl / best / 1e9}), flush=True)
                del Pm, A, B0, B
    ref.close()


if __name__ == "__main__":
    main()
