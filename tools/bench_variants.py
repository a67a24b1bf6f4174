#!/usr/bin/env python
"""Time kernel variants (handle->variant_override via env KBLAS_B200_VARIANT) and the other BASELINE
configurations with CUDA events; prints one JSON line per measurement.  Developer tool, GPU only."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

PEAK = 6554.6


def timeit(fn, restore, reps=5):
    best, tot = 1e9, 0.0
    for _ in range(reps):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return best, tot / reps


def main():
    kb = importlib.import_module("kblas-gpu_b200")
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["-1"])]
    which = sys.argv[2] if len(sys.argv) > 2 else "potrf"
    batch = 1 << 20
    if which == "potrf":
        for prec, dt, es in (("D", torch.float64, 8), ("S", torch.float32, 4)):
            for n in (32, 24, 16, 8):
                P = bench.make_spd(torch, batch, n, dt, 1)
                A = torch.empty_like(P)
                for v in variants:
                    os.environ["KBLAS_B200_VARIANT"] = str(v)
                    h = kb.Handle()
                    best, mean = timeit(lambda: h.potrf_batch_strided("L", n, A, n, n * n, batch, None), lambda: A.copy_(P))
                    algo = n * (n + 1) * es
                    print(json.dumps({"op": f"{prec}potrf", "n": n, "variant": v, "kernel": h.last_kernel, "ms_best": best,
                                      "ms_mean": mean, "Mmat_s": batch / best / 1e3, "algo_GBs": batch * algo / best / 1e6,
                                      "frac": batch * algo / best / 1e6 / PEAK}), flush=True)
                    h.destroy()
                del P, A
    elif which == "large":
        # BASELINE config 4: pointer-array posv, n = 64/128/256, 16 right-hand-side rows, batch 64K
        m = 16
        for prec, dt, es in (("D", torch.float64, 8), ("S", torch.float32, 4)):
            for n in (64, 128, 256):
                b = (1 << 16) if n < 256 or es == 4 else (1 << 15)
                P = bench.make_spd(torch, b, n, dt, 1)
                A = torch.empty_like(P)
                B0 = torch.rand((b, n, m), device="cuda", dtype=dt)
                B = torch.empty_like(B0)
                perm = torch.randperm(b, device="cuda")
                pa = (A.data_ptr() + perm * (n * n * es)).contiguous()
                pb = (B.data_ptr() + perm * (m * n * es)).contiguous()
                def restore():
                    A.copy_(P)
                    B.copy_(B0)
                flops = n ** 3 / 3 + n ** 2 / 2 + n / 6 + 2 * m * n * n
                algo = (n * (n + 1) + 2 * m * n) * es
                for v in variants:
                    os.environ["KBLAS_B200_VARIANT"] = str(v)
                    h = kb.Handle()
                    h.posv_batch_wsquery("R", m, n, b)
                    h.allocate_workspace()
                    for name, fn in (("posv_ptr", lambda: h.posv_batch("R", "L", m, n, pa, n, pb, m, b, None, prec=prec)),
                                     ("potrf_ptr", lambda: h.potrf_batch("L", n, pa, n, b, None, prec=prec))):
                        best, mean = timeit(fn, restore, reps=3)
                        fl = flops if name == "posv_ptr" else n ** 3 / 3 + n ** 2 / 2 + n / 6
                        print(json.dumps({"op": f"{prec}{name}", "n": n, "m": m, "batch": b, "variant": v, "kernel": h.last_kernel,
                                          "ms_best": best, "Mprob_s": b / best / 1e3, "TFLOPs": b * fl / best / 1e9,
                                          "algo_GBs": b * algo / best / 1e6, "frac_hbm": b * algo / best / 1e6 / PEAK}), flush=True)
                    h.destroy()
                del P, A, B0, B
    else:
        ns = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32, 24, 16, 8]
        for prec, dt, es in (("D", torch.float64, 8), ("S", torch.float32, 4)):
            for n in ns:
                m = n
                P = bench.make_spd(torch, batch, n, dt, 1)
                L = P.clone()
                h0 = kb.Handle()
                h0.potrf_batch_strided("L", n, L, n, n * n, batch, None)
                h0.destroy()
                B0 = torch.rand((batch, n, m), device="cuda", dtype=dt)
                B = torch.empty_like(B0)
                algo_trs = (n * (n + 1) // 2 + 2 * m * n) * es
                for v in variants:
                    os.environ["KBLAS_B200_VARIANT"] = str(v)
                    h = kb.Handle()
                    for name, fn in (
                        ("potrs_R", lambda: h.potrs_batch_strided("R", "L", m, n, L, n, n * n, B, m, m * n, batch)),
                        ("potrs_L", lambda: h.potrs_batch_strided("L", "L", n, m, L, n, n * n, B, n, m * n, batch)),
                        ("trsm_LLN", lambda: h.trsm_batch_strided("L", "L", "N", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                        ("trsm_LLT", lambda: h.trsm_batch_strided("L", "L", "T", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                        ("trsm_RLN", lambda: h.trsm_batch_strided("R", "L", "N", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                        ("trsm_RLT", lambda: h.trsm_batch_strided("R", "L", "T", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                    ):
                        best, mean = timeit(fn, lambda: B.copy_(B0))
                        print(json.dumps({"op": f"{prec}{name}", "n": n, "m": m, "variant": v, "kernel": h.last_kernel, "ms_best": best,
                                          "ms_mean": mean, "Mprob_s": batch / best / 1e3, "algo_GBs": batch * algo_trs / best / 1e6,
                                          "frac": batch * algo_trs / best / 1e6 / PEAK}), flush=True)
                    h.destroy()
                del P, L, B0, B


if __name__ == "__main__":
    main()
