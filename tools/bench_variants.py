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
    else:
        for prec, dt, es in (("D", torch.float64, 8), ("S", torch.float32, 4)):
            for n in (32, 16, 8):
                m = n
                P = bench.make_spd(torch, batch, n, dt, 1)
                L = P.clone()
                h = kb.Handle()
                h.potrf_batch_strided("L", n, L, n, n * n, batch, None)
                B0 = torch.rand((batch, n, m), device="cuda", dtype=dt)
                B = torch.empty_like(B0)
                algo_trs = (n * (n + 1) // 2 + 2 * m * n) * es
                for name, fn in (
                    ("potrs_R", lambda: h.potrs_batch_strided("R", "L", m, n, L, n, n * n, B, m, m * n, batch)),
                    ("trsm_LLN", lambda: h.trsm_batch_strided("L", "L", "N", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                    ("trsm_LLT", lambda: h.trsm_batch_strided("L", "L", "T", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                    ("trsm_RLN", lambda: h.trsm_batch_strided("R", "L", "N", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                    ("trsm_RLT", lambda: h.trsm_batch_strided("R", "L", "T", "N", m, n, 0.28, L, n, n * n, B, m, m * n, batch)),
                ):
                    best, mean = timeit(fn, lambda: B.copy_(B0))
                    print(json.dumps({"op": f"{prec}{name}", "n": n, "m": m, "kernel": h.last_kernel, "ms_best": best,
                                      "ms_mean": mean, "Mprob_s": batch / best / 1e3, "algo_GBs": batch * algo_trs / best / 1e6,
                                      "frac": batch * algo_trs / best / 1e6 / PEAK}), flush=True)
                h.destroy()
                del P, L, B0, B


if __name__ == "__main__":
    main()
