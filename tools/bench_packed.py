#!/usr/bin/env python
"""Time kblasx?pptrf_batch_strided (packed lower storage) per data-movement variant, next to kblas?potrf_batch_strided
on full storage; CUDA events, inputs larger than L2 restored from a pristine copy before every run.  GPU only.
usage: python tools/bench_packed.py [variants, e.g. 20,21,22,23,-1] [sizes, e.g. 32,24,16,8]"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

PEAK = 6554.6
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, restore, reps=7):
    ts = []
    for _ in range(reps):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    kb = importlib.import_module("kblas-gpu_b200")
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["20", "21", "22", "23", "-1"])]
    sizes = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["32", "24", "16", "8"])]
    batch = 1 << 20
    for prec, dt, es in (("D", torch.float64, 8), ("S", torch.float32, 4)):
        for n in sizes:
            sz = n * (n + 1) // 2
            A = bench.make_spd(torch, batch, n, dt, 1)
            P0 = torch.empty((batch, sz), device="cuda", dtype=dt)
            h0 = kb.Handle()
            assert h0.tri_pack_batch_strided("L", n, A, n, n * n, P0, sz, batch) == 1
            P = torch.empty_like(P0)
            algo = n * (n + 1) * es
            W = torch.empty_like(A)
            best, med = timeit(lambda: h0.potrf_batch_strided("L", n, W, n, n * n, batch, None), lambda: W.copy_(A))
            print(json.dumps({"op": f"{prec}potrf(full storage)", "n": n, "kernel": h0.last_kernel, "ms_best": round(best, 4),
                              "ms_median": round(med, 4), "frac": round(batch * algo / best / 1e6 / PEAK, 4)}), flush=True)
            # pack / unpack cost (utility kernels)
            best, med = timeit(lambda: h0.tri_pack_batch_strided("L", n, A, n, n * n, P, sz, batch), lambda: None, reps=3)
            print(json.dumps({"op": f"{prec}tri_pack", "n": n, "ms_best": round(best, 4)}), flush=True)
            h0.destroy()
            del W
            for v in variants:
                os.environ["KBLAS_B200_VARIANT"] = str(v)
                h = kb.Handle()
                best, med = timeit(lambda: h.pptrf_batch_strided("L", n, P, sz, batch, None), lambda: P.copy_(P0))
                print(json.dumps({"op": f"{prec}pptrf(packed)", "n": n, "variant": v, "kernel": h.last_kernel, "ms_best": round(best, 4),
                                  "ms_median": round(med, 4), "Mmat_s": round(batch / best / 1e3, 1),
                                  "algo_GBs": round(batch * algo / best / 1e6, 1), "frac": round(batch * algo / best / 1e6 / PEAK, 4),
                                  "frac_median": round(batch * algo / med / 1e6 / PEAK, 4)}), flush=True)
                h.destroy()
            os.environ.pop("KBLAS_B200_VARIANT", None)
            del A, P0, P


if __name__ == "__main__":
    main()
