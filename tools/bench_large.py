#!/usr/bin/env python
"""BASELINE config 4 timings: pointer-array dpotrf / dposv, n = 64 / 128 / 256, 16 right-hand-side rows, batch 64K, per
kernel variant (KBLAS_B200_VARIANT: 30 = one warp per matrix, operands from global/L2; -1 = shared-memory resident factor
with the default warps per matrix; 31 / 32 / 33 = 2 / 4 / 8 warps).  CUDA events, best / median of 5.  GPU only."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

PEAK = 6554.6


def timeit(fn, restore, reps=5):
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
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["30", "-1"])]
    sizes = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["64", "128", "256"])]
    ops = sys.argv[3].split(",") if len(sys.argv) > 3 else ["potrf_ptr", "posv_ptr"]
    m, b = 16, 1 << 16
    fp64 = bench.measure_fp64_peak(torch)
    print(json.dumps({"fp64_peak_tflops_matmul": fp64}), flush=True)
    for n in sizes:
        P = bench.make_spd(torch, b, n, torch.float64, 1)
        A = torch.empty_like(P)
        B0 = torch.rand((b, n, m), device="cuda", dtype=torch.float64)
        B = torch.empty_like(B0)
        perm = torch.randperm(b, device="cuda")
        pa = (A.data_ptr() + perm * (n * n * 8)).contiguous()
        pb = (B.data_ptr() + perm * (m * n * 8)).contiguous()

        def restore():
            A.copy_(P)
            B.copy_(B0)
        pf = n ** 3 / 3 + n ** 2 / 2 + n / 6
        for v in variants:
            os.environ["KBLAS_B200_VARIANT"] = str(v)
            h = kb.Handle()
            h.posv_batch_wsquery("R", m, n, b)
            h.allocate_workspace()
            for name in ops:
                if name == "posv_ptr":
                    fn, fl, by = (lambda: h.posv_batch("R", "L", m, n, pa, n, pb, m, b, None, prec="D")), pf + 2 * m * n * n, (n * (n + 1) + 2 * m * n) * 8
                else:
                    fn, fl, by = (lambda: h.potrf_batch("L", n, pa, n, b, None, prec="D")), pf, n * (n + 1) * 8
                best, med = timeit(fn, restore)
                print(json.dumps({"op": "D" + name, "n": n, "batch": b, "variant": v, "kernel": h.last_kernel, "ms_best": round(best, 3),
                                  "ms_median": round(med, 3), "TFLOPs": round(b * fl / best / 1e9, 2), "frac_fp64": round(b * fl / best / 1e9 / fp64, 3),
                                  "frac_hbm": round(b * by / best / 1e6 / PEAK, 3)}), flush=True)
            h.destroy()
        os.environ.pop("KBLAS_B200_VARIANT", None)
        del P, A, B0, B


if __name__ == "__main__":
    main()
