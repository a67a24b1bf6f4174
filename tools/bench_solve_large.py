#!/usr/bin/env python
"""time the k > 32 solves (side R and L, strided) for a few shapes: default dispatch vs the FMA kernels (variant 44).  GPU only."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools.bench_variants import timeit
kb = importlib.import_module("kblas-gpu_b200")
for prec, dt in (("D", torch.float64),):
    for k, m, batch in ((64, 16, 65536), (64, 32, 65536), (64, 64, 32768), (128, 32, 32768), (128, 128, 8192), (256, 64, 8192), (100, 40, 32768)):
        P = bench.make_spd(torch, batch, k, dt, 1)
        h0 = kb.Handle(); h0.potrf_batch_strided("L", k, P, k, k * k, batch, None); h0.destroy()
        B0 = torch.rand((batch, k, m), device="cuda", dtype=dt)   # side R: m x k, ldb = m
        B = torch.empty_like(B0)
        for v in (-1, 44):
            os.environ["KBLAS_B200_VARIANT"] = str(v)
            h = kb.Handle()
            for name, fn in (("potrs_R", lambda: h.potrs_batch_strided("R", "L", m, k, P, k, k * k, B, m, m * k, batch)),
                             ("trsm_RLT", lambda: h.trsm_batch_strided("R", "L", "T", "N", m, k, 1.0, P, k, k * k, B, m, m * k, batch)),
                             ("trsm_RLN", lambda: h.trsm_batch_strided("R", "L", "N", "N", m, k, 1.0, P, k, k * k, B, m, m * k, batch))):
                best, mean = timeit(fn, lambda: B.copy_(B0), reps=3)
                print(json.dumps({"op": prec + name, "k": k, "m": m, "batch": batch, "variant": v, "kernel": h.last_kernel, "ms_best": round(best, 4)}), flush=True)
            h.destroy()
        del P, B0, B
