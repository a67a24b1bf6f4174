#!/usr/bin/env python
"""time kblasDpotrf_batch_strided called directly on PINNED HOST memory (zero-copy over PCIe):
the kernel fetches only the lines of the lower triangle and writes only its sectors."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
kb = importlib.import_module("kblas-gpu_b200")
n, batch = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
P = bench.make_spd(torch, batch, n, torch.float64, 1)
hbuf = torch.empty((batch, n, n), dtype=torch.float64, pin_memory=True)
h = kb.Handle()
for it in range(3):
    hbuf.copy_(P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    rc = h.potrf_batch_strided("L", n, hbuf.data_ptr(), n, n * n, batch, None, prec="D")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert rc == 1
    L = torch.triu(hbuf[:512].cuda()).transpose(1, 2)
    A = P[:512].transpose(1, 2)
    res = ((A - L @ L.transpose(1, 2)).flatten(1).norm(dim=1) / A.flatten(1).norm(dim=1)).max().item()
    print(f"zero-copy in-place potrf: {ms:.1f} ms  {batch / ms / 1e3:.2f} M matrices/s  residual {res:.2e}  ({h.last_kernel})", flush=True)
