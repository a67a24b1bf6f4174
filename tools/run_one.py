#!/usr/bin/env python
"""run one op twice (for ncu): python tools/run_one.py potrf|pptrf|potrs|trsm_LLN|posv_ptr|potrf_ptr n [batch]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
kb = importlib.import_module("kblas-gpu_b200")
op, n = sys.argv[1], int(sys.argv[2])
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
dt = torch.float32 if os.environ.get("F32") else torch.float64
prec = "S" if os.environ.get("F32") else "D"
m = (16 if op.endswith("ptr") else n)
P = bench.make_spd(torch, batch, n, dt, 1)
B = torch.rand((batch, n, m), device="cuda", dtype=dt)
h = kb.Handle()
h.posv_batch_wsquery("R", m, n, batch); h.posv_batch_strided_wsquery("R", m, n, batch); h.allocate_workspace()
for it in range(2):
    A = P.clone()
    if op == "potrf":
        rc = h.potrf_batch_strided("L", n, A, n, n * n, batch, None)
    elif op == "pptrf":
        sz = n * (n + 1) // 2
        PP = torch.empty((batch, sz), device="cuda", dtype=dt)
        h.tri_pack_batch_strided("L", n, A, n, n * n, PP, sz, batch)
        rc = h.pptrf_batch_strided("L", n, PP, sz, batch, None)
    elif op == "potrs":
        h.potrf_batch_strided("L", n, A, n, n * n, batch, None)
        rc = h.potrs_batch_strided("R", "L", m, n, A, n, n * n, B, m, m * n, batch)
    elif op.startswith("trsm_"):
        s, _, t = op[5], op[6], op[7]
        h.potrf_batch_strided("L", n, A, n, n * n, batch, None)
        rc = h.trsm_batch_strided(s, "L", t, "N", m, n, 0.28, A, n, n * n, B, m, m * n, batch)
    else:
        pa = (A.data_ptr() + torch.arange(batch, device="cuda") * (n * n * A.element_size())).contiguous()
        pb = (B.data_ptr() + torch.arange(batch, device="cuda") * (m * n * B.element_size())).contiguous()
        rc = (h.posv_batch("R", "L", m, n, pa, n, pb, m, batch, None, prec=prec) if op == "posv_ptr"
              else h.potrf_batch("L", n, pa, n, batch, None, prec=prec))
    torch.cuda.synchronize()
    assert rc == 1, rc
print("ok", h.last_kernel)
