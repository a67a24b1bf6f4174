#!/usr/bin/env python
"""fp32 potrf accuracy of the kernel variants against an fp64 factorisation of the same matrices:
python tools/accuracy_f32.py [n ...]   (env KBLAS_B200_VARIANT is set per run: default / 9 = FMA panel kernel)"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ns = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
batch = 2048
eps = 2.0 ** -23
for n in ns:
    P = bench.make_spd(torch, batch, n, torch.float32, 1)          # (batch, n, n), symmetric
    L64 = torch.linalg.cholesky(P.double())
    nrm = torch.linalg.matrix_norm(P.double())                       # Frobenius, per matrix
    for name, v in (("tensor-path", None), ("fma-path", "9")):
        if v is None:
            os.environ.pop("KBLAS_B200_VARIANT", None)
        else:
            os.environ["KBLAS_B200_VARIANT"] = v
        kb = importlib.import_module("kblas-gpu_b200")
        h = kb.Handle()
        h.potrf_batch_strided_wsquery(n, batch); h.allocate_workspace()
        A = P.clone()
        assert h.potrf_batch_strided("L", n, A, n, n * n, batch, None) == 1
        torch.cuda.synchronize()
        # memory is column-major: A[b] viewed row-major is the transpose -> lower factor = triu of the view, transposed
        L = torch.triu(A).transpose(1, 2).double()
        dL = ((L - L64).abs().amax(dim=(1, 2)) / nrm).max().item()
        res = (torch.linalg.matrix_norm(L @ L.transpose(1, 2) - P.double()) / nrm).max().item()
        print(f"n={n:4d} {name:12s} {h.last_kernel:28s} max|L-L64|/|A| = {dL / eps:8.2f} eps   |A-LL^T|/|A| = {res / eps:8.2f} eps"
              f"   (limits {100 * n} / {10 * n} eps)")
        h.destroy()
