#!/usr/bin/env python
"""ours vs reference timing tables (gpurun_out/t_*.jsonl vs gpurun_out/reference_ops.jsonl), ms per call"""
import json, sys, os
d0 = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ours, ref = {}, {}
for f in ("t_potrf.jsonl", "t_solve.jsonl", "t_large.jsonl"):
    for l in open(os.path.join(d0, f)):
        d = json.loads(l); ours[(d["op"], d["n"])] = (d["ms_best"], d["kernel"], d.get("frac_hbm"), d.get("TFLOPs"))
for l in open(os.path.join(d0, "reference_ops.jsonl")):
    d = json.loads(l); ref[(d["op"], d["n"])] = d["ms_best"]
print(f"{'op':12s} {'n':>4s} {'ours ms':>9s} {'ref ms':>9s} {'speedup':>8s} {'frac_hbm':>8s}  kernel")
for k in sorted(ours):
    o = ours[k]; r = ref.get(k)
    print(f"{k[0]:12s} {k[1]:4d} {o[0]:9.3f} {r if r is None else round(r, 3)!s:>9s} {'' if r is None else 'x%.2f' % (r / o[0]):>8s} "
          f"{'' if o[2] is None else '%.3f' % o[2]:>8s}  {o[1]}")
