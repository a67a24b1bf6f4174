#!/usr/bin/env python
"""Where a kernel's warp-stall samples land: `ncu --page source --csv` export -> samples per SASS opcode and
per contiguous address window (so that phases of an unrolled kernel show up), with the dominant stall reason."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
total = sum(int(r[col["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:110]); print("instructions:", len(body), " samples:", total)
by_op = collections.Counter(); ex_op = collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:2])
    by_op[op] += int(r[col["# Samples"]] or 0); ex_op[op] += int(r[col["Instructions Executed"]] or 0)
print("\nby opcode (samples %, warp-instructions executed)")
for op, s in by_op.most_common(14): print(f"  {op:22s} {100 * s / total:5.1f}%  {ex_op[op]:>12d}")
print(f"\nby window of {win} instructions (samples %, dominant stalls, opcode mix)")
for w0 in range(0, len(body), win):
    blk = body[w0:w0 + win]
    s = sum(int(r[col["# Samples"]] or 0) for r in blk)
    if s < 0.01 * total: continue
    st = collections.Counter()
    for r in blk:
        for h in stall_cols: st[h[6:]] += int(r[col[h]] or 0)
    ops = collections.Counter(re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]]).group(2) for r in blk)
    print(f"  [{w0:5d}] {100 * s / total:5.1f}%  " + " ".join(f"{k}:{100 * v / max(1, sum(st.values())):.0f}" for k, v in st.most_common(3)) + "   " + " ".join(f"{k}x{v}" for k, v in ops.most_common(4)))
