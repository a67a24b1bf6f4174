#!/usr/bin/env python
"""SASS opcode histogram of the kernels a default call path dispatches to (cuobjdump -sass on the built library); the
evidence table of B200_PROFILING.md: UBLKCP / UTMA* = TMA, LDGSTS = cp.async, DMMA / HMMA = legacy mma.sync tensor path,
UTC*MMA / LDTM / STTM = tcgen05 (none here: fp64 has no tcgen05 kind).  CPU only.
usage: python tools/sass_histogram.py > profiles/r02_sass_histograms.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kblas-gpu_b200", "lib", "libkblas-gpu.so")

# (label, regex on the demangled kernel name): the default dispatch of every BASELINE configuration
WANT = [
    ("dpotrf n=32 strided (headline)", r"potrf_reg_kernel<double, 32, 8, 8, 1, true, true, true>"),
    ("spotrf n=32 strided", r"potrf_reg_kernel<float, 32, 8, 4, 3, true, true, true>"),
    ("dpotrf n=8 strided", r"potrf_reg_kernel<double, 8, 8, 4, 4, true, true, false>"),
    ("dpotrs / dtrsm k=32 strided, side R fused fwd+bwd", r"tri_solve_dual_kernel<double, 32, false, 2, 4, true>"),
    ("dtrsm k=32 strided, side L forward (config 3, 16-byte accesses)", r"tri_left_vec_kernel<double, 32, 0, 2, 6>"),
    ("strsm k=32 strided, side L forward (config 3, 16-byte accesses)", r"tri_left_vec_kernel<float, 32, 0, 4, 6>"),
    ("dtrsm k=32 side L forward, element-wise fallback (pointer array / unaligned)", r"tri_solve_dual_kernel<double, 32, true, 0, 2, false>"),
    ("dtrsm k=32 strided, side R backward, one vector per lane", r"tri_right_vec_kernel<double, 32, 1, 4, 4>"),
    ("dpotrf n>32 pointer array (config 4)", r"potrf_panel_mma_kernel<double, 32, false>"),
    ("spotrf n>32 pointer array", r"potrf_panel_mma_kernel<float, 32, false>"),
    ("dposv solve n>32, 16 rhs rows, pointer array (config 4, DMMA)", r"tri_solve_mma_kernel<2, 16, 4, false>"),
    ("sposv solve n>32, 16 rhs rows, pointer array (FMA)", r"tri_solve_blocked_kernel<float, false, 2, 16, 4, false>"),
    ("lauum n>32 (blocked, in place)", r"lauum_blocked_kernel<double, 4, true>"),
    ("non-uniform dtrsm, side R forward", r"tri_solve_nonuniform_kernel<double, false, true>"),
    ("dpptrf n=32 packed (TMA in + out)", r"potrf_packed_kernel<double, 32, 8, 1, true, true, true, false>"),
    ("dpptrf n=8 packed (lane per matrix)", r"potrf_packed_lane_kernel<double, 8, 4, 4>"),
    ("dpotrf 32<n<=256 smem-resident (opt-in)", r"potrf_smem_kernel<8, 1, false>"),
    ("dgemm_batch / dsyrk_batch tile", r"gemm_tile_kernel<double, true, false, false, false, 4>"),
    ("sgemm_batch tile (3xTF32)", r"gemm_tile_kernel<float, true, false, false, false, 4>"),
]
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "DMMA", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "DFMA", "FFMA",
       "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    names = list(funcs)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    dem = [re.sub(r"\(bool\)1", "true", re.sub(r"\(bool\)0", "false", re.sub(r"\(int\)", "", d))).replace("void kblasx::", "") for d in dem]
    print(f"# SASS opcode histograms, {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB) / 2**20:.1f} MiB, {len(names)} kernels), cuobjdump -sass")
    print("# " + " ".join(f"{k:>7s}" for k in ["total"] + KEY))
    for label, rx in WANT:
        hit = [i for i, d in enumerate(dem) if re.search(rx.replace("(", r"\(").replace(")", r"\)"), d)]
        if not hit:
            print(f"{label}: NOT FOUND ({rx})")
            continue
        c = funcs[names[hit[0]]]
        print(f"{label}\n    {dem[hit[0]][:150]}")
        print("  " + " ".join(f"{v:7d}" for v in [sum(c.values())] + [sum(n for op, n in c.items() if op.startswith(k)) for k in KEY]))
    tot = collections.Counter()
    for c in funcs.values():
        tot.update(c)
    print("whole library")
    print("  " + " ".join(f"{v:7d}" for v in [sum(tot.values())] + [sum(n for op, n in tot.items() if op.startswith(k)) for k in KEY]))


if __name__ == "__main__":
    main()
