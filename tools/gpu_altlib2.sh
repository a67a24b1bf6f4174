#!/usr/bin/env bash
# time the small solves with alternative builds of the library (swap the .so in place on the box copy)
cd kblas-gpu_b200/lib
cp libkblas-gpu.so libkblas-gpu-base.so
for v in base $@; do
  cp libkblas-gpu-$v.so libkblas-gpu.so
  echo "== build $v"
  (cd ../..; python tools/bench_variants.py ${VARIANTS:--1} solve ${NS:-16,8} 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d['op']+'_v'+str(d['variant']),d['n'],d['kernel'],round(d['ms_best'],3))")
done
