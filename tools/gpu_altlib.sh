#!/usr/bin/env bash
# time the large-n configurations with alternative builds of the library (swap the .so in place on the box copy)
cd kblas-gpu_b200/lib
cp libkblas-gpu.so libkblas-gpu-base.so
for v in base $@; do
  cp libkblas-gpu-$v.so libkblas-gpu.so
  echo "== build $v"
  (cd ../..; python tools/bench_variants.py -1 large 2>/dev/null | python tools/_pv2.py | grep -E "posv")
done
