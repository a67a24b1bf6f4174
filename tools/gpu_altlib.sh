#!/usr/bin/env bash
# time potrf_ptr large n with alternative builds of the library (swap the .so in place on the box copy)
./tools/bin/microbench B 2>&1 | grep -E "TF32|clock" > gpurun_out/micro_tf32.txt; cat gpurun_out/micro_tf32.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
cd kblas-gpu_b200/lib
cp libkblas-gpu.so libkblas-gpu-base.so
for v in base $@; do
  cp libkblas-gpu-$v.so libkblas-gpu.so
  echo "== build $v"
  (cd ../..; python tools/bench_variants.py -1 large 2>/dev/null | python tools/_pl.py | grep -E "potrf|posv")
done
