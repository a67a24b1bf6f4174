#!/usr/bin/env bash
# round 2, GPU call B: packed-layout parity + variant timings, backtrace of the reference test program crash at N=32,
# slimmed round-2 goldens
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests/test_packed.py -m gpu -x -q > gpurun_out/r2b/pytest_packed.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2b/pytest_packed.log
tail -15 gpurun_out/r2b/pytest_packed.log
timeout 900 python tools/bench_packed.py 20,21,22,23 32,24,16,8 > gpurun_out/r2b/bench_packed.jsonl 2> gpurun_out/r2b/bench_packed.err
cat gpurun_out/r2b/bench_packed.jsonl; tail -3 gpurun_out/r2b/bench_packed.err
timeout 300 /usr/local/cuda/bin/cuda-gdb -batch -ex run -ex bt --args oracle/_ref/bin/ref/test_dpotrf_batch -N 32 --batch 100 -s > gpurun_out/r2b/gdb_ref.txt 2>&1
tail -30 gpurun_out/r2b/gdb_ref.txt
timeout 600 python tests/golden/make_golden_r2.py gpurun_out/golden > gpurun_out/r2b/golden.log 2>&1; tail -2 gpurun_out/r2b/golden.log; ls -la gpurun_out/golden
