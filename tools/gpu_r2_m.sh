#!/usr/bin/env bash
mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n_kernel_variants and 36" > gpurun_out/r2m/pytest.log 2>&1; tail -3 gpurun_out/r2m/pytest.log | cut -c1-200
timeout 600 python tools/bench_large.py 30,36 64,128,256 potrf_ptr > gpurun_out/r2m/bench_large.jsonl 2> gpurun_out/r2m/bench_large.err; cut -c1-200 gpurun_out/r2m/bench_large.jsonl; tail -2 gpurun_out/r2m/bench_large.err
