#!/usr/bin/env bash
# round 2, GPU call H: re-scheduled shared-memory resident Cholesky (strips + dynamic hand-out + pivot look-ahead): parity, timings
mkdir -p gpurun_out/r2h
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n_kernel_variants or element_exact or left_side" > gpurun_out/r2h/pytest.log 2>&1; tail -6 gpurun_out/r2h/pytest.log | cut -c1-300
timeout 900 python tools/bench_large.py 30,31,34,32,35,33 64,128,256 potrf_ptr > gpurun_out/r2h/bench_large.jsonl 2> gpurun_out/r2h/bench_large.err
cut -c1-230 gpurun_out/r2h/bench_large.jsonl; tail -3 gpurun_out/r2h/bench_large.err
