#!/usr/bin/env bash
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_k or posv or potrs or random_shapes or config4 or live_large" > gpurun_out/r2w/pytest.log 2>&1; tail -3 gpurun_out/r2w/pytest.log
timeout 900 python tools/bench_solve_large.py > gpurun_out/r2w/solve_large.jsonl 2> gpurun_out/r2w/solve_large.err; tail -2 gpurun_out/r2w/solve_large.err
timeout 900 python tools/bench_variants.py -1 large > gpurun_out/r2w/bench_large.jsonl 2> gpurun_out/r2w/bench_large.err; tail -2 gpurun_out/r2w/bench_large.err
