#!/usr/bin/env bash
mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n or variants or live_large or config4 or posv or golden" > gpurun_out/r2x/pytest.log 2>&1; tail -3 gpurun_out/r2x/pytest.log
timeout 900 python tools/bench_variants.py -1 large > gpurun_out/r2x/bench_large.jsonl 2> gpurun_out/r2x/bench_large.err; tail -2 gpurun_out/r2x/bench_large.err
