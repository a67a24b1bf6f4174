#!/usr/bin/env bash
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "posv or config4 or live_large or golden" > gpurun_out/r2z/pytest.log 2>&1; tail -3 gpurun_out/r2z/pytest.log
timeout 900 python tools/bench_variants.py -1,45 large > gpurun_out/r2z/bench_large.jsonl 2> gpurun_out/r2z/bench_large.err; tail -2 gpurun_out/r2z/bench_large.err
