#!/usr/bin/env bash
mkdir -p gpurun_out/r2nn
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "left_side or potrs or posv or random_shapes" > gpurun_out/r2nn/pytest.log 2>&1; tail -3 gpurun_out/r2nn/pytest.log
timeout 600 python tools/bench_variants.py -1,53 trsm 32,24 > gpurun_out/r2nn/bench.jsonl 2> gpurun_out/r2nn/bench.err; tail -2 gpurun_out/r2nn/bench.err
