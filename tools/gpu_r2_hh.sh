#!/usr/bin/env bash
mkdir -p gpurun_out/r2hh
for v in 48 49; do KBLAS_B200_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf_strided_large_n or live_large" > gpurun_out/r2hh/pytest_v$v.log 2>&1; tail -2 gpurun_out/r2hh/pytest_v$v.log; done
timeout 900 python tools/bench_variants.py -1,48,49 large > gpurun_out/r2hh/bench_large.jsonl 2> gpurun_out/r2hh/bench_large.err; tail -2 gpurun_out/r2hh/bench_large.err
