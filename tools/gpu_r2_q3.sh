#!/usr/bin/env bash
mkdir -p gpurun_out/r2q
timeout 600 python tools/bench_variants.py -1,40 trsm 32,24,16 > gpurun_out/r2q/bench_trsm_c.jsonl 2> gpurun_out/r2q/bench_trsm_c.err; tail -2 gpurun_out/r2q/bench_trsm_c.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trsm or potrs or posv or left" > gpurun_out/r2q/pytest_default.log 2>&1; tail -3 gpurun_out/r2q/pytest_default.log
