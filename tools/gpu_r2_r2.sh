#!/usr/bin/env bash
mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_vector or left_side_16 or trsm or potrs" > gpurun_out/r2r/pytest.log 2>&1; tail -5 gpurun_out/r2r/pytest.log
timeout 600 python tools/bench_variants.py -1 trsm 32,24,16 > gpurun_out/r2r/bench_trsm_final.jsonl 2> gpurun_out/r2r/bench_trsm_final.err; tail -2 gpurun_out/r2r/bench_trsm_final.err
