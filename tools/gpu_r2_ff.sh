#!/usr/bin/env bash
mkdir -p gpurun_out/r2ff
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf_strided_vs_oracle or pointer_array_shuffled or golden or live_against or non_spd or element_exact" > gpurun_out/r2ff/pytest.log 2>&1; tail -3 gpurun_out/r2ff/pytest.log
timeout 600 python tools/bench_variants.py -1 potrf > gpurun_out/r2ff/bench_potrf.jsonl 2> gpurun_out/r2ff/bench_potrf.err; tail -2 gpurun_out/r2ff/bench_potrf.err
