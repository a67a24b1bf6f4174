#!/usr/bin/env bash
mkdir -p gpurun_out/r2q
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "left_side_16_byte" > gpurun_out/r2q/pytest_vec.log 2>&1; tail -5 gpurun_out/r2q/pytest_vec.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q/pytest_all.log 2>&1; tail -3 gpurun_out/r2q/pytest_all.log
