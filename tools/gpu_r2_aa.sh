#!/usr/bin/env bash
mkdir -p gpurun_out/r2aa
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "upper_and_unit or return_codes or nonuniform" > gpurun_out/r2aa/pytest_new.log 2>&1; tail -15 gpurun_out/r2aa/pytest_new.log
