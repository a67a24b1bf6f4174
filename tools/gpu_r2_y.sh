#!/usr/bin/env bash
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_inverse_family.py -m gpu -x -q > gpurun_out/r2y/pytest.log 2>&1; tail -15 gpurun_out/r2y/pytest.log
