#!/usr/bin/env bash
mkdir -p gpurun_out/r2cc
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2cc/pytest_all.log 2>&1; tail -3 gpurun_out/r2cc/pytest_all.log
