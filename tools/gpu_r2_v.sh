#!/usr/bin/env bash
mkdir -p gpurun_out/r2v
timeout 900 python tools/bench_solve_large.py > gpurun_out/r2v/solve_large.jsonl 2> gpurun_out/r2v/solve_large.err; tail -2 gpurun_out/r2v/solve_large.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v/pytest_all.log 2>&1; tail -3 gpurun_out/r2v/pytest_all.log
