#!/usr/bin/env bash
mkdir -p gpurun_out/r2q
timeout 600 python tools/bench_variants.py 40 trsm 32,24,16 > gpurun_out/r2q/bench_trsm_b.jsonl 2> gpurun_out/r2q/bench_trsm_b.err; tail -2 gpurun_out/r2q/bench_trsm_b.err
