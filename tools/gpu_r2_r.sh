#!/usr/bin/env bash
mkdir -p gpurun_out/r2r
timeout 600 python tools/bench_variants.py -1,42 trsm 32 > gpurun_out/r2r/bench_trsm_r.jsonl 2> gpurun_out/r2r/bench_trsm_r.err; tail -2 gpurun_out/r2r/bench_trsm_r.err
