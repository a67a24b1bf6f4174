#!/usr/bin/env bash
mkdir -p gpurun_out/r2bb
timeout 600 python tools/bench_variants.py -1,46 trsm 16,8 > gpurun_out/r2bb/bench_trsm.jsonl 2> gpurun_out/r2bb/bench_trsm.err; tail -2 gpurun_out/r2bb/bench_trsm.err
