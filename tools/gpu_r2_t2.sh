#!/usr/bin/env bash
mkdir -p gpurun_out/r2t
timeout 900 python tools/bench_variants.py -1 large > gpurun_out/r2t/bench_large2.jsonl 2> gpurun_out/r2t/bench_large2.err; tail -2 gpurun_out/r2t/bench_large2.err
