#!/usr/bin/env bash
# ours vs the unmodified reference GPU library (oracle/_ref/libkblas_ref.so) on every BASELINE configuration, same box, same buffers
mkdir -p gpurun_out/vsref
timeout 900 python tools/bench_variants.py -1 potrf > gpurun_out/vsref/t_potrf.jsonl 2> gpurun_out/vsref/t_potrf.err
timeout 900 python tools/bench_variants.py -1 trsm 32,24,16,8 > gpurun_out/vsref/t_solve.jsonl 2> gpurun_out/vsref/t_solve.err
timeout 900 python tools/bench_variants.py -1 large > gpurun_out/vsref/t_large.jsonl 2> gpurun_out/vsref/t_large.err
timeout 1800 python tools/bench_reference_ops.py > gpurun_out/vsref/reference_ops.jsonl 2> gpurun_out/vsref/reference_ops.err; tail -2 gpurun_out/vsref/reference_ops.err
python tools/compare_tables.py gpurun_out/vsref > gpurun_out/vsref/ours_vs_reference.txt; tail -5 gpurun_out/vsref/ours_vs_reference.txt
