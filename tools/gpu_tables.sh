#!/usr/bin/env bash
# timing tables of every BASELINE configuration (ours) + fp32 accuracy of the tensor-path panel kernel
mkdir -p gpurun_out
python tools/accuracy_f32.py 64 128 256 40 100 2>&1 | tee gpurun_out/accuracy_f32.txt
python tools/bench_variants.py -1 potrf > gpurun_out/t_potrf.jsonl 2>/dev/null
python tools/bench_variants.py -1 solve > gpurun_out/t_solve.jsonl 2>/dev/null
python tools/bench_variants.py -1 large > gpurun_out/t_large.jsonl 2>/dev/null
cat gpurun_out/t_solve.jsonl | python tools/_pl.py
