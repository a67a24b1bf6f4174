#!/usr/bin/env bash
# ncu --set full of one kernel: tools/gpu_ncu_one.sh <tag> <kernel regex> <run_one args...>
mkdir -p gpurun_out
tag=$1; rx=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/prof_$tag -f python tools/run_one.py "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
