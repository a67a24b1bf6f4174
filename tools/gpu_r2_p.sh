#!/usr/bin/env bash
mkdir -p gpurun_out/r2s
cap() { tag=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r2s/prof_$tag -f "$@" > gpurun_out/r2s/ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2s/prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2s/prof_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2s/prof_$tag.ncu-rep --page source --csv > gpurun_out/r2s/prof_${tag}_src.csv 2>/dev/null
  rm -f gpurun_out/r2s/prof_$tag.ncu-rep; tail -1 gpurun_out/r2s/ncu_$tag.log; }
F32=1 cap strsm32_LLN tri_left_vec python tools/run_one.py trsm_LLN 32 1048576
cap dtrsm32_LLN tri_left_vec python tools/run_one.py trsm_LLN 32 1048576
