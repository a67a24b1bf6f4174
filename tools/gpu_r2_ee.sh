#!/usr/bin/env bash
mkdir -p gpurun_out/r2ee
cap() { tag=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r2ee/prof_$tag -f "$@" > gpurun_out/r2ee/ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2ee/prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2ee/prof_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2ee/prof_$tag.ncu-rep --page source --csv > gpurun_out/r2ee/prof_${tag}_src.csv 2>/dev/null
  rm -f gpurun_out/r2ee/prof_$tag.ncu-rep; tail -1 gpurun_out/r2ee/ncu_$tag.log; }
F32=1 cap spotrf32 potrf_reg python tools/run_one.py potrf 32 1048576
cap dpotrs32 tri_solve_dual python tools/run_one.py potrs 32 1048576
