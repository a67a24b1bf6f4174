#!/usr/bin/env bash
mkdir -p gpurun_out/r2u
cap() { tag=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r2u/prof_$tag -f "$@" > gpurun_out/r2u/ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2u/prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2u/prof_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2u/prof_$tag.ncu-rep --page source --csv > gpurun_out/r2u/prof_${tag}_src.csv 2>/dev/null
  rm -f gpurun_out/r2u/prof_$tag.ncu-rep; tail -1 gpurun_out/r2u/ncu_$tag.log; }
cap dposv256_mma tri_solve_mma python tools/run_one.py posv_ptr 256 32768
cap dposv64_mma tri_solve_mma python tools/run_one.py posv_ptr 64 65536
