#!/usr/bin/env bash
# bench (both arms) + ncu evidence for the headline kernel
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json; tail -2 gpurun_out/bench_ours.err
python bench.py --impl reference-gpu > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-400; tail -2 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_reg -s 3 -c 1 -o gpurun_out/prof_potrf32_c -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
