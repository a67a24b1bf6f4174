#!/usr/bin/env bash
# round 2, GPU call C: whole parity suite, bench line with the configs sweep, packed n=8/16 defaults, gdb of the
# reference test program on OUR library, ncu of the packed kernel
mkdir -p gpurun_out/r2c
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2c/pytest.log
tail -12 gpurun_out/r2c/pytest.log
timeout 300 /usr/local/cuda/bin/cuda-gdb -batch -ex run -ex bt --args oracle/_ref/bin/ours/test_dpotrf_batch -N 32 --batch 100 -s > gpurun_out/r2c/gdb_ours.txt 2>&1
tail -12 gpurun_out/r2c/gdb_ours.txt
timeout 600 python tools/bench_packed.py -1,24 8,16 > gpurun_out/r2c/bench_packed_small.jsonl 2> gpurun_out/r2c/bench_packed_small.err
cat gpurun_out/r2c/bench_packed_small.jsonl | cut -c1-260
timeout 1200 python bench.py > gpurun_out/r2c/bench_ours.json 2> gpurun_out/r2c/bench_ours.err; tail -40 gpurun_out/r2c/bench_ours.err
ncu --set full --clock-control none --import-source on -k regex:potrf_packed_kernel -s 2 -c 1 -o gpurun_out/r2c/prof_pptrf32 -f python tools/run_one.py pptrf 32 1048576 > gpurun_out/r2c/ncu_pptrf32.log 2>&1
ncu -i gpurun_out/r2c/prof_pptrf32.ncu-rep --page raw --csv > gpurun_out/r2c/prof_pptrf32_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c/prof_pptrf32.ncu-rep --page source --csv > gpurun_out/r2c/prof_pptrf32_src.csv 2>/dev/null
rm -f gpurun_out/r2c/prof_pptrf32.ncu-rep
tail -3 gpurun_out/r2c/ncu_pptrf32.log
