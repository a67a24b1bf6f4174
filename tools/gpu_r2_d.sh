#!/usr/bin/env bash
# round 2, GPU call D: the shared-memory resident large-n Cholesky -- parity, variant timings; reference programs (O0 build);
# bench with the fixed configs sweep
mkdir -p gpurun_out/r2d
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_n or posv_pointer_array_large or live_large or golden_r2 or config4 or two_devices" > gpurun_out/r2d/pytest_large.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2d/pytest_large.log
tail -15 gpurun_out/r2d/pytest_large.log
timeout 900 python tools/bench_large.py 30,-1,31,32,33 64,128,256 potrf_ptr > gpurun_out/r2d/bench_large.jsonl 2> gpurun_out/r2d/bench_large.err
cat gpurun_out/r2d/bench_large.jsonl; tail -3 gpurun_out/r2d/bench_large.err
timeout 600 python -m pytest tests/test_link_compat.py -m gpu -x -q > gpurun_out/r2d/pytest_link.log 2>&1; tail -5 gpurun_out/r2d/pytest_link.log
for t in dpotrf dposv; do stdbuf -o0 oracle/_ref/bin/ours/test_${t}_batch --range 32:256:32 --batch 1000 -SR -c --nruns 2 > gpurun_out/r2d/refprog_ours_$t.txt 2>&1; stdbuf -o0 oracle/_ref/bin/ref/test_${t}_batch --range 32:256:32 --batch 1000 -SR -c --nruns 2 > gpurun_out/r2d/refprog_ref_$t.txt 2>&1; done
tail -9 gpurun_out/r2d/refprog_ours_dposv.txt gpurun_out/r2d/refprog_ref_dposv.txt
timeout 1200 python bench.py > gpurun_out/r2d/bench_ours.json 2> gpurun_out/r2d/bench_ours.err; grep configs gpurun_out/r2d/bench_ours.err
