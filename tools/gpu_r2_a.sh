#!/usr/bin/env bash
# round 2, GPU call A: parity suite (new pointer-array / large-n / link-compat tests), round-2 goldens, reference test
# programs on both libraries (raw output kept), a bench line for this box
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a/smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
timeout 600 python tests/golden/make_golden_r2.py gpurun_out/golden > gpurun_out/r2a/golden.log 2>&1; tail -3 gpurun_out/r2a/golden.log
for who in ours ref; do for t in dpotrf dtrsm dpotrs dposv spotrf; do
  for s in "" "-s"; do
    echo "== $who $t $s" >> gpurun_out/r2a/refprogs.txt
    timeout 120 oracle/_ref/bin/$who/test_${t}_batch -N 32 --batch 1000 -c --nruns 2 -SR $s >> gpurun_out/r2a/refprogs.txt 2>&1
  done
done; done
timeout 120 oracle/_ref/bin/ours/test_dpotrf_batch --range 8:32:8 --batch 100000 -s -w --nruns 5 >> gpurun_out/r2a/refprogs.txt 2>&1
timeout 120 oracle/_ref/bin/ref/test_dpotrf_batch --range 8:32:8 --batch 100000 -s -w --nruns 5 >> gpurun_out/r2a/refprogs.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2a/bench_ours.json 2> gpurun_out/r2a/bench_ours.err; cut -c1-600 gpurun_out/r2a/bench_ours.json
