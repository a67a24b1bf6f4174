#!/usr/bin/env bash
# round 2, GPU call E: whole suite; bench line (configs sweep, packed e2e); launch list + ncu --set full of the headline
# kernel and of the packed kernel; ncu of the opt-in shared-memory resident kernel (DRAM traffic evidence)
mkdir -p gpurun_out/r2e
timeout 2700 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2e/pytest.log
tail -12 gpurun_out/r2e/pytest.log | cut -c1-250
timeout 1500 python bench.py > gpurun_out/r2e/bench_ours.json 2> gpurun_out/r2e/bench_ours.err; grep -c configs gpurun_out/r2e/bench_ours.err; tail -3 gpurun_out/r2e/bench_ours.err
timeout 900 python bench.py --impl reference --steps 5 > gpurun_out/r2e/bench_reference.json 2> gpurun_out/r2e/bench_reference.err; cut -c1-300 gpurun_out/r2e/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2e/launches_bench_dpotrf32.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/r2e/b_ncu.log 2>&1
cap() {  # tag, kernel regex, command...
  tag=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r2e/prof_$tag -f "$@" > gpurun_out/r2e/ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2e/prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2e/prof_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2e/prof_$tag.ncu-rep --page source --csv > gpurun_out/r2e/prof_${tag}_src.csv 2>/dev/null
  rm -f gpurun_out/r2e/prof_$tag.ncu-rep
  tail -2 gpurun_out/r2e/ncu_$tag.log
}
cap dpotrf32 potrf_reg_kernel python tools/run_one.py potrf 32 1048576
cap dpptrf32 potrf_packed_kernel python tools/run_one.py pptrf 32 1048576
cap dpptrf8 potrf_packed_lane_kernel python tools/run_one.py pptrf 8 1048576
KBLAS_B200_VARIANT=33 cap dpotrf256_smem potrf_smem_kernel python tools/run_one.py potrf_ptr 256 16384
cap dpotrf256_dmma potrf_panel_mma_kernel python tools/run_one.py potrf_ptr 256 16384
du -sh gpurun_out/r2e
