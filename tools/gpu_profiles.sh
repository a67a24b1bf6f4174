#!/usr/bin/env bash
# round evidence: bench (both arms), launch list, ncu --set full of the dominant kernel of every configuration,
# timing tables (ours and the reference library).  Raw CSV exports come back in gpurun_out/; tools/ncu_summary.py
# condenses them into profiles/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -1 gpurun_out/bench_ours.err
python bench.py --impl reference-gpu --steps 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
cap() {  # tag, kernel regex, command...
  tag=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/prof_$tag -f "$@" > gpurun_out/ncu_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$tag.ncu-rep
}
ncu --set full --clock-control none --import-source on -k regex:potrf_reg -s 3 -c 1 -o gpurun_out/prof_potrf32 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
ncu -i gpurun_out/prof_potrf32.ncu-rep --page raw --csv > gpurun_out/prof_potrf32_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_potrf32.ncu-rep --page source --csv > gpurun_out/prof_potrf32_src.csv 2>/dev/null
rm -f gpurun_out/prof_potrf32.ncu-rep
cap potrs32 tri_solve_dual python tools/run_one.py potrs 32 1048576
cap trsm32RLT tri_solve_dual python tools/run_one.py trsm_RLT 32 1048576
cap trsm32LLN tri_solve_dual python tools/run_one.py trsm_LLN 32 1048576
cap potrs16 tri_solve_dual python tools/run_one.py potrs 16 1048576
cap potrf256d potrf_panel_mma python tools/run_one.py potrf_ptr 256 16384
F32=1 cap potrf256s potrf_panel_mma python tools/run_one.py potrf_ptr 256 16384
cap posv256d tri_solve_blocked python tools/run_one.py posv_ptr 256 16384
python tools/bench_variants.py -1 potrf > gpurun_out/t_potrf.jsonl 2>/dev/null
python tools/bench_variants.py -1 solve > gpurun_out/t_solve.jsonl 2>/dev/null
python tools/bench_variants.py -1 large > gpurun_out/t_large.jsonl 2>/dev/null
python tools/bench_reference_ops.py > gpurun_out/reference_ops.jsonl 2> gpurun_out/reference_ops.err
du -sh gpurun_out
