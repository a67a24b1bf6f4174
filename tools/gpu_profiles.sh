#!/usr/bin/env bash
# ncu evidence for the dominant kernels of each BASELINE configuration + timing tables (ours and reference)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -1 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potrf_reg -c 12 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_reg -s 3 -c 1 -o gpurun_out/prof_potrf32 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:tri_solve_small -s 1 -c 1 -o gpurun_out/prof_potrs32 -f python tools/run_one.py potrs 32 > gpurun_out/ncu_potrs.log 2>&1
ncu --set full --clock-control none -k regex:tri_solve_small -s 1 -c 1 -o gpurun_out/prof_trsm32 -f python tools/run_one.py trsm_LLN 32 > gpurun_out/ncu_trsm.log 2>&1
ncu --set full --clock-control none -k regex:potrf_panel_dmma -s 1 -c 1 -o gpurun_out/prof_potrf256 -f python tools/run_one.py potrf_ptr 256 16384 > gpurun_out/ncu_p256.log 2>&1
python tools/bench_variants.py -1 potrf > gpurun_out/t_potrf.jsonl 2>/dev/null
python tools/bench_variants.py -1 solve > gpurun_out/t_solve.jsonl 2>/dev/null
python tools/bench_variants.py -1 large > gpurun_out/t_large.jsonl 2>/dev/null
for r in potrf32 potrs32 trsm32 potrf256; do
  ncu -i gpurun_out/prof_$r.ncu-rep --page raw --csv > gpurun_out/prof_${r}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_potrf32.ncu-rep --page source --csv > gpurun_out/prof_potrf32_src.csv 2>/dev/null
rm -f gpurun_out/prof_potrs32.ncu-rep gpurun_out/prof_trsm32.ncu-rep gpurun_out/prof_potrf256.ncu-rep
du -sh gpurun_out
