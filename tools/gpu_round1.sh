#!/usr/bin/env bash
# first GPU session: smoke, parity tests, goldens, bench (both arms), variants, ncu evidence
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_potrf_strided_large_n --deselect tests/test_gpu_parity.py::test_trsm_large_k --deselect tests/test_gpu_parity.py::test_posv_pointer_array_large_n > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; tail -3 gpurun_out/golden.log
python tools/bench_variants.py -1,1,2 potrf > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; cat gpurun_out/variants.jsonl
python tools/bench_variants.py -1 solve > gpurun_out/solve.jsonl 2> gpurun_out/solve.err; cat gpurun_out/solve.jsonl
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_reg -s 3 -c 1 -o gpurun_out/prof_potrf32 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
