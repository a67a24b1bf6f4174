#!/usr/bin/env bash
# end-of-round check: full GPU suite, smoke, both bench arms, timing tables
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python bench.py --impl reference-gpu --steps 10 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; cut -c1-200 gpurun_out/bench_refgpu.json
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-300 gpurun_out/bench_ours.json
python tools/bench_variants.py -1 potrf > gpurun_out/t_potrf.jsonl 2>/dev/null
python tools/bench_variants.py -1 solve > gpurun_out/t_solve.jsonl 2>/dev/null
python tools/bench_variants.py -1 large > gpurun_out/t_large.jsonl 2>/dev/null
