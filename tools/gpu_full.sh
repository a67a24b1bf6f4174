#!/usr/bin/env bash
# full GPU parity suite + timing tables
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python tools/bench_variants.py -1 potrf > gpurun_out/t_potrf.jsonl 2>/dev/null
python tools/bench_variants.py -1 solve > gpurun_out/t_solve.jsonl 2>/dev/null
python tools/bench_variants.py -1 large > gpurun_out/t_large.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
