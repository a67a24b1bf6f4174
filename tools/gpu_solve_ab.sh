#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trsm or potrs or posv" 2>&1 | tail -5
python tools/bench_variants.py -1,8 solve 32 > gpurun_out/t_solve_ab.jsonl 2>gpurun_out/t_solve_ab.err; tail -3 gpurun_out/t_solve_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/t_solve_ab.jsonl'):
    d=json.loads(l); print(d['op'],d['n'],d['variant'],d['kernel'],round(d['ms_best'],3))
PY
