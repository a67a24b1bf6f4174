#!/usr/bin/env bash
# round 2, GPU call L: persistent tri_dual (L2 prefetch of the next task) -- parity + A/B timing against the round-1 launch shape
mkdir -p gpurun_out/r2l
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trsm or potrs or posv or golden or live" > gpurun_out/r2l/pytest.log 2>&1; tail -5 gpurun_out/r2l/pytest.log | cut -c1-250
timeout 900 python tools/bench_variants.py 40,-1 solve 32,24,16 > gpurun_out/r2l/solve_ab.jsonl 2> gpurun_out/r2l/solve_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r2l/solve_ab.jsonl'):
    r=json.loads(l); print(f"{r['op']:10s} n={r['n']:3d} v={r['variant']:3d} {r['kernel']:20s} best {r['ms_best']:.3f} mean {r['ms_mean']:.3f} frac {r['frac']:.3f}")
PY
