#!/usr/bin/env bash
# full GPU parity suite + large-n timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python tools/bench_variants.py -1 large > gpurun_out/large.jsonl 2> gpurun_out/large.err; python - <<'PY'
import json
for l in open('gpurun_out/large.jsonl'):
    d=json.loads(l); print(d['op'],d['n'],d['batch'],d['kernel'],round(d['ms_best'],3),'ms',round(d['Mprob_s'],2),'M/s',round(d['TFLOPs'],2),'TF',round(d['frac_hbm'],3))
PY
tail -3 gpurun_out/large.err
