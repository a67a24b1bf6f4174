#!/usr/bin/env bash
# quick GPU check: parity tests (n<=32 subset) + kernel timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_potrf_strided_large_n --deselect tests/test_gpu_parity.py::test_trsm_large_k --deselect tests/test_gpu_parity.py::test_posv_pointer_array_large_n > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python tools/bench_variants.py ${1:--1} potrf > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; python - <<'PY'
import json
for l in open('gpurun_out/variants.jsonl'):
    d=json.loads(l); print(d['op'],d['n'],d['variant'],d['kernel'],round(d['ms_best'],3),round(d['Mmat_s'],1),round(d['frac'],3))
PY
tail -3 gpurun_out/variants.err
python tools/bench_variants.py -1 solve > gpurun_out/solve.jsonl 2> gpurun_out/solve.err; python - <<'PY'
import json
for l in open('gpurun_out/solve.jsonl'):
    d=json.loads(l); print(d['op'],d['n'],d['kernel'],round(d['ms_best'],3),round(d['Mprob_s'],1),round(d['frac'],3))
PY
tail -3 gpurun_out/solve.err
