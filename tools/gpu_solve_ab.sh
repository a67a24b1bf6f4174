#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trsm or potrs or posv or golden or live" 2>&1 | tail -5
python tools/bench_variants.py ${VARIANTS:--1} solve ${NS:-32} 2>/dev/null | python tools/_pv2.py
