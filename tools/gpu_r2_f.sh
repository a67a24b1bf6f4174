#!/usr/bin/env bash
# round 2, GPU call F (1 GPU): BASELINE config 5 at N = 1 -- strided dpotrf n=32, batch 8M (64 GiB) on ONE B200; plus the
# remaining parity checks of this build
mkdir -p gpurun_out/r2f
timeout 600 python -m pytest tests/test_link_compat.py tests/test_gemm_syrk.py -m gpu -x -q > gpurun_out/r2f/pytest.log 2>&1; tail -4 gpurun_out/r2f/pytest.log | cut -c1-200
timeout 1500 python bench.py --batch 8388608 --steps 6 --warmup 3 > gpurun_out/r2f/bench_8M_1gpu.json 2> gpurun_out/r2f/bench_8M_1gpu.err; tail -2 gpurun_out/r2f/bench_8M_1gpu.err; cut -c1-400 gpurun_out/r2f/bench_8M_1gpu.json
