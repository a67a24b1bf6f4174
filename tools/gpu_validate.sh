#!/usr/bin/env bash
# full validation on a B200 box (gpurun -- bash tools/gpu_validate.sh): build check, whole GPU suite, smoke, both bench arms
mkdir -p gpurun_out/validate
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/validate/smoke.log 2>&1; tail -2 gpurun_out/validate/smoke.log
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/validate/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/validate/pytest.log
tail -6 gpurun_out/validate/pytest.log | cut -c1-250
timeout 900 python bench.py --impl reference --steps 5 > gpurun_out/validate/bench_reference.json 2> gpurun_out/validate/bench_reference.err; cut -c1-200 gpurun_out/validate/bench_reference.json
timeout 1500 python bench.py > gpurun_out/validate/bench_ours.json 2> gpurun_out/validate/bench_ours.err; tail -2 gpurun_out/validate/bench_ours.err; cut -c1-300 gpurun_out/validate/bench_ours.json
