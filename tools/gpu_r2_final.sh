#!/usr/bin/env bash
# round 2, final validation: build check, whole GPU suite, smoke, both bench arms
mkdir -p gpurun_out/r2fin
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2fin/smoke.log 2>&1; tail -2 gpurun_out/r2fin/smoke.log
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/r2fin/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2fin/pytest.log
tail -6 gpurun_out/r2fin/pytest.log | cut -c1-250
timeout 900 python bench.py --impl reference --steps 5 > gpurun_out/r2fin/bench_reference.json 2> gpurun_out/r2fin/bench_reference.err; cut -c1-200 gpurun_out/r2fin/bench_reference.json
timeout 1500 python bench.py > gpurun_out/r2fin/bench_ours.json 2> gpurun_out/r2fin/bench_ours.err; tail -2 gpurun_out/r2fin/bench_ours.err; cut -c1-300 gpurun_out/r2fin/bench_ours.json
