#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_pipeline" 2>&1 | tail -5
for mode in tri full; do
  KBLAS_B200_HOSTCOPY=$mode python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/e2e_$mode.err | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$mode', json.dumps(d['e2e'])); print('   kernel', d['value'], d['roofline']['frac'])"
  tail -2 gpurun_out/e2e_$mode.err
done
KBLAS_B200_E2E=user python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('user-side pipeline', json.dumps(d['e2e']))"
