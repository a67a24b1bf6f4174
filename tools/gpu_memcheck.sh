#!/usr/bin/env bash
# compute-sanitizer memcheck over the small-batch GPU parity tests (out-of-bounds / misaligned accesses in the kernels)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "potrf_strided_vs_oracle or pointer_array_shuffled or strided_large_n or trsm_strided_vs_oracle or trsm_large_k or small_packed or potrs_and_posv or posv_pointer_array or host_pipeline" \
  > gpurun_out/memcheck.log 2>&1
echo "exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -20
tail -3 gpurun_out/memcheck.log
