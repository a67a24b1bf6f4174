#!/usr/bin/env bash
# compute-sanitizer over the small-batch GPU parity tests of every kernel family: memcheck (out-of-bounds / misaligned
# accesses, including the TMA bulk copies of the packed kernels) and racecheck (shared-memory hazards of the staged kernels)
mkdir -p gpurun_out/sanitize
K="potrf_strided_vs_oracle or pointer_array_shuffled or strided_large_n or large_n_kernel_variants or trsm_strided_vs_oracle or trsm_large_k or trsm_pointer_array_default or potrs_posv_pointer_array_default or potrs_and_posv or posv_pointer_array or left_side or one_vector or upper_and_unit or nonuniform or host_pipeline or pointer_and_value_helpers"
timeout 2400 compute-sanitizer --tool memcheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/sanitize/memcheck_parity.log 2>&1
echo "memcheck parity exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize/memcheck_parity.log | tail -3
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_packed.py tests/test_gemm_syrk.py tests/test_inverse_family.py -x -q -m gpu -k "not full_size" > gpurun_out/sanitize/memcheck_packed_gemm.log 2>&1
echo "memcheck packed+gemm exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize/memcheck_packed_gemm.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_packed.py -x -q -m gpu -k "strided_vs_oracle or pointer_array_and_info" > gpurun_out/sanitize/racecheck_packed.log 2>&1
echo "racecheck packed exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize/racecheck_packed.log | tail -4
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_n_kernel_variants and (33 or 32 or 31)" > gpurun_out/sanitize/racecheck_smem.log 2>&1
echo "racecheck smem exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize/racecheck_smem.log | tail -4
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 120 --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_inverse_family.py -x -q -m gpu -k "left_side_16_byte or one_vector or (trsm_large_k and D) or (inverse_family_large_n and 100)" > gpurun_out/sanitize/racecheck_solve.log 2>&1
echo "racecheck solve exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize/racecheck_solve.log | tail -4
