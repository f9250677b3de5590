#!/bin/bash
# compute-sanitizer passes over a small, kernel-covering subset of the GPU tests (1 GPU; slow: ~20-50x).
#   bash tools/gpu_sanitize.sh            memcheck + racecheck (+ synccheck on the kernels that use mbarriers / named syncs)
# Logs land in gpurun_out/sanitize_*.log; the summary lines are echoed.
mkdir -p gpurun_out
SEL='goldens or api_forms or (random_mxv and 9) or (random_mxm and 9) or complemented_null_mask or (masked_pull_hub_rows_against_oracle and 1) or (masked_push_forced_against_oracle and 1) or test_ref_ or rmat_triangle_count_masked_mxm or (large_spmv_all_kernel_variants and mode0) or large_spmv_sparse_u or (assign_scalar_against_model and 5) or (matrix_extract_against_scipy and 2) or (hypersparse_operands and 1)'
for tool in memcheck racecheck synccheck; do
  K="$SEL"
  timeout ${SAN_TIMEOUT:-700} compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 \
      python -m pytest tests/test_parity_gpu.py tests/test_matrix_ops_gpu.py tests/test_vector_ops_gpu.py -v -m gpu -x -p no:cacheprovider -k "$K" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY| passed| failed|Error:|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | tail -8; grep -c PASSED gpurun_out/sanitize_$tool.log
done
