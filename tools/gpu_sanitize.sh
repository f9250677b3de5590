#!/bin/bash
# compute-sanitizer passes over a small, kernel-covering subset of the GPU tests (1 GPU; slow: ~20-50x).
#   bash tools/gpu_sanitize.sh            memcheck + racecheck
# Logs land in gpurun_out/sanitize_*.log; the summary lines are echoed.
mkdir -p gpurun_out
SEL='goldens or api_forms or (random_mxv and (0- or 9- or 19-)) or (random_mxm and (0- or 7-)) or masked_pull_hub_rows_against_oracle[0] or masked_push_forced_against_oracle[1] or test_ref_'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 \
      python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
