#!/bin/bash
# compute-sanitizer over small launches of every kernel family (tests/gpu_sanitize_worker.py): memcheck, racecheck, synccheck
O=gpurun_out/r02_sanitize
mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 170 compute-sanitizer --tool $tool --error-exitcode 9 python tests/gpu_sanitize_worker.py > $O/sanitize_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_WORKER_OK|kernels ok|ns ok|layout .* ok" $O/sanitize_$tool.log | tail -6 | tee -a $O/summary.txt
done
