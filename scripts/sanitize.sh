#!/bin/bash
# compute-sanitizer over one small end-to-end run (benchmarks/sanitizer_target.py): memcheck and racecheck.
# The logs go to gpurun_out/ (copy the summaries to profiles/).   scripts/sanitize.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 1500 $S --tool $tool --log-file gpurun_out/sanitizer_${tag}_$tool.log --print-limit 30 \
      python benchmarks/sanitizer_target.py > gpurun_out/sanitizer_${tag}_$tool.out 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tag}_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_${tag}_$tool.log | sort | uniq -c | head -12
done
