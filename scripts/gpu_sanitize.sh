#!/bin/bash
# compute-sanitizer over a small exercise of every kernel family (scripts/sanitize_small.py)
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck}; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc $?"; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard\|Invalid\|rms\|solve\|drt_cross" gpurun_out/sanitize_$tool.log | head -14
done
