#!/bin/bash
# GPU box, one GPU: compute-sanitizer memcheck / racecheck / synccheck over scripts/sanitize_cases.py; logs under
# gpurun_out/ (summaries are copied to profiles/ by hand).  Usage: scripts/gpu_sanitize.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_cases.py > gpurun_out/${tag}_sanitize_plain.log 2>&1; echo "plain rc=$?"
tail -2 gpurun_out/${tag}_sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 50 python scripts/sanitize_cases.py > gpurun_out/${tag}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases: ok" gpurun_out/${tag}_sanitize_${tool}.log | tail -3
done
