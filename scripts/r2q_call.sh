#!/bin/bash
# compute-sanitizer pass over every kernel family: memcheck, racecheck, synccheck
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_workload.py > gpurun_out/r2q_sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; tail -6 gpurun_out/r2q_sanitizer_$tool.txt
done
