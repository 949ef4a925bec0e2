#!/bin/bash
# compute-sanitizer over one small launch of every hand-written kernel family (run on the GPU box).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --report-api-errors no --print-limit 20 python tools/sanitizer_targets.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.txt | tail -1)"
done
