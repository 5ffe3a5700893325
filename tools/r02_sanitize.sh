#!/bin/bash
# compute-sanitizer passes of round 2 (run under gpurun; logs in gpurun_out/)
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_driver.py > gpurun_out/r02k_sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/r02k_sanitizer_$tool.log
done
tail -5 gpurun_out/r02k_sanitizer_*.log
