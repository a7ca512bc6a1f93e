#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck initcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_k0_k5.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/san_$tool.log
  tail -6 gpurun_out/san_$tool.log
done
