#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (one GPU).  usage: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r02_sanitize}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_small.py > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"; tail -12 gpurun_out/${TAG}_plain.log
for tool in memcheck racecheck synccheck; do
  SANITIZE_WALKERS=4200 timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/${TAG}_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done|Error|hazard" gpurun_out/${TAG}_$tool.log | sort | uniq -c | head -8
done
