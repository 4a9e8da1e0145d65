#!/bin/bash
# one full ncu capture of the fused kernel (after a plain run of the same command exited 0).  usage: bash tools/gpu_ncu.sh <tag> [kernel regex] [extra bench args]
TAG=${1:-n}; KREG=${2:-chi2_mixed}; shift 2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $@"
mkdir -p gpurun_out
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREG -s 3 -c 1 -f -o gpurun_out/${TAG}_${KREG} $BENCH > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
