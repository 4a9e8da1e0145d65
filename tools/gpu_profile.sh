#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list + full captures of the fused kernel and the channel-stream
# kernel (each after a plain run of the same command exited 0).  usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01_x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu1.log 2>&1
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed $BENCH > gpurun_out/${TAG}_ncu2.log 2>&1
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:simulate_tiles -s 2 -c 1 -f -o gpurun_out/${TAG}_simulate_tiles $BENCH > gpurun_out/${TAG}_ncu3.log 2>&1
echo done
