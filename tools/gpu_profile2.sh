#!/bin/bash
# One gpurun call: GPU tests, smoke, bench lines (own arm + reference arm), ncu launch list + full captures of the
# fused kernel at K=1 and K=4 (each after a plain run of the same command exited 0).  usage: bash tools/gpu_profile2.sh <tag>
set -u
TAG=${1:-r01_x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_k1 $BENCH > gpurun_out/${TAG}_ncu2.log 2>&1
BENCH4="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload benzonitrile_k4"
timeout 300 $BENCH4 > gpurun_out/${TAG}_plain4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_k4 $BENCH4 > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
echo done
