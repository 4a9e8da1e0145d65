#!/bin/bash
# round 2, 1 GPU: span kernel staging only the span's segment of a tile
TAG=${1:-r02_o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extended.py -m gpu -q -k "one_pass or boundary or full_size_four or joint_two or dsn_like" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
{
timeout 300 python tools/bench_stream.py
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py
timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
timeout 300 python tools/bench_stream.py 256 joint_k4
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py 256 joint_k4
} > gpurun_out/${TAG}_stream.jsonl 2> gpurun_out/${TAG}_stream_err.log
cat gpurun_out/${TAG}_stream.jsonl; tail -5 gpurun_out/${TAG}_stream_err.log
BENCH="python tools/bench_stream.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simulate_span -s 3 -c 1 -f -o gpurun_out/${TAG}_simulate_span $BENCH > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu span rc=$?"
