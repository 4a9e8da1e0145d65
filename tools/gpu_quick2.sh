#!/bin/bash
# usage: tools/gpu_quick2.sh  -- full gpu test-suite + headline bench + config-1 + survey (each under its own timeout)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -c 1800 gpurun_out/bench_q.json
timeout 200 python bench.py --workload hc5n_dsn --walkers 128 --no-cpu-baseline 2>/dev/null | cut -c1-700
timeout 300 python tools/bench_survey.py --steps 5 2> gpurun_out/survey.err | cut -c1-300; grep -o '"host_queue[^,]*\|"gpu_launches[^,]*\|"graph[^,]*' gpurun_out/survey.err | head
