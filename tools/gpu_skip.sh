#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | tail -12
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/skip_k1.json
python -c "
import json;d=json.load(open('gpurun_out/skip_k1.json'));r=d['roofline']
print('k1 value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'fused_ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3))"
timeout 200 python tools/bench_sampler.py --steps 60 --warmup 60 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('workload','walkers_global','value','ms_per_step','launches_per_step','rebuilds','graph_replays','acceptance_local','fused_ms_last_half_step')}); print(d['lists'])"
