#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | tail -8
for wl in "hc5n_dsn --walkers 128 --batches 1" "hc5n_dsn --walkers 128 --batches 4 --steps 40" "benzonitrile_k1"; do
timeout 300 python bench.py --no-cpu-baseline --workload $wl 2>/dev/null > gpurun_out/e2e.json
python -c "
import json;d=json.load(open('gpurun_out/e2e.json'))
print('$wl value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches'], 'wall', round(d['wall_s_timed_region']*1e3/d['steps'],4))"
done
timeout 120 python tools/bench_sampler.py --workload hc5n_dsn --walkers 128 --steps 300 --warmup 50 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('workload','walkers_global','value','ms_per_step','graph_replays')})"
