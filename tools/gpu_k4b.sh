#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "tmc1 or four_component or joint or random or survey or partition" 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | tail -5
for wl in benzonitrile_k4 hc7n_hfs_k4; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null > gpurun_out/k4_$wl.json
  python -c "
import json;d=json.load(open('gpurun_out/k4_$wl.json'));r=d['roofline']
print('$wl','value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'fused_ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3))"
done
