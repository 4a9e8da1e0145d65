#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | tail -8
for wl in benzonitrile_k1 benzonitrile_k4; do
timeout 300 python bench.py --no-cpu-baseline --workload $wl 2>/dev/null > gpurun_out/sim_$wl.json
python -c "
import json;d=json.load(open('gpurun_out/sim_$wl.json'));r=d['roofline_stream']
print('$wl value',round(d['value']),'stream frac',round(r['frac'],3),'ms',round(r['ms'],4),'GB/s',round(r['achieved']))"
done
