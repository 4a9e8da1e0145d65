#!/bin/bash
# quick GPU check: all GPU tests (no -x) + one bench line.  usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | tail -25
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'fused_ms',r['avg_launch_ms'],'frac',r['frac'])"
tail -3 gpurun_out/${TAG}_bench_err.log
