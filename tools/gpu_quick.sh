#!/bin/bash
# quick check on one GPU: full GPU suite + one default bench line.  usage: bash tools/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity.json 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); s=d.get("sampler") or {}; r=d.get("roofline") or {}
print("value", d["value"], "ms", round(d["ms_per_step"],4), "fused", r.get("avg_launch_ms"), "e2e", (d.get("e2e") or {}).get("value"), "| sampler", s.get("value"), s.get("ms_per_step"), "build_ms", s.get("list_build_ms_in_timed_region"))
for k in ("posterior_batch","roofline_stream"):
    if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","frac","ms","fused_ms","ms_per_step_median_rank0","list_rebuilds_incl_warmup","reach_ordered_batches")})
P
