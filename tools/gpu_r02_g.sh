#!/bin/bash
# round 2, 1 GPU: sampler half-steps replayed as CUDA graphs (A/B), full GPU suite
TAG=${1:-r02_g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --mode sampler --steps 300 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_graph.json 2> gpurun_out/${TAG}_smode_graph_err.log; echo "smode graph rc=$?"
CHALTE_SAMPLER_GRAPHS=0 timeout 600 python bench.py --mode sampler --steps 300 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_plain.json 2> gpurun_out/${TAG}_smode_plain_err.log; echo "smode plain rc=$?"
timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
python - <<P
import json
for f in ("smode_graph","smode_plain","bench"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "| sampler", round(s.get("value")), s.get("ms_per_step"), "graphs", s.get("half_steps_replayed_as_graphs"), "queue", s.get("host_queue_ms_per_step"))
        print("    ", {k: s[k] for k in s if "timed_region" in k})
    except Exception as e: print(f, "ERR", e)
P
