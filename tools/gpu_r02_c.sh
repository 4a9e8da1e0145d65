#!/bin/bash
# round 2, 1 GPU: full GPU suite with the two-list sampler, bench line, sampler block with the split off for comparison
TAG=${1:-r02_c}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_err.log
CHALTE_TWO_LISTS=0 timeout 900 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench_onelist.json 2> gpurun_out/${TAG}_bench_onelist_err.log; echo "bench one-list rc=$?"
timeout 900 python bench.py --no-cpu-baseline --sustained-s 0 --mode sampler --steps 100 --warmup 20 > gpurun_out/${TAG}_bench_smode.json 2> gpurun_out/${TAG}_bench_smode_err.log; echo "bench sampler-mode rc=$?"; tail -3 gpurun_out/${TAG}_bench_smode_err.log
python - <<P
import json
for f in ("bench","bench_onelist","bench_smode"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d["sampler"]
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "| sampler", round(s["value"]), "ms/step", round(s["ms_per_step"],4), "rebuilds", s["list_rebuilds_in_timed_region"], "reruns", s["half_steps_rerun_in_timed_region"], "queue_ms", round(s["host_queue_ms_per_step"],3), "fused_last", s["fused_ms_last_half_step"], "launches", s["launches_per_step"])
        print("   lists", s["lists"])
        print("   posterior", {k: (d.get("posterior_batch") or {}).get(k) for k in ("value","ms_per_step","fused_ms")})
    except Exception as e: print(f, "ERR", e)
P
