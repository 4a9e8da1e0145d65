#!/bin/bash
# round 2, first GPU call (1 GPU): GPU tests (parity with the residual-form epilogue), smoke, bench line with the new blocks
TAG=${1:-r02_a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench_err.log
python - <<P
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "fused_ms", round(d["roofline"]["avg_launch_ms"],4))
    for k in ("sampler","posterior_batch","sustained","fp64"):
        b=d.get(k) or {}
        print(k, {kk: b.get(kk) for kk in ("value","ms_per_step","fused_ms","max_abs_dlogp_mixed_vs_fp64","seconds","list_rebuilds_in_timed_region","half_steps_rerun_in_timed_region","launches_per_step","host_queue_ms_per_step")})
    print("cpu", d.get("cpu_baseline")); print("stream", (d.get("roofline_stream") or {}).get("frac")); print("clocks", d.get("clocks"))
except Exception as e: print("ERR", e)
P
