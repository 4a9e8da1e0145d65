#!/bin/bash
# round 2, 1 GPU: reach-ordered plain log-prob batches (A/B), span kernel buffer variants
TAG=${1:-r02_n}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity.json 2>/dev/null
V=$PWD/cha1_mcmc_b200/csrc/variants
{
timeout 300 python tools/bench_stream.py
for v in r4b2 r8b2; do CHALTE_LIB=$V/libchalte_$v.so timeout 300 python tools/bench_stream.py; done
CHALTE_LIB=$V/libchalte_r4b2.so timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
} > gpurun_out/${TAG}_stream.jsonl 2> gpurun_out/${TAG}_stream_err.log
cat gpurun_out/${TAG}_stream.jsonl; tail -5 gpurun_out/${TAG}_stream_err.log
timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
CHALTE_SORT_ROWS=0 timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench_nosort.json 2> gpurun_out/${TAG}_bench_nosort_err.log; echo "bench nosort rc=$?"
CHALTE_SORT_ROWS=1 timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 --no-extras > gpurun_out/${TAG}_bench_sort1.json 2> gpurun_out/${TAG}_bench_sort1_err.log; echo "bench sort1 rc=$?"
python - <<P
import json
for f in ("bench","bench_nosort","bench_sort1"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}; r=d.get("roofline") or {}
        print(f, "value", d["value"], "ms", round(d["ms_per_step"],4), "fused", r.get("avg_launch_ms"), "e2e", (d.get("e2e") or {}).get("value"), "| sampler", s.get("value"), s.get("ms_per_step"))
        for k in ("posterior_batch","roofline_stream"):
            if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","frac","ms","fused_ms","ms_per_step_median_rank0","ms_per_step_max_rank0","list_rebuilds_incl_warmup","reach_ordered_batches")})
    except Exception as e: print(f, "ERR", e)
P
