#!/bin/bash
# round 2, 1 GPU: one-pass channel-stream kernel (spans + TMA bulk stores) -- parity, A/B against memset + tiles, span
# sizes; fused kernel with MUFU strengths + integer group-sum conversion as default
TAG=${1:-r02_k}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity.json 2>/dev/null
V=$PWD/cha1_mcmc_b200/csrc/variants
{
timeout 300 python tools/bench_stream.py
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py
for v in span256 span1024 span512x8 span256x8; do CHALTE_LIB=$V/libchalte_$v.so timeout 300 python tools/bench_stream.py; done
timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
timeout 300 python tools/bench_stream.py 100
} > gpurun_out/${TAG}_stream.jsonl 2> gpurun_out/${TAG}_stream_err.log
cat gpurun_out/${TAG}_stream.jsonl; tail -5 gpurun_out/${TAG}_stream_err.log
timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
python - <<P
import json
for f in ("bench",):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}; r=d.get("roofline") or {}
        print(f, "value", d["value"], "ms", round(d["ms_per_step"],4), "fused", r.get("avg_launch_ms"), "e2e", (d.get("e2e") or {}).get("value"), "| sampler", s.get("value"), s.get("ms_per_step"))
        for k in ("posterior_batch","roofline_stream"):
            if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","frac","ms","seconds","clocks","fused_ms")})
    except Exception as e: print(f, "ERR", e)
P
