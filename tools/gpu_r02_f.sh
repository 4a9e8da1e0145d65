#!/bin/bash
# round 2, 1 GPU: full GPU suite (data rounded to fp32 after sigma scaling), bench, A/B against the hi+lo split build,
# sampler mode over 300 steps with sync-point logging
TAG=${1:-r02_f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
CHALTE_LIB=$PWD/cha1_mcmc_b200/csrc/variants/libchalte_split.so timeout 600 python bench.py --no-extras > gpurun_out/${TAG}_bench_split.json 2> gpurun_out/${TAG}_bench_split_err.log; echo "bench split rc=$?"
CHALTE_DEBUG=1 timeout 600 python bench.py --mode sampler --steps 300 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_n1.json 2> gpurun_out/${TAG}_smode_n1_err.log; echo "smode n1 rc=$?"
grep "\[chalte\]" gpurun_out/${TAG}_smode_n1_err.log | tail -45
python - <<P
import json
for f in ("bench","bench_split","smode_n1"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["avg_launch_ms"],4), "e2e", round(d["e2e"]["value"]), "cpu", (d.get("cpu_baseline") or {}).get("max_abs_dlogp_vs_gpu"), "| sampler", s.get("value"), s.get("ms_per_step"))
        print("    ", {k: s[k] for k in s if "timed_region" in k})
    except Exception as e: print(f, "ERR", e)
P
