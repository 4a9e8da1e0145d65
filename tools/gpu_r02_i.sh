#!/bin/bash
# A/B: line strengths through MUFU.EX2 in log2 space (variant build) against the fp64 polynomial (default build)
TAG=${1:-r02_i}
mkdir -p gpurun_out
V=$PWD/cha1_mcmc_b200/csrc/variants/libchalte_mufu.so
rm -f gpurun_out/parity_errors.json
CHALTE_LIB=$V timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extended.py -m gpu -q > gpurun_out/${TAG}_pytest_mufu.log 2>&1; echo "pytest(mufu) rc=$?"; tail -5 gpurun_out/${TAG}_pytest_mufu.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_mufu.json
CHALTE_LIB=$V timeout 600 python bench.py --no-extras > gpurun_out/${TAG}_bench_mufu.json 2> gpurun_out/${TAG}_bench_mufu_err.log; echo "bench(mufu) rc=$?"
timeout 600 python bench.py --no-extras > gpurun_out/${TAG}_bench_def.json 2> gpurun_out/${TAG}_bench_def_err.log; echo "bench(default) rc=$?"
CHALTE_LIB=$V timeout 600 python bench.py --mode sampler --steps 200 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_mufu.json 2>/dev/null; echo "smode(mufu) rc=$?"
python - <<P
import json
for f in ("bench_mufu","bench_def","smode_mufu"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["avg_launch_ms"],4), "cpu_err", (d.get("cpu_baseline") or {}).get("max_abs_dlogp_vs_gpu"))
    except Exception as e: print(f, "ERR", e)
d=json.load(open("gpurun_out/${TAG}_parity_mufu.json"))
for k,v in sorted(d.items()):
    if "mixed" in k and ("tmc1" in k or "hc5n" in k or "benz" in k): print(k, v["max_abs_err_good_fit"], v["max_abs_err_far"])
P
