#!/bin/bash
# round 2, 1 GPU: sampler mode over 300 steps with event logging, default bench, full ncu capture of the fused kernel
TAG=${1:-r02_e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extended.py -m gpu -q -x -k "sampler or odd_rows or full_size" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
CHALTE_DEBUG=1 timeout 600 python bench.py --mode sampler --steps 300 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_n1.json 2> gpurun_out/${TAG}_smode_n1_err.log; echo "smode n1 rc=$?"
grep "\[chalte\]" gpurun_out/${TAG}_smode_n1_err.log | tail -40
timeout 600 python bench.py --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_k1 $BENCH > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu rc=$?"
python - <<P
import json
for f in ("smode_n1","bench"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d["sampler"]
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["avg_launch_ms"],4), "| sampler", round(s["value"]), "ms/step", round(s["ms_per_step"],4))
        print("    ", {k: s[k] for k in s if "timed_region" in k}, s["lists"])
    except Exception as e: print(f, "ERR", e)
P
