#!/bin/bash
# round 2, 1 GPU: full GPU suite, smoke, both bench arms, A/B of the MUFU-strengths variant, ncu launch list and full
# captures (fused kernel K=1 headline; sampler half-step; joint M=2 K=4; channel stream), each after a plain run of
# the same command exited 0.   usage: bash tools/gpu_r02_j.sh <tag>
TAG=${1:-r02_j}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity.json 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo "ref rc=$?"
V=$PWD/cha1_mcmc_b200/csrc/variants/libchalte_mufu.so
if [ -f $V ]; then
  CHALTE_LIB=$V timeout 600 python bench.py --no-extras > gpurun_out/${TAG}_bench_mufu.json 2> gpurun_out/${TAG}_bench_mufu_err.log; echo "bench(mufu) rc=$?"
  CHALTE_LIB=$V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_pytest_mufu.log 2>&1; echo "pytest(mufu) rc=$?"; tail -2 gpurun_out/${TAG}_pytest_mufu.log
  cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity_mufu.json 2>/dev/null
fi
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-s 0"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_k1 $BENCH > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu k1 rc=$?"
# sampler half-step at steady state: skip the half-steps of the burn-in (2 fused launches per step)
SB="python bench.py --mode sampler --steps 6 --warmup 2 --sampler-burn 60 --no-extras --no-cpu-baseline"
timeout 300 $SB > gpurun_out/${TAG}_plain_s.log 2>&1 &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 132 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_sampler $SB > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu sampler rc=$?"
timeout 300 $SB > gpurun_out/${TAG}_plain_s.log 2>&1 &&
timeout 600 $NCU --metrics gpu__time_duration.sum -s 700 -c 120 --csv --log-file gpurun_out/${TAG}_launches_sampler.csv $SB > gpurun_out/${TAG}_ncu3b.log 2>&1; echo "launch list sampler rc=$?"
JB="python bench.py --workload joint_k4 --walkers 8192 --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 300 $JB > gpurun_out/${TAG}_joint.json 2> gpurun_out/${TAG}_plain_j.log &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_joint $JB > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu joint rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-s 0 --sampler-steps 4"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 $NCU --set full -k regex:simulate_tiles -s 2 -c 1 -f -o gpurun_out/${TAG}_simulate_tiles $BENCH > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu stream rc=$?"
python - <<P
import json
for f in ("bench","bench_ref","bench_mufu","joint"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}; r=d.get("roofline") or {}
        print(f, "value", d["value"], "ms", round(d["ms_per_step"],4), "fused", r.get("avg_launch_ms"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", d.get("cpu_baseline"), "| sampler", s.get("value"), s.get("ms_per_step"))
        for k in ("sustained","posterior_batch","fp64","roofline_stream"):
            if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","frac","ms","seconds","clocks","fused_ms")})
    except Exception as e: print(f, "ERR", e)
P
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
