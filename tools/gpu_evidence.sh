#!/bin/bash
# One gpurun call (1 GPU) that produces the round's evidence: full GPU suite, smoke, both bench arms, the other
# workloads, channel-stream A/B, the ncu launch lists and the full ncu captures (fused kernel: headline / sampler
# half-step / joint K=4; channel-stream span kernel) -- every ncu pass after a plain run of the same command exited 0.
# Read the captures on the CPU box: tools/ncu_summary.py, tools/ncu_stalls.py, tools/ncu_metrics_json.py.
# usage: bash tools/gpu_evidence.sh <tag>
TAG=${1:-r02_final}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_errors.json gpurun_out/${TAG}_parity.json 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 600 python bench.py --mode sampler --steps 200 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_sampler.json 2>/dev/null; echo "sampler rc=$?"
timeout 600 python bench.py --workload hc5n_dsn --walkers 128 --steps 200 --warmup 20 --no-extras > gpurun_out/${TAG}_bench_config1.json 2>/dev/null; echo "config1 rc=$?"
timeout 600 python bench.py --workload benzonitrile_k4 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_k4.json 2>/dev/null; echo "k4 rc=$?"
timeout 600 python bench.py --workload joint_k4 --walkers 8192 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_joint.json 2>/dev/null; echo "joint rc=$?"
timeout 600 python bench.py --workload survey --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_survey.json 2>/dev/null; echo "survey rc=$?"
{
timeout 300 python tools/bench_stream.py
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py
timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
CHALTE_SPAN_STREAM=0 timeout 300 python tools/bench_stream.py 256 benzonitrile_k4
} > gpurun_out/${TAG}_stream.jsonl 2> gpurun_out/${TAG}_stream_err.log; cat gpurun_out/${TAG}_stream.jsonl
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-s 0"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_k1 $BENCH > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu k1 rc=$?"
# sampler half-step at steady state: skip the half-steps of the burn-in (2 fused launches per step)
SB="python bench.py --mode sampler --steps 6 --warmup 2 --sampler-burn 60 --no-extras --no-cpu-baseline"
timeout 300 $SB > gpurun_out/${TAG}_plain_s.log 2>&1 &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 132 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_sampler $SB > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu sampler rc=$?"
timeout 300 $SB > gpurun_out/${TAG}_plain_s.log 2>&1 &&
timeout 600 $NCU --metrics gpu__time_duration.sum -s 900 -c 60 --csv --log-file gpurun_out/${TAG}_launches_sampler.csv $SB > gpurun_out/${TAG}_ncu3b.log 2>&1; echo "launch list sampler rc=$?"
JB="python bench.py --workload joint_k4 --walkers 8192 --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 300 $JB > /dev/null 2> gpurun_out/${TAG}_plain_j.log &&
timeout 600 $NCU --set full --import-source on -k regex:chi2_mixed -s 3 -c 1 -f -o gpurun_out/${TAG}_chi2_mixed_joint $JB > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu joint rc=$?"
ST="python tools/bench_stream.py"
timeout 300 $ST > gpurun_out/${TAG}_plain_st.log 2>&1 &&
timeout 600 $NCU --set full --import-source on -k regex:"simulate_span|sim_gcoef|sim_line_tau" -s 9 -c 3 -f -o gpurun_out/${TAG}_simulate_span $ST > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu stream rc=$?"
# the captures are read here (only text travels back: four .ncu-rep files exceed the 64 MiB that gpurun copies)
for k in chi2_mixed_k1 chi2_mixed_sampler chi2_mixed_joint simulate_span; do
  R=gpurun_out/${TAG}_$k.ncu-rep
  [ -f $R ] || continue
  python tools/ncu_summary.py $R > gpurun_out/${TAG}_${k}_ncu_summary.txt
  ncu -i $R --page source --csv > /tmp/${k}.src.csv 2>/dev/null
  { echo "-- stall samples by reason and by opcode (ncu --page source)"; python tools/ncu_stalls.py /tmp/${k}.src.csv;
    echo "-- executed warp instructions by opcode"; python tools/ncu_source_hist.py /tmp/${k}.src.csv; } >> gpurun_out/${TAG}_${k}_ncu_summary.txt 2>&1
done
python tools/ncu_metrics_json.py gpurun_out/${TAG}_ncu_metrics.json chi2_mixed_kernel=gpurun_out/${TAG}_chi2_mixed_k1.ncu-rep \
  chi2_mixed_kernel_sampler_half_step=gpurun_out/${TAG}_chi2_mixed_sampler.ncu-rep chi2_mixed_kernel_joint_k4=gpurun_out/${TAG}_chi2_mixed_joint.ncu-rep \
  channel_stream=gpurun_out/${TAG}_simulate_span.ncu-rep > /dev/null 2>&1; echo "metrics rc=$?"
rm -f gpurun_out/${TAG}_chi2_mixed_sampler.ncu-rep gpurun_out/${TAG}_chi2_mixed_joint.ncu-rep gpurun_out/${TAG}_simulate_span.ncu-rep
python - <<P
import json
for f in ("bench","bench_reference_arm","sampler","bench_config1","bench_k4","bench_joint","survey"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}; r=d.get("roofline") or {}
        print(f, "value", d["value"], "ms", round(d["ms_per_step"],4), "fused", r.get("avg_launch_ms"), "frac", r.get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "| sampler", s.get("value"), s.get("ms_per_step"))
        for k in ("sustained","posterior_batch","fp64","roofline_stream"):
            if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","frac","ms","seconds","fused_ms","traffic")})
    except Exception as e: print(f, "ERR", e)
P
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
