#!/bin/bash
# round-end evidence: GPU tests, smoke, own arm + reference arm bench lines, config-1 line, survey line
TAG=${1:-r01_final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 300 python bench.py --workload hc5n_dsn --walkers 128 --steps 40 --no-cpu-baseline 2>/dev/null > gpurun_out/${TAG}_bench_config1.json; echo "config1 rc=$?"
timeout 300 python bench.py --workload benzonitrile_k4 --no-cpu-baseline 2>/dev/null > gpurun_out/${TAG}_bench_k4.json; echo "k4 rc=$?"
timeout 300 python tools/bench_survey.py --steps 5 2>/dev/null | tail -1 > gpurun_out/${TAG}_survey.json; echo "survey rc=$?"
timeout 200 python tools/bench_sampler.py --steps 60 --warmup 60 2>/dev/null | tail -1 > gpurun_out/${TAG}_sampler.json; echo "sampler rc=$?"
python - <<'P'
import json,sys
t=sys.argv[1] if len(sys.argv)>1 else "r01_final"
for f in ("bench","bench_config1","bench_k4","survey","sampler"):
    try:
        d=json.load(open(f"gpurun_out/r01_final_{f}.json"))
        print(f, round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_stream") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
P
