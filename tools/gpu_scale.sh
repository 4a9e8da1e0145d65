#!/bin/bash
# Scaling evidence at N GPUs of one box (run under gpurun --gpus N): the headline line (log-prob mode + sampler block with the
# all-gather inside), sampler mode as headline, config 4 at its stated size (joint fit, K=4, 65 536 walkers sharded, strong
# scaling), config 5 (survey, strong scaling).  usage: bash tools/gpu_scale.sh <N> <tag> [tests]
N=${1:-2}; TAG=${2:-r02_scale}; TESTS=${3:-}; ONLY=${4:-headline,sampler,joint_k4,survey}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"; fi
if [ -n "$TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_n${N}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_n${N}_pytest_multi.log
fi
run() {  # name, args...
  local name=$1; shift
  case ",$ONLY," in *",$name,"*) ;; *) return;; esac
  timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}_err.log; echo "$name rc=$?"
}
run headline --steps 20 --warmup 3 --no-cpu-baseline --sustained-s 0
run sampler --mode sampler --steps 200 --warmup 20 --no-extras --no-cpu-baseline
run joint_k4 --workload joint_k4 --mode sampler --scaling strong --walkers 65536 --steps 30 --warmup 3 --sampler-burn 100 --no-extras --no-cpu-baseline
run survey --workload survey --steps 5 --warmup 3 --no-cpu-baseline
python - <<P
import json
for f in ("headline","sampler","joint_k4","survey"):
    try:
        d=json.load(open("gpurun_out/${TAG}_n${N}_%s.json" % f)); s=d.get("sampler") or {}
        print(f, "N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), d["scaling"], "| sampler", s.get("value") and round(s["value"]), s.get("ms_per_step"), "coll", s.get("collectives_in_timed_region"), "events", s.get("uncovered_events_in_timed_region"), "rebuilds", s.get("list_rebuilds_in_timed_region"), "graphs", s.get("half_steps_replayed_as_graphs"))
        if s: print("     lists", s.get("lists"))
    except Exception as e: print(f, "ERR", e)
P
for f in gpurun_out/${TAG}_n${N}_*_err.log; do grep -i "error\|Traceback" $f | head -3; done
