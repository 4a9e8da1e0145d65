#!/bin/bash
# round 2, 2 GPUs: full GPU suite (NCCL chain identity with two list sets), bench at N=2 and N=1 (sampler block = the
# path with the collective), sampler-mode line, ncu launch list of the sampler's half-step sequence
TAG=${1:-r02_d}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2_err.log; echo "bench n2 rc=$?"; tail -3 gpurun_out/${TAG}_bench_n2_err.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1_err.log; echo "bench n1 rc=$?"
timeout 600 $TR bench.py --gpus 2 --mode sampler --steps 200 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_n2.json 2> gpurun_out/${TAG}_smode_n2_err.log; echo "smode n2 rc=$?"
timeout 600 python bench.py --mode sampler --steps 200 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_smode_n1.json 2> gpurun_out/${TAG}_smode_n1_err.log; echo "smode n1 rc=$?"
SM="python bench.py --mode sampler --steps 3 --warmup 2 --sampler-burn 60 --no-extras --no-cpu-baseline"
timeout 300 $SM > gpurun_out/${TAG}_plain_sm.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_sampler.csv $SM > gpurun_out/${TAG}_ncu_sm.log 2>&1; echo "ncu rc=$?"
python - <<P
import json
for f in ("bench_n1","bench_n2","smode_n1","smode_n2"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d["sampler"]
        print(f, "value", round(d["value"]), "| sampler", round(s["value"]), "ms/step", round(s["ms_per_step"],4), "coll", s["collectives_in_timed_region"], "B/step", s["collective_bytes_per_step"], "rebuilds", s["list_rebuilds_in_timed_region"], "reruns", s["half_steps_rerun_in_timed_region"], "queue_ms", round(s["host_queue_ms_per_step"],3), "launches", s["launches_per_step"])
        print("    ", s["lists"])
    except Exception as e: print(f, "ERR", e)
P
