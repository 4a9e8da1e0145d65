#!/bin/bash
# round 2, 2-GPU call: full GPU suite (incl. the NCCL chain-identity test), bench at N=2 (logprob + sampler blocks), config 4 shape in sampler mode
TAG=${1:-r02_b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
grep -h "\[posterior\]\|\[parity\] tmc1\|\[parity\] hc5n\|\[parity\] benz" gpurun_out/${TAG}_pytest.log | head -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2_err.log; echo "bench n2 rc=$?"; tail -3 gpurun_out/${TAG}_bench_n2_err.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustained-s 0 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1_err.log; echo "bench n1 rc=$?"
python - <<P
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d["sampler"]
        print(f, "value", round(d["value"]), "sampler", round(s["value"]), "ms/step", round(s["ms_per_step"],4), "coll", s["collectives_in_timed_region"], "bytes/step", s["collective_bytes_per_step"], "rebuilds", s["list_rebuilds_in_timed_region"], "reruns", s["half_steps_rerun_in_timed_region"], "queue_ms", round(s["host_queue_ms_per_step"],3))
    except Exception as e: print(f, "ERR", e)
P
