#!/bin/bash
# 2 GPUs of one box (gpurun --gpus 2): NCCL chain-identity tests, the default bench line and sampler mode under torchrun
TAG=${1:-r02_n2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/${TAG}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus 2 --mode sampler --steps 100 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_sampler.json 2> gpurun_out/${TAG}_sampler_err.log; echo "sampler rc=$?"
timeout 600 $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> /dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
python - <<P
import json
for f in ("bench","sampler"):
    d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); s=d.get("sampler") or {}
    print(f, "N", d["n_gpus"], "value", d["value"], "ms", round(d["ms_per_step"],4), "e2e", (d.get("e2e") or {}).get("value"), "| sampler", s.get("value"), s.get("ms_per_step"), "coll", s.get("collectives_in_timed_region"))
    for k in ("posterior_batch","sustained"):
        if d.get(k): print("    ", k, {a:b for a,b in d[k].items() if a in ("value","ms_per_step","seconds")})
P
grep -i "error\|Traceback" gpurun_out/${TAG}_bench_err.log gpurun_out/${TAG}_sampler_err.log | head -5
