#!/usr/bin/env python
"""Resident-sampler throughput (row N1): stretch-move steps/s and log-prob evals/s with chains in HBM, walkers
sharded over the ranks of one box (one NCCL all-gather of positions per half-step).
  python tools/bench_sampler.py [--walkers 8192] [--steps 30]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sampler.py ...
walkers = per GPU (weak scaling).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="benzonitrile_k1")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler, shard_range
    prob = make_problem(args.workload, default_cat_folder(), device=local, seed=0)
    eng = prob.engine(device=local)
    nwg = args.walkers * world
    p0 = prob.walkers(nwg, seed=11)
    w0, w1 = shard_range(nwg, world, rank)
    smp = DeviceEnsembleSampler(eng, nwg, p0[w0:w1], w0=w0, seed=5, dist=dist if world > 1 else None)
    for _ in range(args.warmup):
        smp.step()
    eng.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = eng.stat("launches")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        smp.step()
    eng.sync(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.cpu())
    _, lp, nacc = smp.state()
    if rank == 0:
        print(json.dumps({"metric": "resident-sampler log-prob evals/sec", "value": nwg * args.steps / dt, "unit": "evals/s",
                          "n_gpus": world, "steps": args.steps, "ms_per_step": 1e3 * dt / args.steps,
                          "walkers_global": nwg, "workload": args.workload,
                          "acceptance_local": nacc / (args.walkers * (args.steps + args.warmup)),
                          "launches_per_step": (eng.stat("launches") - l0) / args.steps,
                          "fused_ms_last_half_step": eng.stat("fused_ns") * 1e-6, "lists": {k: eng.stats()[k] for k in ("pairs", "active_channels", "tiles", "records", "dv_list", "hv_list")},
                          "rebuilds": eng.stat("rebuilds"), "graph_replays": eng.stat("graph_launches"), "build_ms_total": eng.stat("build_us") * 1e-3, "finite": bool(np.all(np.isfinite(lp)))}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
