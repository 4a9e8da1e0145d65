#!/usr/bin/env python
"""BASELINE config 5: all 35 shipped catalogs x {DSN-like, GOTHAM-like} synthetic spectra = 70 fits, 262 144 walkers
split evenly across the fits (3744 each), fits sharded over the ranks of one box by pair count (no collective).
  python tools/bench_survey.py [--walkers-total 262144] [--steps 5]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_survey.py ...
Strong scaling: the survey is fixed, ranks divide it.  Prints one JSON line on rank 0."""
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
    ap.add_argument("--walkers-total", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--molecules", default="", help="comma list (default: all shipped catalogs)")
    ap.add_argument("--per-fit", action="store_true", help="also time every fit alone (rank 0 of a 1-GPU run)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cha1_mcmc_b200.synthetic import default_cat_folder
    from cha1_mcmc_b200 import survey as SV
    folder = default_cat_folder()
    mols = [m for m in args.molecules.split(",") if m] or SV.list_molecules(folder)
    t_setup = time.perf_counter()
    # every rank builds every (cheap) problem description so the cost-balanced assignment is identical everywhere
    probs = [SV.survey_problem(m, k, folder, device=local, seed=7) for m in mols for k in SV.KINDS]
    costs = [SV.fit_cost(p) for p in probs]
    mine = SV.shard_fits(costs, world)[rank]
    per_fit = args.walkers_total // len(probs)
    sv = SV.MoleculeSurvey([probs[i] for i in mine], per_fit, device=local)
    t_setup = time.perf_counter() - t_setup
    for _ in range(args.warmup):
        sv.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = sum(f.eng.stat("launches") for f in sv.fits)
    t0 = time.perf_counter()
    queue = 0.0
    for _ in range(args.steps):
        tq = time.perf_counter()
        sv.step(sync=False)
        queue += time.perf_counter() - tq
        sv.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    l1 = sum(f.eng.stat("launches") for f in sv.fits)
    n_graph = sum(f.eng.stat("graph_launches") for f in sv.fits)
    dt = torch.tensor([wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.cpu())
    detail = None
    if args.per_fit and world == 1:
        detail = {}
        for f in sv.fits:
            f.eng.log_prob_device(f.theta, out=f.out); t1 = time.perf_counter()
            for _ in range(3):
                f.eng.log_prob_device(f.theta, out=f.out, sync=False)
            f.eng.sync()
            ms = (time.perf_counter() - t1) / 3 * 1e3
            detail[f.prob.name] = {"lines": int(f.prob.line_idx[0].size), "channels": int(f.prob.freq.size),
                                   "ms": round(ms, 4), "evals_per_s": round(per_fit / ms * 1e3)}
    finite = all(bool(torch.isfinite(f.out).all()) for f in sv.fits)
    if rank == 0:
        line = {"metric": "walker log-prob evals/sec", "value": per_fit * len(probs) * args.steps / dt, "unit": "evals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "scaling": "strong", "data": "synthetic",
                "config": {"workload": f"config 5: {len(mols)} catalogs x (DSN-like 30.5 kHz 18-25 GHz K=1, GOTHAM-like 1.4 kHz "
                                       f"7-30 GHz K=4) = {len(probs)} fits, +-1.5 km/s windows around lines above 5 % of the "
                                       f"strongest, {per_fit} walkers per fit ({per_fit * len(probs)} in all)",
                           "parallelism": f"fits sharded over {world} GPU(s) by pair count, no collective",
                           "fits_on_rank0": len(mine), "channels_total": int(sum(p.freq.size for p in probs)),
                           "lines_total": int(sum(p.line_idx[0].size for p in probs))},
                "gpu_launches": int(l1 - l0), "graph_replays_total": int(n_graph), "host_queue_ms_per_step": round(1e3 * queue / args.steps, 3), "all_finite": finite, "setup_s": round(t_setup, 2)}
        if detail:
            line["per_fit"] = detail
        print(json.dumps(line))
    sv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
