#!/usr/bin/env python
"""Channel-stream path alone (cha_simulate_dev, 256 walkers x 2^20 channels = 2 GiB of spectra): best of N launches,
CUDA events on the engine stream.  usage: [CHALTE_LIB=...] [CHALTE_SPAN_STREAM=0] python tools/bench_stream.py [walkers]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder  # noqa: E402


def main():
    n_sim = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    workload = sys.argv[2] if len(sys.argv) > 2 else "benzonitrile_k1"
    prob = make_problem(workload, default_cat_folder(), n_chan=1 << 20, device=0, seed=0)
    eng = prob.engine(device=0, precision="mixed")
    stream = torch.cuda.ExternalStream(eng._lib.cha_stream(eng._h), device=torch.device("cuda", 0))
    th = torch.from_numpy(prob.walkers(n_sim, seed=1)).to("cuda:0")
    out = torch.empty((n_sim, prob.freq.size), dtype=torch.float64, device="cuda:0")
    for _ in range(3):
        eng.simulate_device(th, out=out, sync=True)
    times = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.simulate_device(th, out=out, sync=False)
        e1.record(stream)
        eng.sync(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    nbytes = n_sim * prob.freq.size * 8
    best = min(times)
    print(json.dumps({"lib": os.environ.get("CHALTE_LIB", "default"), "span_stream": os.environ.get("CHALTE_SPAN_STREAM", "1"),
                      "workload": workload, "walkers": n_sim, "bytes": nbytes, "best_ms": best, "median_ms": sorted(times)[len(times) // 2],
                      "GBps": nbytes / best / 1e6, "nonzero_frac": float((out[0] != 0).double().mean())}))


if __name__ == "__main__":
    main()
