"""Multi-GPU tests (run only when >= 2 CUDA devices are visible): the resident sampler with its walkers sharded over
2 / 4 / 8 ranks -- the all-gather of positions enqueued by the engine on its own stream (cha_comm_init,
cha_sampler_run), no host synchronisation per step -- gives the same chain as one GPU, bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from tests import helpers as H
from cha1_mcmc_b200.sampler import DeviceEnsembleSampler, shard_range

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(H.GOLD + "/benzonitrile_synth_ref.npz")
_, spec = H.specs_inference(None, H.SYNTH_BOUNDS, 100, 5.8, 7000, 30000)
cat = H.product_cat("benzonitrile")
grid = (g["grid_freq"], g["grid_y"], g["grid_yerr"])
mu, sd = g["free/prior_means"], g["free/prior_stds"]
nw, nsteps = 4096, 100
rng = np.random.default_rng(7)
p0 = mu + rng.standard_normal((nw, 5)) * sd * 0.1
eng = H.make_engine(spec, [cat], grid, [g["line_idx"]], prior=(sd, mu), precision="mixed", device=local)
w0, w1 = shard_range(nw, world, rank)
smp = DeviceEnsembleSampler(eng, nw, p0[w0:w1], w0=w0, seed=2024, dist=dist)
# several runs: the step counter and the chain store carry over; every run ends at a synchronisation point, where the
# narrow list set of the bulk / outlier split is chosen from the class histogram of the WHOLE ensemble
parts = [smp.run(n, store_every=5) for n in (40, 40, nsteps - 80)]
chain = np.concatenate([p[0] for p in parts], axis=1); logp = np.concatenate([p[1] for p in parts], axis=1)
ncoll = eng.stat("collectives"); reruns = eng.stat("reruns"); tight = eng.stat("tight_builds")
out = [None] * world
dist.all_gather_object(out, (chain, logp, smp.state()[2], ncoll, reruns, tight))
if rank == 0:
    full = np.concatenate([o[0] for o in out], axis=0)
    lp = np.concatenate([o[1] for o in out], axis=0)
    np.save(os.environ["OUT_PREFIX"] + f"_chain_w{world}.npy", full)
    np.save(os.environ["OUT_PREFIX"] + f"_logp_w{world}.npy", lp)
    print("RESULT " + json.dumps({"world": world, "nacc": int(sum(o[2] for o in out)), "collectives": [int(o[3]) for o in out],
                                  "reruns": [int(o[4]) for o in out], "tight_builds": [int(o[5]) for o in out]}), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def _run(world, script, prefix):
    env = dict(os.environ)
    env.update({"REPO_ROOT": H.ROOT, "MASTER_ADDR": "127.0.0.1", "OUT_PREFIX": prefix})
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), script],
                         cwd=H.ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, res.stdout[-2000:]
    return json.loads(line[0][7:])


def test_nccl_sharded_sampler_chain_identical_to_single_gpu(tmp_path):
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    prefix = str(tmp_path / "r")
    r1 = _run(1, str(script), prefix)
    c1, l1 = np.load(prefix + "_chain_w1.npy"), np.load(prefix + "_logp_w1.npy")
    assert c1.shape == (4096, 20, 5)
    assert r1["collectives"] == [0] and 0.1 < r1["nacc"] / (4096 * 100) < 0.95
    assert r1["tight_builds"][0] >= 1, "the bulk / outlier split never came into use"
    for world in [w for w in (2, 4, 8) if w <= ndev]:
        rw = _run(world, str(script), prefix)
        cw, lw = np.load(prefix + f"_chain_w{world}.npy"), np.load(prefix + f"_logp_w{world}.npy")
        assert np.array_equal(c1, cw) and np.array_equal(l1, lw), f"sharding over {world} GPUs changed the chain"
        assert rw["nacc"] == r1["nacc"]
        # one all-gather per half-step on every rank (+ the re-runs after a list rebuild, the same on every rank)
        assert len(set(rw["collectives"])) == 1 and len(set(rw["reruns"])) == 1
        assert rw["collectives"][0] >= 2 * 100 and rw["tight_builds"] == r1["tight_builds"] * world
