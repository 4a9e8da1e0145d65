"""world_size-2 gloo tests (CPU) of the N>1 path: walker sharding + the per-half-step all-gather of the sharded
ensemble sampler, and the launch contract of bench.py's reference arm under torchrun.

The sharded driver (cha1_mcmc_b200.sampler.ShardedEnsembleSampler) is the production class; only the backend that
holds the resident local walkers is swapped: on a GPU box it is the CUDA engine, here it is the CPU restatement of the
device sampler (oracle/device_sampler_oracle.py -- the checker, used from tests only)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

WORKER = r'''
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from cha1_mcmc_b200.sampler import ShardedEnsembleSampler, broadcast_bytes, shard_range
from oracle import device_sampler_oracle as D

SCALE = np.array([1.0, 2.0, 0.5, 3.0])
def log_prob(x):
    x = np.atleast_2d(x)
    lp = -0.5 * np.sum((x / SCALE) ** 2, axis=1)
    return np.where(np.abs(x[:, 0]) < 4.0, lp, -np.inf)          # a bound: -inf lanes must survive the exchange

class OracleBackend:
    """CPU stand-in for the engine's resident sampler state (same RNG, split and arithmetic as lte_sampler.cuh)."""
    ndim = 4
    device = torch.device("cpu")
    def init(self, coords_local, nw_global, w0, seed, a):
        self.c = torch.from_numpy(coords_local.copy()); self.lp = log_prob(coords_local)
        self.w0, self.seed, self.a, self.nacc = w0, seed, a, 0
    def local_coords(self):
        return self.c
    def half_step(self, step, split, all_coords):
        c, lp, n = D.half_step(all_coords.numpy(), self.lp, self.w0, self.c.shape[0], step, split, self.seed, log_prob, self.a)
        self.c.copy_(torch.from_numpy(c)); self.lp = lp; self.nacc += n
    def get(self):
        return self.c.numpy().copy(), self.lp.copy(), self.nacc

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nw, nsteps, seed = 48, 25, 99
p0 = np.random.default_rng(3).standard_normal((nw, 4)) * SCALE
w0, w1 = shard_range(nw, world, rank)
smp = ShardedEnsembleSampler(OracleBackend(), nw, p0[w0:w1], w0=w0, seed=seed, dist=dist)
chain, logp = smp.run(nsteps)
_, _, nacc = smp.state()
# bad sharding must be refused, not silently mis-gathered
try:
    ShardedEnsembleSampler(OracleBackend(), nw, p0[w0:w1 - 1], w0=w0, seed=seed, dist=dist)
    refused = False
except ValueError:
    refused = True
# the communicator id of the in-engine exchange travels from rank 0 to every rank through this helper (128 raw bytes)
token = broadcast_bytes(dist, bytes(range(128)) if rank == 0 else None)
out = [None] * world
dist.all_gather_object(out, (w0, w1, chain, logp, nacc, refused, token))
if rank == 0:
    full = np.concatenate([o[2] for o in out], axis=0)           # (nw, nsteps, ndim)
    ref_chain, ref_lp, ref_acc = D.run(p0, log_prob, nsteps, seed=seed, shards=1)
    res = {"identical": bool(np.array_equal(np.swapaxes(full, 0, 1), ref_chain)),
           "logp_identical": bool(np.array_equal(np.concatenate([o[3][:, -1] for o in out]), ref_lp)),
           "nacc": int(sum(o[4] for o in out)), "ref_acc": int(ref_acc), "refused": all(o[5] for o in out),
           "ranges": [[o[0], o[1]] for o in out], "world": world,
           "token_ok": all(o[6] == bytes(range(128)) for o in out)}
    print("RESULT " + json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def _torchrun(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update({"REPO_ROOT": H.ROOT, "MASTER_ADDR": "127.0.0.1", "OMP_NUM_THREADS": "1"})
    env.update(env_extra or {})
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000)] + args,
                          cwd=H.ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_shard_range_tiles_the_ensemble():
    from cha1_mcmc_b200.sampler import shard_range
    for nw, world in ((8192, 1), (8192, 8), (65536, 4), (50, 4), (7, 3)):
        r = [shard_range(nw, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nw
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_sharded_sampler_world2_gloo_chain_is_identical_to_single_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    res = _torchrun([str(script)])
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, res.stdout[-2000:] + res.stderr[-2000:]
    r = json.loads(line[0][7:])
    assert r["world"] == 2 and r["ranges"] == [[0, 24], [24, 48]]
    assert r["identical"] and r["logp_identical"], "sharding changed the chain"
    assert r["nacc"] == r["ref_acc"] and r["refused"] and r["token_ok"]


def test_bench_reference_arm_under_torchrun_rank0_only():
    """bench.py --impl reference with 2 ranks: rank 0 alone prints the JSON line, rank 1 exits 0 without work."""
    res = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                     "--n-chan", "4096", "--cpu-sample", "2"], timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "evals/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"] == "walker log-prob evals/sec"
