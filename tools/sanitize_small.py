#!/usr/bin/env python
"""Small invocations of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py
Problems are small (a few thousand channels) so that the instrumented run finishes in a minute; the code paths are the
production ones: fused kernel (fast / SKIP / general variants, K = 1, 2, 4, joint), reach-ordered batch, graph replay,
channel stream (span kernel forced, and zero-fill + tiles), fp64 kernels, resident sampler with two list sets."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CHALTE_SPAN_STREAM", "2")
os.environ.setdefault("CHALTE_SORT_ROWS", "1")

from tests import test_gpu_extended as T                      # noqa: E402  (problem builders only)
from cha1_mcmc_b200 import LTEEngine                          # noqa: E402


def engine(sp, pcats, lidx, grid, prec="mixed", prior=None):
    eng = LTEEngine(device=0, precision=prec)
    eng.set_model(sp)
    for m, c in enumerate(pcats):
        eng.set_molecule(m, c, line_idx=lidx[m])
    eng.set_spectrum(*grid)
    if prior is not None:
        eng.set_prior(*prior)
    return eng


def main():
    n_big = int(os.environ.get("SANITIZE_WALKERS", "4500"))
    for mols, K, n_chan in ((["benzonitrile"], 1, 6000), (["benzonitrile"], 2, 3000), (["1-cyanonapthalene", "indene_hfs"], 4, 2048 + 130)):
        so, sp, ocats, pcats, grid, lidx, theta, stds = T._small_problem(mols, K, n_chan, seed=11)
        th = T._ball(sp, theta, stds, n_big if K == 1 else 300, seed=2, scale=0.5)
        th[::53, -1] *= 2.0                                  # reach classes differ; some rows leave the fast path
        th[7, 1] = np.nan
        with engine(sp, pcats, lidx, grid) as eng:
            ll = eng.log_like(th)
            ll2 = eng.log_like(th[:200]); ll3 = eng.log_like(th[:200])       # graph capture + replay
            assert np.array_equal(ll2, ll3) and np.array_equal(ll2, ll[:200], equal_nan=True)
            sim = eng.simulate(th[:37])
            print(mols, K, "mixed ok: finite", int(np.isfinite(ll).sum()), "of", ll.size, "sorted batches", eng.stat("sorted_batches"),
                  "sim nonzero", float((sim != 0).mean()))
        os.environ["CHALTE_SPAN_STREAM"] = "0"
        with engine(sp, pcats, lidx, grid) as eng:
            sim2 = eng.simulate(th[:37])
        os.environ["CHALTE_SPAN_STREAM"] = "2"
        assert np.array_equal(sim, sim2)
        with engine(sp, pcats, lidx, grid, prec="fp64") as eng:
            l64 = eng.log_like(th[:64]); s64 = eng.simulate(th[:3])
            m = np.isfinite(l64)
            print("   fp64 ok: max |d lnlike| mixed vs fp64", float(np.max(np.abs(l64[m] - ll[:64][m]))),
                  "sim rel", float(np.max(np.abs(s64 - sim[:3])) / np.max(np.abs(s64))))
    # resident sampler, two list sets, chain store
    g = np.load(os.path.join(ROOT, "tests", "golden", "benzonitrile_synth_ref.npz"))
    from tests import helpers as H
    _, spec = H.specs_inference(None, H.SYNTH_BOUNDS, 100, 5.8, 7000, 30000)
    cat = H.product_cat("benzonitrile")
    mu, sd = g["free/prior_means"], g["free/prior_stds"]
    rng = np.random.default_rng(7)
    p0 = mu + rng.standard_normal((1024, 5)) * sd * 0.1
    eng = H.make_engine(spec, [cat], (g["grid_freq"], g["grid_y"], g["grid_yerr"]), [g["line_idx"]], prior=(sd, mu), precision="mixed", device=0)
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler
    smp = DeviceEnsembleSampler(eng, 1024, p0, w0=0, seed=3)
    for n in (12, 12, 12):
        chain, logp = smp.run(n, store_every=4)
    print("sampler ok: chain", chain.shape, "finite", bool(np.isfinite(logp).all()), "tight builds", eng.stat("tight_builds"))
    eng.close()
    print("sanitize_small done")


if __name__ == "__main__":
    main()
