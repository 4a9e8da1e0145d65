"""CPU restatement of the on-device stretch move (csrc/lte_sampler.cuh) -- test infrastructure only.

Same counter-based RNG (Philox4x32-10 keyed by (seed, step, global walker id)), same parity split, same
proposal / acceptance arithmetic; the log-probability is whatever callable the test passes (the oracle's).
Used to check (a) the CUDA sampler step by step and (b) that sharding walkers over ranks does not change
the chain (gloo world_size-2 test)."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays."""
    c0 = np.asarray(c0, np.uint64); c1 = np.asarray(c1, np.uint64); c2 = np.asarray(c2, np.uint64); c3 = np.asarray(c3, np.uint64)
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M32
        n1 = p1 & M32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M32
        n3 = p0 & M32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def u01(x):
    return (x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def half_step(all_coords, local_logp, w0, n_local, step, split, seed, log_prob_fn, a=2.0):
    """Update walkers [w0, w0+n_local) of colour `split`.  all_coords: (nw_global, ndim) BEFORE the half-step.
    Returns (new local coords, new local logp, n_accepted)."""
    nwg, ndim = all_coords.shape
    gid = np.arange(w0, w0 + n_local)
    mv = gid[(gid & 1) == split]
    coords = all_coords[w0:w0 + n_local].copy(); logp = local_logp.copy()
    if mv.size == 0:
        return coords, logp, 0
    r0, r1, r2, _ = philox4x32_10(np.full(mv.size, step & 0xFFFFFFFF), np.full(mv.size, (step >> 32) & 0xFFFFFFFF),
                                  mv, np.zeros(mv.size), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    zr = (a - 1.0) * u01(r0) + 1.0
    z = zr * zr / a
    nc = nwg // 2
    j = np.minimum((u01(r1) * nc).astype(np.int64), nc - 1)
    partner = 2 * j + (1 - split)
    s = all_coords[mv]; c = all_coords[partner]
    q = c - (c - s) * z[:, None]
    factor = (ndim - 1.0) * np.log(z)
    new_lp = np.asarray(log_prob_fn(q), dtype=float)
    li = mv - w0
    with np.errstate(invalid="ignore"):
        acc = (factor + new_lp - logp[li]) > np.log(u01(r2))
    coords[li[acc]] = q[acc]; logp[li[acc]] = new_lp[acc]
    return coords, logp, int(acc.sum())


def run(coords0, log_prob_fn, nsteps, seed, a=2.0, shards=1):
    """Whole-ensemble chain computed shard by shard (shards only changes the bookkeeping, never the result)."""
    nw, ndim = coords0.shape
    coords = coords0.copy(); logp = np.asarray(log_prob_fn(coords), float)
    bounds = np.linspace(0, nw, shards + 1).astype(int)
    chain = np.empty((nsteps, nw, ndim)); nacc = 0
    for step in range(nsteps):
        for split in (0, 1):
            snap = coords.copy()
            for b0, b1 in zip(bounds[:-1], bounds[1:]):
                c, lp, n = half_step(snap, logp[b0:b1], b0, b1 - b0, step, split, seed, log_prob_fn, a)
                coords[b0:b1] = c; logp[b0:b1] = lp; nacc += n
        chain[step] = coords
    return chain, logp, nacc
