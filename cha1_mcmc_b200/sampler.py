"""Ensemble samplers around the vectorised log-probability.

* ``EnsembleSampler`` -- host-side affine-invariant sampler with emcee 3.1.6's interface subset the reference
  uses (``run_mcmc(pos, n)``, ``.chain``; call sites inference.py:456-473).  emcee itself is a third-party
  dependency that is not in the reference tree nor in this image; the stretch move is restated from its
  published algorithm.  log_prob_fn is called VECTORISED: (n, ndim) -> (n,).
* ``DeviceEnsembleSampler`` -- chains resident in HBM, stretch move + prior + likelihood on the device
  (cha_sampler_*), walkers sharded over ranks with one all-gather of positions per half-step, enqueued by the engine
  on its own stream (no host synchronisation per step).
* ``ShardedEnsembleSampler`` -- the same move with the exchange driven from the host through torch.distributed and a
  pluggable backend; with the CPU restatement as backend it is the gloo world-2 test of the sharding logic.
"""
from __future__ import annotations

import numpy as np


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, args=(), kwargs=None, a=2.0, pool=None, vectorize=True,
                 live_dangerously=False):
        if nwalkers < 2 * ndim and not live_dangerously:
            raise ValueError("The number of walkers needs to be more than twice the dimension of your parameter space")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn, self.args, self.kwargs = log_prob_fn, tuple(args), dict(kwargs or {})
        self.vectorize = vectorize
        # emcee seeds a private RandomState from the global NumPy state at construction
        self._random = np.random.RandomState()
        self._random.set_state(np.random.get_state())
        self._store_c = np.empty((0, self.nwalkers, self.ndim))    # capacity >= iteration; grown geometrically
        self._store_l = np.empty((0, self.nwalkers))
        self.naccepted = np.zeros(self.nwalkers, dtype=int)
        self.iteration = 0

    def reserve(self, nsteps):
        """Room for `nsteps` more stored steps (the chain is appended in place, not re-concatenated per call)."""
        need = self.iteration + int(nsteps)
        if need > self._store_c.shape[0]:
            cap = max(need, 2 * self._store_c.shape[0])
            c = np.empty((cap, self.nwalkers, self.ndim)); l = np.empty((cap, self.nwalkers))
            c[:self.iteration] = self._store_c[:self.iteration]; l[:self.iteration] = self._store_l[:self.iteration]
            self._store_c, self._store_l = c, l

    # -- emcee-compatible accessors --
    def get_chain(self):
        return self._store_c[:self.iteration]

    def get_log_prob(self):
        return self._store_l[:self.iteration]

    @property
    def _chain(self):
        return self._store_c[:self.iteration]

    @property
    def chain(self):
        """(nwalkers, nsteps, ndim): the layout the reference saves as chain.npy (inference.py:462)."""
        return np.swapaxes(self._chain, 0, 1)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iteration, 1)

    def compute_log_prob(self, coords):
        if not np.all(np.isfinite(coords)):
            raise ValueError("At least one parameter value was infinite or NaN")
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(coords, *self.args, **self.kwargs), dtype=float)
        else:
            lp = np.array([self.log_prob_fn(c, *self.args, **self.kwargs) for c in coords], dtype=float)
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def run_mcmc(self, initial_state, nsteps, log_prob0=None):
        coords = np.array(initial_state, dtype=float)
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        # a bare ndarray start means emcee recomputes log-prob of every walker
        lp = self.compute_log_prob(coords) if log_prob0 is None else np.array(log_prob0, dtype=float)
        self.reserve(nsteps)
        for s in range(nsteps):
            coords, lp, acc = self._stretch_step(coords, lp)
            self.naccepted += acc
            self._store_c[self.iteration] = coords; self._store_l[self.iteration] = lp
            self.iteration += 1
        return coords, lp

    def _stretch_step(self, coords, lp):
        nw, nd, rng, a = self.nwalkers, self.ndim, self._random, self.a
        coords = coords.copy(); lp = lp.copy()
        accepted = np.zeros(nw, dtype=bool)
        all_inds = np.arange(nw)
        inds = all_inds % 2
        rng.shuffle(inds)
        for split in range(2):
            S1 = inds == split
            s = coords[S1]; c = coords[~S1]
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.0) * rng.rand(Ns) + 1) ** 2.0 / a
            factors = (nd - 1.0) * np.log(zz)
            rint = rng.randint(Nc, size=(Ns,))
            q = c[rint] - (c[rint] - s) * zz[:, None]
            new_lp = self.compute_log_prob(q)
            lnpdiff = factors + new_lp - lp[S1]
            acc = lnpdiff > np.log(rng.rand(Ns))
            idx = all_inds[S1][acc]
            coords[idx] = q[acc]; lp[idx] = new_lp[acc]; accepted[idx] = True
        return coords, lp, accepted


def shard_range(nwalkers_global: int, world: int, rank: int):
    """Walker range [w0, w1) of `rank`: contiguous, equal shares, remainder to the first ranks."""
    base, rem = divmod(int(nwalkers_global), int(world))
    w0 = rank * base + min(rank, rem)
    return w0, w0 + base + (1 if rank < rem else 0)


class _EngineBackend:
    """The CUDA engine's resident sampler state behind the four calls the sharded driver needs."""

    def __init__(self, engine):
        import torch
        self.torch, self.eng = torch, engine
        self.ndim = engine.spec.ndim
        self.device = torch.device("cuda", engine.device)

    def init(self, coords_local, nw_global, w0, seed, a):
        self.eng.sampler_init(coords_local, nw_global=nw_global, w0=w0, seed=seed, a=a)
        pc, _ = self.eng.sampler_device_ptrs()
        n = coords_local.shape[0]

        class _Arr:           # torch view of library-owned device memory (no copy) via the CUDA array interface
            pass
        arr = _Arr()
        arr.__cuda_array_interface__ = {"shape": (n, self.ndim), "typestr": "<f8", "data": (int(pc), False),
                                        "version": 3, "strides": None}
        self._view = self.torch.as_tensor(arr, device=self.device)

    def local_coords(self):
        return self._view

    # The engine runs on its own stream; NCCL collectives are enqueued on torch's current stream.
    def _streams(self):
        if not hasattr(self, "_es"):
            self._es = self.torch.cuda.ExternalStream(self.eng._lib.cha_stream(self.eng._h), device=self.device)
        return self._es, self.torch.cuda.current_stream(self.device)

    def after_engine(self):
        """Before the positions are exchanged: the host waits for the engine AND lets it validate the half-step just
        queued (cha_sync re-runs it after a list rebuild if the lists had not covered it).  A re-run must see the
        exchanged positions of its own step, so across ranks at most one half-step is ever in flight; on a single
        rank, where no exchange happens, half-steps queue up back to back."""
        self.eng.sync()

    def before_engine(self):
        """the engine stream waits for everything queued on torch's current stream."""
        es, cur = self._streams()
        ev = self.torch.cuda.Event(); ev.record(cur); es.wait_event(ev)

    def half_step(self, step, split, all_coords):
        self.eng.sampler_half_step(step, split, all_coords.data_ptr())

    def get(self):
        return self.eng.sampler_get()


class ShardedEnsembleSampler:
    """Walkers [w0, w0+n_local) of a global ensemble live on this rank.  ``dist`` is an initialised
    torch.distributed module (or None for one rank).  Per half-step ONE all-gather of positions
    (nw_global x ndim float64) is the only exchange; the red/blue split is the parity of the global walker id
    and the RNG is Philox keyed by (seed, step, walker id), so the chain does not depend on the sharding.
    ``backend`` holds the resident local state (the CUDA engine in production)."""

    def __init__(self, backend, nwalkers_global, coords_local, w0=0, seed=0, a=2.0, dist=None):
        import torch
        self.torch = torch
        self.backend, self.dist = backend, dist
        self.nw_global = int(nwalkers_global)
        if self.nw_global % 2:
            raise ValueError("the ensemble needs an even number of walkers (two equal half-ensembles)")
        coords_local = np.ascontiguousarray(coords_local, dtype=np.float64)
        self.ndim = backend.ndim
        self.n_local, self.w0 = coords_local.shape[0], int(w0)
        self.world = dist.get_world_size() if self._distributed() else 1
        if self.world > 1:
            # all_gather_into_tensor needs equal shares; the walker ranges must tile the ensemble in rank order
            if self.n_local * self.world != self.nw_global or self.w0 != dist.get_rank() * self.n_local:
                raise ValueError("walkers must be split into equal contiguous shares in rank order (shard_range)")
        elif self.n_local != self.nw_global or self.w0 != 0:
            raise ValueError("a single rank must hold the whole ensemble")
        backend.init(coords_local, self.nw_global, self.w0, int(seed), float(a))
        self.all_coords = torch.empty((self.nw_global, self.ndim), dtype=torch.float64, device=backend.device)
        self.step_index = 0

    def _distributed(self):
        return self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1

    def step(self):
        b = self.backend
        for split in (0, 1):
            if self._distributed():
                # one all-gather of positions per half-step, ordered against the engine's stream by events
                if hasattr(b, "after_engine"):
                    b.after_engine()
                self.dist.all_gather_into_tensor(self.all_coords, b.local_coords().contiguous())
                if hasattr(b, "before_engine"):
                    b.before_engine()
                b.half_step(self.step_index, split, self.all_coords)
            else:
                # single rank: the local walkers ARE the ensemble; a half-step only rewrites walkers of its own
                # colour and reads partners of the other, so the resident array serves as all_coords in place
                b.half_step(self.step_index, split, b.local_coords())
        self.step_index += 1

    def run(self, nsteps, store_every=1):
        """Returns the local chain (n_local, nstored, ndim) and log-probs (n_local, nstored)."""
        chain, logp = [], []
        for s in range(nsteps):
            self.step()
            if (s + 1) % store_every == 0:
                c, lp, _ = self.backend.get()
                chain.append(c); logp.append(lp)
        return np.swapaxes(np.array(chain), 0, 1), np.swapaxes(np.array(logp), 0, 1)

    def state(self):
        return self.backend.get()


def broadcast_bytes(dist, payload, src=0):
    """Rank `src`'s bytes on every rank, through an initialised torch.distributed process group (any backend)."""
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class DeviceEnsembleSampler:
    """The resident sampler of the CUDA engine (cha_sampler_*): chains live in HBM; stretch move, prior and
    likelihood run on the device; a run of steps is queued on the engine's stream without host synchronisation.

    With ``dist`` spanning several ranks (one process per GPU) the walkers are sharded in equal contiguous shares
    and the engine itself enqueues the one exchange of the move -- an NCCL all-gather of positions per half-step --
    on its stream.  ``dist`` is used ONCE, to hand rank 0's communicator id to the other ranks.  The red/blue split
    is the parity of the global walker id, the RNG is Philox keyed by (seed, step, walker id) and the line/channel
    lists are sized from the proposals of the whole ensemble, so the chain does not depend on the sharding."""

    def __init__(self, engine, nwalkers_global, coords_local, w0=0, seed=0, a=2.0, dist=None):
        self.eng = engine
        self.nw_global = int(nwalkers_global)
        if self.nw_global % 2:
            raise ValueError("the ensemble needs an even number of walkers (two equal half-ensembles)")
        coords_local = np.ascontiguousarray(coords_local, dtype=np.float64)
        self.ndim = engine.spec.ndim
        self.n_local, self.w0 = coords_local.shape[0], int(w0)
        self.world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 1:
            if self.n_local * self.world != self.nw_global or self.w0 != self.rank * self.n_local:
                raise ValueError("walkers must be split into equal contiguous shares in rank order (shard_range)")
            comm_id = broadcast_bytes(dist, engine.comm_unique_id() if self.rank == 0 else None)
            engine.comm_init(self.rank, self.world, comm_id)
        elif self.n_local != self.nw_global or self.w0 != 0:
            raise ValueError("a single rank must hold the whole ensemble")
        else:
            engine.comm_init(0, 1, bytes(engine.COMM_ID_BYTES))
        engine.sampler_init(coords_local, nw_global=self.nw_global, w0=self.w0, seed=int(seed), a=float(a))
        self.step_index = 0

    def step(self, n=1):
        """Queue n steps (no host synchronisation; ``sync`` or any read waits)."""
        self.eng.sampler_run(self.step_index, n, 0)
        self.step_index += n

    def sync(self):
        self.eng.sync()

    def run(self, nsteps, store_every=1):
        """Returns the local chain (n_local, nstored, ndim) and log-probs (n_local, nstored)."""
        first = self.eng.sampler_chain_len()
        self.eng.sampler_run(self.step_index, nsteps, store_every)
        self.step_index += nsteps
        c, lp = self.eng.sampler_chain_read(first)
        self.eng.sampler_chain_clear()
        return np.ascontiguousarray(np.swapaxes(c, 0, 1)), np.ascontiguousarray(np.swapaxes(lp, 0, 1))

    def state(self):
        return self.eng.sampler_get()
