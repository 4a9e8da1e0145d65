"""BASELINE config 5: every shipped catalog, each with its own synthetic DSN-like (30.5 kHz channels,
18-25 GHz, 70 m dish, inference.py layout) and GOTHAM-like (1.4 kHz, 7-30 GHz, 100 m dish, 4-component TMC1 layout)
spectrum; the walkers are split evenly across the (molecule, spectrum) fits and the fits are sharded across GPUs.

The reference fits one molecule per process run (inference.py:585-640); a survey is that script started once per
molecule.  Here every fit is one engine handle (own stream, own resident tables) and a rank drives its share of the
handles back to back, so small fits (hc3n: 3 lines) overlap on the device instead of paying launch latency in turn.
Fits never exchange data: sharding is by whole fit, balanced on the (line, channel) pair count (longest first)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .catalog import MolCat, find_catalog
from .constants import ckm
from .engine import LTEEngine, ModelSpec
from .synthetic import DSN_DNU, GOTHAM_DNU, TMC1_MEANS, TMC1_STDS, SyntheticProblem

# inference.py:585-631 template (the reference's own DSN run) -- theta* of the DSN-like fits
DSN_BOUNDS = {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e16], 'Tex': [3.5, 12.0], 'vlsr': [3.0, 5.5], 'dV': [0.4, 1.5]}
DSN_TEMPLATE = dict(ss=52.0, tex=8.0, vlsr=4.3, dv=0.7575, aligned=4.10, dish=70, ll=18000, ul=25000)
KINDS = ("dsn", "gotham")


def velocity_window_grid(line_freqs, half_width_kms, dnu, v_centre):
    """Channels of a comb of spacing dnu within +-half_width_kms of each line (centre f*(1 - v_centre/ckm)): the
    windows read_file keeps (inference.py:271-273, +-1.5 km/s), merged and sorted."""
    f = np.asarray(line_freqs, dtype=float)
    if f.size == 0:
        return np.empty(0)
    c = f * (1.0 - v_centre / ckm)
    h = half_width_kms / ckm * f
    lo = np.ceil((c - h) / dnu).astype(np.int64)
    hi = np.floor((c + h) / dnu).astype(np.int64) + 1
    order = np.argsort(lo, kind="stable")
    lo, hi = lo[order], hi[order]
    end = np.maximum.accumulate(hi)
    start = np.maximum(lo, np.r_[lo[0], end[:-1]])
    n = np.maximum(end - start, 0)
    keep = n > 0
    start, n = start[keep], n[keep]
    base = np.repeat(start - np.r_[0, np.cumsum(n)[:-1]], n)
    return (base + np.arange(int(n.sum()))).astype(np.float64) * dnu


def list_molecules(cat_folder: str) -> List[str]:
    import os
    out = set()
    for fn in os.listdir(cat_folder):
        for ext in (".cat.gz", ".cat"):
            if fn.endswith(ext):
                out.add(fn[:-len(ext)])
                break
    return sorted(out)


def survey_problem(mol: str, kind: str, cat_folder: str, device: int = 0, seed: int = 0, peak_tau: float = 0.05,
                   rel_threshold: float = 0.05, noise_k: float = 0.005, half_width_kms: float = 1.5,
                   max_lines: int = 0) -> SyntheticProblem:
    """One fit of the survey.  Lines: those whose stick intensity under MolSim(C=3.4e12, dV=0.89, T=7) exceeds
    rel_threshold of the strongest (the selection rule of read_file, inference.py:266-268, 322-327).  theta*: the
    reference's templates with the column density scaled so the strongest selected line has optical depth peak_tau.
    max_lines > 0 keeps only the strongest max_lines of the selection (scaled-down test cases)."""
    path = find_catalog(cat_folder, mol)
    if path is None:
        raise FileNotFoundError(f"no catalog for {mol} in {cat_folder}")
    cat = MolCat(mol, path)
    if kind == "dsn":
        t = DSN_TEMPLATE
        spec = ModelSpec.inference(t["ss"], DSN_BOUNDS, t["dish"], t["aligned"], t["ll"], t["ul"])
        dnu, v_centre = DSN_DNU, t["vlsr"] - t["aligned"]
        ss0, tex0, dv0 = t["ss"], t["tex"], t["dv"]
    elif kind == "gotham":
        spec = ModelSpec.tmc1(4, 1)
        dnu, v_centre = GOTHAM_DNU, 5.8
        ss0, tex0, dv0 = float(TMC1_MEANS[1]), float(TMC1_MEANS[8]), float(TMC1_MEANS[13])
    else:
        raise ValueError(kind)
    with LTEEngine(device=device, precision="fp64") as eng:
        eng.set_model(spec)
        eng.set_molecule(0, cat, spec.ll, spec.ul, line_idx=None)
        f_sim, int_sim, _ = eng.stick_spectrum(0, cat.frequency.size, 3.4e12, 7.0, 0.89, ss0, spec.dish_size)
        if f_sim.size == 0:
            raise ValueError(f"{mol}: no line inside ({spec.ll}, {spec.ul}] MHz")
        sel = np.flatnonzero(int_sim > rel_threshold * int_sim.max())
        if 0 < max_lines < sel.size:
            sel = np.sort(sel[np.argsort(int_sim[sel], kind="stable")[::-1][:max_lines]])
        _, _, tau1 = eng.stick_spectrum(0, cat.frequency.size, 1.0e12, tex0, dv0, ss0, spec.dish_size)
        ncol = min(1.0e12 * peak_tau / tau1[sel].max(), 1.0e15)       # stay a decade inside the Ncol bound (1e16)
        if kind == "dsn":
            theta = np.array([ncol, tex0, t["vlsr"], dv0])
            stds = np.array([0.1 * ncol, 3.0, 0.06, 0.22])              # inference.py:601 template_stds
        else:
            share = TMC1_MEANS[4:8] / TMC1_MEANS[4:8].max()
            theta = TMC1_MEANS.copy(); stds = TMC1_STDS.copy()
            theta[4:8] = ncol * share
            stds[4:8] = TMC1_STDS[4:8] / TMC1_MEANS[4:8] * theta[4:8]
        freq = velocity_window_grid(f_sim[sel], half_width_kms, dnu, v_centre)
        eng.set_molecule(0, cat, spec.ll, spec.ul, line_idx=sel)
        eng.set_spectrum(freq, np.zeros_like(freq), np.ones_like(freq))
        truth = eng.simulate(theta[None, :])[0]
    rng = np.random.default_rng(seed)
    y = truth + rng.normal(0.0, noise_k, freq.size)
    yerr = np.sqrt(noise_k ** 2 + (0.1 * y) ** 2)
    return SyntheticProblem(f"{mol}:{kind}", spec, [cat], [sel], freq, y, yerr, theta, theta.copy(), stds)


def fit_cost(prob: SyntheticProblem) -> float:
    """Relative cost of one evaluation: (line, channel) pairs inside +-6 sigma of theta*, times K, plus the channels."""
    spec = prob.spec
    cat, sel = prob.cats[0], prob.line_idx[0]
    i0, i1 = cat.trim_bounds(spec.ll, spec.ul)
    f = cat.frequency[i0:i1][sel]
    dv = prob.theta_true[spec.idx_dv]
    dnu = np.median(np.diff(prob.freq)) if prob.freq.size > 1 else 1.0
    per_line = 2.0 * 6.0 * (dv / 2.355) / ckm * f / dnu + 1.0
    return float(spec.K * per_line.sum() + 2.0 * prob.freq.size)


def shard_fits(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of whole fits to ranks (deterministic: ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i); load[r] += costs[i]
    return [sorted(x) for x in out]


def shard_fit_walkers(costs: Sequence[float], world: int, walkers_per_fit: int, quantum: int = 128):
    """Assignment of the survey to ranks when whole fits are too coarse a unit: the walkers of one fit are
    independent, so a fit whose cost exceeds a fraction of a rank's fair share is cut into pieces of whole walker blocks
    (`quantum` walkers) that may land on different ranks; the pieces are then placed longest first like whole fits.
    Returns, per rank, a sorted list of (fit index, first walker, one-past-last walker).  Deterministic."""
    total = float(sum(costs))
    if world <= 1 or total <= 0.0:
        return [[(i, 0, walkers_per_fit) for i in range(len(costs))]] + [[] for _ in range(max(world, 1) - 1)]
    cap = total / world / 3.0                                   # no piece above a third of a rank's fair share
    blocks = max(1, -(-walkers_per_fit // quantum))
    pieces = []                                                 # (cost, fit, w0, w1)
    for i, c in enumerate(costs):
        n = min(blocks, max(1, int(-(-c // cap)))) if cap > 0 else 1
        edges = [min(walkers_per_fit, ((blocks * k) // n) * quantum) for k in range(n)] + [walkers_per_fit]
        for k in range(n):
            if edges[k + 1] > edges[k]:
                pieces.append((c * (edges[k + 1] - edges[k]) / walkers_per_fit, i, edges[k], edges[k + 1]))
    pieces.sort(key=lambda p: (-p[0], p[1], p[2]))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for c, i, a, b in pieces:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append((i, a, b)); load[r] += c
    return [sorted(x) for x in out]


@dataclass
class _Fit:
    prob: SyntheticProblem
    eng: LTEEngine
    theta: object      # torch.Tensor [nw, ndim] on the device
    out: object        # torch.Tensor [nw]


class MoleculeSurvey:
    """The fits one rank owns: engines resident, walkers resident, one `step()` = every fit's walkers evaluated once."""

    def __init__(self, problems: Sequence[SyntheticProblem], walkers_per_fit: int, device: int = 0, precision="mixed",
                 seed: int = 1, ranges=None, seeds=None):
        """ranges: per problem the (first, one-past-last) walker of the fit this rank evaluates (default: all of them);
        seeds: per problem the seed of the fit's walker ball (so that a fit cut over ranks draws one ensemble)."""
        import torch
        self.fits: List[_Fit] = []
        for k, p in enumerate(problems):
            eng = p.engine(device=device, precision=precision)
            a, b = ranges[k] if ranges is not None else (0, walkers_per_fit)
            ball = p.walkers(walkers_per_fit, seed=seeds[k] if seeds is not None else seed + k)[a:b]
            th = torch.from_numpy(np.ascontiguousarray(ball)).to(f"cuda:{device}")
            self.fits.append(_Fit(p, eng, th, torch.empty(b - a, dtype=torch.float64, device=f"cuda:{device}")))
        self._torch = torch
        self.device = device
        self._order = sorted(self.fits, key=lambda f: -fit_cost(f.prob))

    @property
    def n_evals(self) -> int:
        return sum(int(f.theta.shape[0]) for f in self.fits)

    def step(self, sync: bool = True):
        """Queue every fit on its own stream (most expensive first, so the short ones fill the tail), then wait."""
        self._torch.cuda.current_stream(self.device).synchronize()
        for f in self._order:
            f.eng.log_prob_device(f.theta, out=f.out, sync=False, wait_torch=False)
        if sync:
            self.sync()

    def sync(self):
        for f in self.fits:
            f.eng.sync()

    def log_prob(self) -> List[np.ndarray]:
        self.step()
        return [f.out.cpu().numpy() for f in self.fits]

    def close(self):
        for f in self.fits:
            f.eng.close()
        self.fits = []
        self._order = []
