"""Seeded synthetic spectra of the shapes BASELINE.json names (there is no network for real GOTHAM data).

SURVEY.md 8(d) config 3: GOTHAM-like grid = union of +-W-channel windows at 1.4 kHz around every selected
line, W chosen so the grid has exactly ``n_chan`` channels; y = model(theta*) + N(0, 5 mK);
sigma_j = sqrt((5 mK)^2 + (0.1 y_j)^2) (mirrors inference.py:290); walkers = theta* + randn * s/10 redrawn
until inside the bounds (mirrors inference.py:442-451)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .catalog import MolCat, find_catalog
from .constants import ckm
from .engine import LTEEngine, ModelSpec

GOTHAM_DNU = 1.4e-3      # MHz
DSN_DNU = 30.518e-3      # MHz (data/DSN sample)

# TMC1_four_component.py:292-294 template (HC9N) -- theta* of the 4-component configs
TMC1_MEANS = np.array([37, 25, 56, 22, 2.47e12, 11.19e12, 2.20e12, 5.64e12, 6.7, 5.624, 5.790, 5.910, 6.033, 0.117])
TMC1_STDS = np.array([2.5, 2.0, 6.5, 2.0, 0.30e12, 1.75e12, 0.265e12, 1.185e12, 0.1, 0.0015, 0.001, 0.0035, 0.002, 0.002])


def window_grid(line_freqs: np.ndarray, n_chan: int, dnu: float = GOTHAM_DNU, v_centre: float = 0.0) -> np.ndarray:
    """Sorted channel frequencies: union of +-W channel windows (comb of spacing dnu) around
    f_i*(1 - v_centre/ckm), with W the smallest half-width giving >= n_chan channels, truncated to
    exactly n_chan from the high-frequency end."""
    centres = np.round(np.asarray(line_freqs, float) * (1.0 - v_centre / ckm) / dnu).astype(np.int64)
    centres = np.unique(centres)

    def count(W):
        lo = centres - W; hi = centres + W + 1
        # merged length of sorted intervals
        end = np.maximum.accumulate(hi)
        start = np.maximum(lo, np.r_[lo[0], end[:-1]])
        return int(np.sum(np.maximum(hi - start, 0)))

    lo_w, hi_w = 0, 1
    while count(hi_w) < n_chan:
        hi_w *= 2
        if hi_w > (1 << 26):
            raise ValueError("cannot reach the requested channel count")
    while lo_w < hi_w:
        mid = (lo_w + hi_w) // 2
        if count(mid) >= n_chan:
            hi_w = mid
        else:
            lo_w = mid + 1
    W = lo_w
    idx = np.unique((centres[:, None] + np.arange(-W, W + 1)[None, :]).ravel())
    idx = idx[:n_chan]
    return idx.astype(np.float64) * dnu


@dataclass
class SyntheticProblem:
    name: str
    spec: ModelSpec
    cats: List[MolCat]
    line_idx: List[Optional[np.ndarray]]
    freq: np.ndarray
    y: np.ndarray
    yerr: np.ndarray
    theta_true: np.ndarray
    prior_means: np.ndarray
    prior_stds: np.ndarray

    def engine(self, device=0, precision="mixed") -> LTEEngine:
        eng = LTEEngine(device=device, precision=precision)
        eng.set_model(self.spec)
        for m, c in enumerate(self.cats):
            eng.set_molecule(m, c, line_idx=self.line_idx[m])
        eng.set_spectrum(self.freq, self.y, self.yerr)
        eng.set_prior(self.prior_stds, self.prior_means)
        return eng

    def walkers(self, n: int, seed: int = 1) -> np.ndarray:
        """theta* + randn * s/10, redrawn until within bounds (inference.py:442-451)."""
        if not self.spec.within_bounds(self.theta_true):
            raise ValueError(f"{self.name}: theta* is outside the bounds, no walker ball can be drawn around it")
        rng = np.random.default_rng(seed)
        out = np.empty((n, self.spec.ndim))
        todo = np.arange(n)
        while todo.size:
            trial = self.theta_true + rng.standard_normal((todo.size, self.spec.ndim)) * (self.prior_stds / 10.0)
            ok = np.array([self.spec.within_bounds(t) for t in trial])
            out[todo[ok]] = trial[ok]
            todo = todo[~ok]
        return out


def _trimmed_freqs(cat: MolCat, ll, ul):
    i0, i1 = cat.trim_bounds(ll, ul)
    return cat.frequency[i0:i1]


def make_problem(name: str, cat_folder: str, n_chan: int = 1 << 20, device: int = 0, seed: int = 0,
                 noise_k: float = 0.005) -> SyntheticProblem:
    """name:
       'hc5n_dsn'        : config 1, the reduced DSN sample shipped with the reference (real data)
       'benzonitrile_k1' : config 3, inference.py 5-dim layout (free source size), aligned_velocity 5.8
       'benzonitrile_k4' : config 3, 14-dim TMC1 layout
       'hc7n_hfs_k4'     : config 2 shape (hyperfine catalog, free source size, 4 vlsr components)
       'joint_k4'        : config 4, 1-cyanonapthalene + indene_hfs sharing ss/Tex/vlsr/dV (18-dim)"""
    def load(mol):
        p = find_catalog(cat_folder, mol)
        if p is None:
            raise FileNotFoundError(f"no catalog for {mol} in {cat_folder}")
        return MolCat(mol, p)

    if name == "hc5n_dsn":
        # BASELINE config 1: the reference's own CPU-runnable case -- hc5n_hfs on the reduced DSN sample (22 channels,
        # 9 lines), inference.py:585-631 template priors, fixed source size.  Real data, no synthesis.
        g = np.load(os.path.join(os.path.dirname(cat_folder), "hc5n_dsn_ref.npz"))
        bounds = {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e14], 'Tex': [3.5, 12.0], 'vlsr': [3.0, 5.5], 'dV': [0.4, 1.5]}
        spec = ModelSpec.inference(52.0, bounds, 70, 4.10, 18000, 25000)
        theta = np.array(g["fixed/prior_means"], dtype=float); theta[0] = float(g["fixed/mle_ncol"])
        return SyntheticProblem(name, spec, [load("hc5n_hfs")], [np.asarray(g["fixed/line_idx"])], g["fixed/grid_freq"],
                                g["fixed/grid_y"], g["fixed/grid_yerr"], theta, np.array(g["fixed/prior_means"], dtype=float),
                                np.array(g["fixed/prior_stds"], dtype=float))
    if name == "benzonitrile_k1":
        cats = [load("benzonitrile")]
        bounds = {'source_size': [0.0, 200.0], 'Ncol': [1e8, 1e14], 'Tex': [2.7, 15.0], 'vlsr': [5.0, 6.6], 'dV': [0.05, 0.3]}
        spec = ModelSpec.inference(None, bounds, 100, 5.8, 7000, 30000)
        theta = np.array([40.0, 2.15e11, 6.7, 5.8, 0.117])
        stds = np.array([4.0, 0.3e11, 0.1, 0.002, 0.002])
        v_centre = 0.0
    elif name in ("benzonitrile_k4", "hc7n_hfs_k4"):
        cats = [load(name[:-3])]
        spec = ModelSpec.tmc1(4, 1)
        theta = TMC1_MEANS.copy(); stds = TMC1_STDS.copy()
        if name == "benzonitrile_k4":
            theta[4:8] /= 10.0; stds[4:8] /= 10.0
        v_centre = 5.8
    elif name == "joint_k4":
        cats = [load("1-cyanonapthalene"), load("indene_hfs")]
        spec = ModelSpec.tmc1(4, 2)
        theta = np.r_[TMC1_MEANS[:4], TMC1_MEANS[4:8] / 10.0, TMC1_MEANS[4:8] / 2.0, TMC1_MEANS[8:]]
        stds = np.r_[TMC1_STDS[:4], TMC1_STDS[4:8] / 10.0, TMC1_STDS[4:8] / 2.0, TMC1_STDS[8:]]
        v_centre = 5.8
    else:
        raise ValueError(name)
    lines = np.sort(np.concatenate([_trimmed_freqs(c, spec.ll, spec.ul) for c in cats]))
    freq = window_grid(lines, n_chan, GOTHAM_DNU, v_centre)
    # noiseless truth from the fp64 kernels, then seeded noise
    prob = SyntheticProblem(name, spec, cats, [None] * len(cats), freq, np.zeros_like(freq), np.ones_like(freq),
                            theta, theta.copy(), stds)
    with prob.engine(device=device, precision="fp64") as eng:
        truth = eng.simulate(theta[None, :])[0]
    rng = np.random.default_rng(seed)
    prob.y = truth + rng.normal(0.0, noise_k, freq.size)
    prob.yerr = np.sqrt(noise_k ** 2 + (0.1 * prob.y) ** 2)
    return prob


def default_cat_folder() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), "tests", "golden", "catalog")
