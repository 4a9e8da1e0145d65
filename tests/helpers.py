"""Shared fixture builders: the SAME problem described once for the oracle (checker) and once for the
engine (thing under test)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

HC5N_BOUNDS = {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e14], 'Tex': [3.5, 12.0], 'vlsr': [3.0, 5.5], 'dV': [0.4, 1.5]}
SYNTH_BOUNDS = {'source_size': [0.0, 200.0], 'Ncol': [1e8, 1e14], 'Tex': [2.7, 15.0], 'vlsr': [5.0, 6.6], 'dV': [0.05, 0.3]}


def cat_path(name):
    return os.path.join(GOLD, "catalog", name + ".cat.gz")


def oracle_cat(name):
    from oracle import lte_oracle as O
    return O.parse_catalog(cat_path(name), name_for_q=name + ".cat")


def product_cat(name):
    from cha1_mcmc_b200 import MolCat
    return MolCat(name, cat_path(name))


def make_engine(spec, cats, grid, line_idx, prior=None, precision="mixed", device=0):
    """spec: cha1_mcmc_b200.ModelSpec; cats: list of product MolCat; grid: (x, y, yerr)."""
    from cha1_mcmc_b200 import LTEEngine
    eng = LTEEngine(device=device, precision=precision)
    eng.set_model(spec)
    for m, c in enumerate(cats):
        eng.set_molecule(m, c, line_idx=None if line_idx is None else line_idx[m])
    eng.set_spectrum(*grid)
    if prior is not None:
        eng.set_prior(prior[0], prior[1])     # (stds, means)
    return eng


def specs_inference(fixed, bounds, dish, al, ll, ul):
    from oracle import lte_oracle as O
    from cha1_mcmc_b200 import ModelSpec
    return O.spec_inference(fixed, bounds, dish, al, ll, ul), ModelSpec.inference(fixed, bounds, dish, al, ll, ul)


def specs_tmc1(K=4, n_mol=1):
    from oracle import lte_oracle as O
    from cha1_mcmc_b200 import ModelSpec
    return O.spec_tmc1(K, n_mol), ModelSpec.tmc1(K, n_mol)


def same_inf_pattern(a, b):
    """-inf / NaN lanes must coincide (engine never returns NaN: NaN in the reference == -inf here)."""
    fa = np.isfinite(a); fb = np.isfinite(b)
    return np.array_equal(fa, fb)
