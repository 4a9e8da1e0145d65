"""CUDA path vs the oracle beyond the reference-generated goldens: the BASELINE configs that have no real data
(joint fit, state-sum molecules, full-size synthetic grids), the stand-alone entry points, the on-device sampler,
the reference-shaped host class end to end, and the edge cases of the boundary.  All calls go through the C-ABI."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

LL_ATOL = 1e-3          # BASELINE.json north_star


def _assert_ll(got, want, grid, prec, rel=H.FAR_REL, tag="extended", weight=None):
    """The one tolerance rule of tests/helpers.py: fp64 path ~1e-12; mixed path pure 1e-3 absolute on well-fitting
    rows, 1e-3 + rel * chi2/2 on rows far from the data."""
    H.check_lnlike(got, want, grid[2], prec, tag, far_rel=rel, weight=weight)


def _oracle_pair(spec_o, ocats, grid, lidx, prior):
    """C restatement of the reference (the checker), same problem description as the engine."""
    from oracle.c_oracle import COracle
    return COracle(spec_o, ocats, (grid[0], grid[1], grid[2], lidx), prior=prior)


def _small_problem(mols, K, n_chan, seed, dnu=None, v_centre=5.8, noise=0.005, ncol_scale=1.0):
    """Synthetic GOTHAM-like problem on a small grid: returns (oracle spec, product spec, oracle cats, product cats,
    grid, theta*, stds)."""
    from cha1_mcmc_b200 import synthetic as SY
    from oracle import lte_oracle as O
    so, sp = H.specs_tmc1(K, len(mols))
    ocats = [H.oracle_cat(m) for m in mols]
    pcats = [H.product_cat(m) for m in mols]
    n_ss = K
    base_ncol = SY.TMC1_MEANS[4:8][:K]
    theta = np.r_[SY.TMC1_MEANS[:4][:K], *[base_ncol / (10.0 if m == 0 else 2.0) for m in range(len(mols))],
                  SY.TMC1_MEANS[8], SY.TMC1_MEANS[9:13][:K], SY.TMC1_MEANS[13]]
    stds = np.r_[SY.TMC1_STDS[:4][:K], *[SY.TMC1_STDS[4:8][:K] / (10.0 if m == 0 else 2.0) for m in range(len(mols))],
                 SY.TMC1_STDS[8], SY.TMC1_STDS[9:13][:K], SY.TMC1_STDS[13]]
    assert theta.size == sp.ndim and n_ss == K
    for row in sp.idx_ncol:
        theta[row] *= ncol_scale; stds[row] *= ncol_scale
    lines = np.concatenate([SY._trimmed_freqs(c, sp.ll, sp.ul) for c in pcats])
    strength = np.concatenate([c.logint[slice(*c.trim_bounds(sp.ll, sp.ul))] for c in pcats])
    rng = np.random.default_rng(seed)
    # grid around the 48 strongest lines + 16 random ones (every line of every catalog still enters the model)
    strongest = lines[np.argsort(strength)[::-1][:48]]
    pick = np.unique(np.r_[strongest, rng.choice(lines, size=min(lines.size, 16), replace=False)])
    freq = SY.window_grid(pick, n_chan, dnu or SY.GOTHAM_DNU, v_centre)
    lidx = [np.arange(c.trim_bounds(sp.ll, sp.ul)[1] - c.trim_bounds(sp.ll, sp.ul)[0]) for c in pcats]
    truth = O.simulate(so, ocats, lidx, freq, theta, windowed=True)
    y = truth + rng.normal(0.0, noise, freq.size)
    yerr = np.sqrt(noise ** 2 + (0.1 * y) ** 2)
    return so, sp, ocats, pcats, (freq, y, yerr), lidx, theta, stds


def _ball(spec, theta, stds, n, seed, scale=0.1):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        t = theta + rng.standard_normal(theta.size) * stds * scale
        if spec.within_bounds(t):
            out.append(t)
    return np.array(out)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 4 (scaled): joint fit of two state-sum molecules sharing ss/Tex/vlsr/dV, K = 4
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp64", "mixed"])
def test_joint_two_molecule_fit_matches_oracle(prec):
    so, sp, ocats, pcats, grid, lidx, theta, stds = _small_problem(["1-cyanonapthalene", "indene_hfs"], 4, 3000, seed=3,
                                                                    ncol_scale=20.0)
    th = np.vstack([_ball(sp, theta, stds, 12, 5), _ball(sp, theta, stds, 6, 6, scale=3.0)])
    co = _oracle_pair(so, ocats, grid, lidx, (stds, theta))
    want_ll, want_lp = co.lnlike(th), co.lnprob(th)
    eng = H.make_engine(sp, pcats, grid, None, prior=(stds, theta), precision=prec)
    got_ll, got_lp = eng.log_like(th), eng.log_prob(th)
    _assert_ll(got_ll, want_ll, grid, prec)
    _assert_ll(got_lp, want_lp, grid, prec)
    # the walker ball (first 12 rows) is the regime BASELINE's 1e-3 absolute tolerance is about
    np.testing.assert_allclose(got_ll[:12], want_ll[:12], atol=LL_ATOL if prec == "mixed" else 1e-8, rtol=0)
    # model spectra
    ref = co.simulate(th[:3])
    mod = eng.simulate(th[:3])
    peak = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(mod - ref) / peak) < (1e-11 if prec == "fp64" else 1e-5)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 5 (scaled): molecules that take the state-sum partition function, and the analytic branches
# ---------------------------------------------------------------------------------------------------------
NCOL_SCALE = {"C8H-": 1.0, "cyclopentadiene": 3000.0, "hc2nc": 60.0, "hc3n": 1.0, "phenol": 100.0, "1-cyanonapthalene": 20.0}


@pytest.mark.parametrize("mol", sorted(NCOL_SCALE))
def test_every_partition_function_branch_matches_oracle(mol):
    so, sp, ocats, pcats, grid, lidx, theta, stds = _small_problem([mol], 1, 2048, seed=11, ncol_scale=NCOL_SCALE[mol])
    th = _ball(sp, theta, stds, 16, 2, scale=1.0)
    th[:, sp.idx_tex] = np.linspace(3.0, 40.0, len(th))          # sweep Tex: Q(T) is the point of this test
    co = _oracle_pair(so, ocats, grid, lidx, (stds, theta))
    want = co.lnlike(th)
    for prec in ("fp64", "mixed"):
        with H.make_engine(sp, pcats, grid, None, prior=(stds, theta), precision=prec) as eng:
            _assert_ll(eng.log_like(th), want, grid, prec)


def test_dsn_like_grid_wide_channels_matches_oracle():
    """DSN-like sampling (30.5 kHz channels ~ 0.43 km/s: lines only ~2 channels wide), inference.py layout."""
    from cha1_mcmc_b200 import synthetic as SY
    from oracle import lte_oracle as O
    bounds = {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e14], 'Tex': [3.5, 12.0], 'vlsr': [3.0, 5.5], 'dV': [0.2, 1.5]}
    so, sp = H.specs_inference(None, bounds, 70, 4.10, 18000, 25000)
    ocat, pcat = H.oracle_cat("hc7n_hfs"), H.product_cat("hc7n_hfs")
    i0, i1 = pcat.trim_bounds(18000, 25000)
    lidx = [np.arange(i1 - i0)]
    freq = SY.window_grid(pcat.frequency[i0:i1], 1500, SY.DSN_DNU, 0.0)
    theta = np.array([50.0, 4e12, 7.0, 4.3, 0.6]); stds = np.array([5.0, 1e12, 1.0, 0.1, 0.1])
    truth = O.simulate(so, [ocat], lidx, freq, theta, windowed=True)
    rng = np.random.default_rng(0)
    y = truth + rng.normal(0, 0.01, freq.size); yerr = np.sqrt(0.01 ** 2 + (0.1 * y) ** 2)
    th = _ball(sp, theta, stds, 48, 3, scale=1.0)
    co = _oracle_pair(so, [ocat], (freq, y, yerr), lidx, (stds, theta))
    want = co.lnprob(th)
    for prec in ("fp64", "mixed"):
        with H.make_engine(sp, [pcat], (freq, y, yerr), None, prior=(stds, theta), precision=prec) as eng:
            _assert_ll(eng.log_prob(th), want, (freq, y, yerr), prec)


# ---------------------------------------------------------------------------------------------------------
# stand-alone entry points: cha_stick_spectrum (MolSim.run_sim) and cha_make_model (make_model_numba)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mol", ["hc5n_hfs", "benzonitrile", "indene_hfs", "phenol"])
def test_stick_spectrum_matches_reference_molsim(mol):
    """tau_sim / int_sim of MolSim(C=1e12, dV=0.3, T=7, source_size=40, ll=7000, ul=30000) as stored by the
    unmodified reference (oracle/make_golden.py::golden_catalogs)."""
    g = np.load(H.GOLD + "/catalogs_ref.npz")
    cat = H.product_cat(mol)
    from cha1_mcmc_b200 import LTEEngine, ModelSpec
    with LTEEngine(device=0, precision="fp64") as eng:
        eng.set_model(ModelSpec.tmc1(1, 1))
        eng.set_molecule(0, cat, 7000, 30000)
        f, ints, tau = eng.stick_spectrum(0, cat.frequency.size, 1.0e12, 7.0, 0.3, 40.0, 100)
    assert np.array_equal(f, g[f"{mol}/freq_sim"])
    np.testing.assert_allclose(tau, g[f"{mol}/tau_sim"], rtol=1e-12)      # exp(-El/kT) with |arg| up to ~500
    # 1 - exp(-tau) for tau ~ 1e-13 carries the absolute rounding of exp() near 1 (one ulp of 1.0 times dJ ~ 1 K)
    np.testing.assert_allclose(ints, g[f"{mol}/int_sim"], rtol=1e-12, atol=1e-15)


def test_make_model_entry_point_matches_oracle_and_reference_models():
    g = np.load(H.GOLD + "/hc5n_dsn_ref.npz")
    from oracle import lte_oracle as O
    from cha1_mcmc_b200 import LTEEngine, ModelSpec
    ocat = H.oracle_cat("hc5n_hfs")
    th = g["fixed/theta"][:16]
    x = g["fixed/grid_freq"]
    with LTEEngine(device=0) as eng:
        eng.set_model(ModelSpec.inference(52.0, H.HC5N_BOUNDS, 70, 4.10, 18000, 25000))
        for t, ref in zip(th, g["fixed/models"]):
            f, tau = O.line_taus(ocat, t[0], t[1], t[3], 18000, 25000)
            sel = g["fixed/line_idx"]
            got = eng.make_model(f[sel], tau[sel], x, t[2], t[3], t[1], 52.0, 4.10, 70)
            np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-14 * np.abs(ref).max())


# ---------------------------------------------------------------------------------------------------------
# edge cases of the boundary
# ---------------------------------------------------------------------------------------------------------
def test_boundary_edge_cases():
    g = np.load(H.GOLD + "/hc5n_dsn_ref.npz")
    so, sp = H.specs_inference(52.0, H.HC5N_BOUNDS, 70, 4.10, 18000, 25000)
    cat = H.product_cat("hc5n_hfs")
    x, y, e = g["fixed/grid_freq"], g["fixed/grid_y"], g["fixed/grid_yerr"]
    pr = (g["fixed/prior_stds"], g["fixed/prior_means"])
    th = g["fixed/theta"]
    ref = g["fixed/lnprob"]
    for prec in ("fp64", "mixed"):
        with H.make_engine(sp, [cat], (x, y, e), [g["fixed/line_idx"]], prior=pr, precision=prec) as eng:
            # empty batch, batch of one, ragged batch sizes around the 128-walker block
            assert eng.log_prob(np.empty((0, 4))).shape == (0,)
            for n in (1, 2, 127, 128, 129, 201):
                got = eng.log_prob(th[:n])
                H.check_lnlike(got, ref[:n], e, prec, f"edge/batch of {n}", prior=g["fixed/lnprior"][:n])
            # a walker's value does not depend on what else is in the batch, nor on its position in it
            a = eng.log_prob(th)
            b = eng.log_prob(th[::-1])[::-1]
            assert np.array_equal(a, b)
            c = np.concatenate([eng.log_prob(th[:77]), eng.log_prob(th[77:])])
            assert np.array_equal(a, c)
            # non-finite parameters -> -inf lane, never NaN, neighbours untouched
            bad = th[:8].copy(); bad[3, 1] = np.nan; bad[5, 3] = np.inf; bad[6, 3] = -0.1
            got = eng.log_prob(bad)
            assert np.all(np.isneginf(got[[3, 5, 6]])) and not np.any(np.isnan(got))
            keep = [0, 1, 2, 4, 7]
            assert np.array_equal(got[keep], a[keep])
        # channel order given by the caller does not matter (the reference sums over channels in the given order;
        # fp64 sum order changes the last bits only)
        perm = np.random.default_rng(0).permutation(x.size)
        with H.make_engine(sp, [cat], (x[perm], y[perm], e[perm]), [g["fixed/line_idx"]], prior=pr, precision=prec) as eng2:
            got = eng2.log_prob(th[:64])
            H.check_lnlike(got, ref[:64], e, prec, "edge/permuted channels", prior=g["fixed/lnprior"][:64])
            mod = eng2.simulate(th[:4])
            refm = g["fixed/models"][:4][:, perm]
            assert np.max(np.abs(mod - refm)) / np.abs(refm).max() < (1e-12 if prec == "fp64" else 1e-5)
        # no selected line at all: model == 0, lnlike is the constant -0.5*sum(y^2/s^2 - ln(1/s^2))
        with H.make_engine(sp, [cat], (x, y, e), [np.array([], dtype=np.int64)], prior=pr, precision=prec) as eng3:
            want = -0.5 * np.sum(y ** 2 / e ** 2 - np.log(1 / e ** 2))
            got = eng3.log_like(th[:5])
            np.testing.assert_allclose(got, want, rtol=1e-14)
            assert np.all(eng3.simulate(th[:2]) == 0.0)
        # a spectrum that no line window touches (all channels inactive)
        xf = np.linspace(19000.0, 19000.5, 40)
        with H.make_engine(sp, [cat], (xf, y[:1].repeat(40), e[:1].repeat(40)), [g["fixed/line_idx"]], prior=pr,
                           precision=prec) as eng4:
            want = -0.5 * np.sum((y[0] / e[0]) ** 2 - np.log(1 / e[0] ** 2)) * 40
            np.testing.assert_allclose(eng4.log_like(th[:3]), want, rtol=1e-13)


def test_error_reporting_never_throws_across_the_abi():
    from cha1_mcmc_b200 import LTEEngine, EngineError, ModelSpec
    with LTEEngine(device=0) as eng:
        with pytest.raises(EngineError):
            eng.log_prob(np.zeros((4, 4)))                  # nothing configured yet
        eng.set_model(ModelSpec.inference(52.0, H.HC5N_BOUNDS, 70, 4.10, 18000, 25000))
        with pytest.raises(EngineError):
            eng.log_prob(np.zeros((4, 4)))                  # no spectrum / molecule
        cat = H.product_cat("hc5n_hfs")
        eng.set_molecule(0, cat, line_idx=np.array([10 ** 6]))
        eng.set_spectrum(np.array([20000.0]), np.array([0.0]), np.array([1.0]))
        with pytest.raises(EngineError, match="out of range"):
            eng.log_like(np.array([[3e12, 8.0, 4.3, 0.7]]))


# ---------------------------------------------------------------------------------------------------------
# full BASELINE size (2^20 channels, 8192 walkers): size-independent properties + oracle on a sample
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_size_problem():
    from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder
    prob = make_problem("benzonitrile_k1", default_cat_folder(), n_chan=1 << 20, device=0, seed=0)
    return prob


def test_full_size_benzonitrile_properties(full_size_problem):
    prob = full_size_problem
    th = prob.walkers(8192, seed=1)
    with prob.engine(precision="mixed") as eng:
        lp = eng.log_prob(th)
        assert lp.shape == (8192,) and np.all(np.isfinite(lp))
        # (1) sharding independence: any split of the batch gives bit-identical per-walker values
        parts = np.concatenate([eng.log_prob(th[a:b]) for a, b in ((0, 1000), (1000, 4096), (4096, 8192))])
        assert np.array_equal(parts, lp)
        # (2) lnprob = lnprior + lnlike
        np.testing.assert_allclose(eng.log_prior(th[:256]) + eng.log_like(th[:256]), lp[:256], rtol=1e-14)
        # (3) chi-square recomputed on the host from the spectrum the channel-stream kernel writes
        sim = eng.simulate(th[:4])
        assert sim.shape == (4, prob.freq.size)
        w = 1.0 / prob.yerr ** 2
        host = np.array([-0.5 * np.sum((prob.y - m) ** 2 * w - np.log(w)) for m in sim])
        np.testing.assert_allclose(eng.log_like(th[:4]), host, atol=LL_ATOL, rtol=0)
        # (4) the truth maximises the likelihood along Ncol (MLE sanity at full size)
        t0 = prob.theta_true.copy()
        scan = np.repeat(t0[None], 21, 0); scan[:, 1] *= np.linspace(0.5, 1.5, 21)
        ll = eng.log_like(scan)
        assert 7 <= int(np.argmax(ll)) <= 13
        mixed_lp = lp
    # (5) mixed vs the all-fp64 kernel (reference operation order, full 10 dV masks) at full size
    with prob.engine(precision="fp64") as eng64:
        lp64 = eng64.log_prob(th)                         # all 8192 walkers (23 ms on the device)
        np.testing.assert_allclose(mixed_lp, lp64, atol=LL_ATOL, rtol=0)
    # (6) the C restatement of the reference algorithm on 24 walkers at full size (O(L*C) each: ~8 s on 16 threads)
    from bench import to_oracle_spec
    from oracle.c_oracle import COracle
    from oracle import lte_oracle as O
    ocat = O.parse_catalog(prob.cats[0].catalog_file, name_for_q="benzonitrile.cat")
    i0, i1 = prob.cats[0].trim_bounds(prob.spec.ll, prob.spec.ul)
    co = COracle(to_oracle_spec(prob.spec), [ocat], (prob.freq, prob.y, prob.yerr, [np.arange(i1 - i0)]),
                 prior=(prob.prior_stds, prob.prior_means))
    np.testing.assert_allclose(mixed_lp[:24], co.lnprob(th[:24]), atol=LL_ATOL, rtol=0)


# ---------------------------------------------------------------------------------------------------------
# on-device ensemble sampler (row N1) vs its CPU restatement, and the reference-shaped class end to end
# ---------------------------------------------------------------------------------------------------------
def _hc5n_setup(prec="fp64"):
    g = np.load(H.GOLD + "/hc5n_dsn_ref.npz")
    so, sp = H.specs_inference(52.0, H.HC5N_BOUNDS, 70, 4.10, 18000, 25000)
    cat, ocat = H.product_cat("hc5n_hfs"), H.oracle_cat("hc5n_hfs")
    grid = (g["fixed/grid_freq"], g["fixed/grid_y"], g["fixed/grid_yerr"])
    pr = (g["fixed/prior_stds"], g["fixed/prior_means"])
    eng = H.make_engine(sp, [cat], grid, [g["fixed/line_idx"]], prior=pr, precision=prec)
    co = _oracle_pair(so, [ocat], grid, [g["fixed/line_idx"]], pr)
    return g, sp, eng, co


def test_device_sampler_follows_its_cpu_restatement_step_by_step():
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler
    from oracle import device_sampler_oracle as D
    g, sp, eng, co = _hc5n_setup("fp64")
    mu, sd = g["fixed/prior_means"].copy(), g["fixed/prior_stds"]
    mu[0] = float(g["fixed/mle_ncol"])
    p0 = _ball(sp, mu, sd, 64, 4)
    smp = DeviceEnsembleSampler(eng, 64, p0, seed=42)
    chain, logp = smp.run(30)
    ref_chain, ref_lp, ref_acc = D.run(p0, co.lnprob, 30, seed=42)
    assert chain.shape == (64, 30, 4)
    np.testing.assert_allclose(np.swapaxes(chain, 0, 1), ref_chain, rtol=1e-10)
    np.testing.assert_allclose(logp[:, -1], ref_lp, atol=1e-8)
    assert smp.state()[2] == ref_acc
    assert 0.15 < ref_acc / (30 * 64) < 0.95
    eng.close()


def test_device_sampler_large_ensemble_with_reach_sorted_batches_follows_its_cpu_restatement():
    """>= 1024 proposals per half-step: the evaluation batch is ordered by reach class (reach_sort_kernel) and the
    chain must not notice.  A wide initial ball makes the classes differ."""
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler
    from oracle import device_sampler_oracle as D
    g, sp, eng, co = _hc5n_setup("fp64")
    mu, sd = g["fixed/prior_means"].copy(), g["fixed/prior_stds"]
    mu[0] = float(g["fixed/mle_ncol"])
    p0 = _ball(sp, mu, sd, 4096, 4, scale=1.0)
    smp = DeviceEnsembleSampler(eng, 4096, p0, seed=7)
    chain, logp = smp.run(5)
    ref_chain, ref_lp, ref_acc = D.run(p0, co.lnprob, 5, seed=7)
    np.testing.assert_allclose(np.swapaxes(chain, 0, 1), ref_chain, rtol=1e-10)
    np.testing.assert_allclose(logp[:, -1], ref_lp, atol=1e-8)
    assert smp.state()[2] == ref_acc
    eng.close()
    # mixed path, sorted batches against unsorted evaluation of the same positions
    g, sp, eng, co = _hc5n_setup("mixed")
    smp = DeviceEnsembleSampler(eng, 4096, p0, seed=7)
    smp.run(3)
    c, lp, _ = smp.state()
    direct = eng.log_prob(c)
    assert H.same_inf_pattern(lp, direct)
    m = np.isfinite(direct)
    np.testing.assert_allclose(lp[m], direct[m], atol=2e-4, rtol=0)
    # Two list sets: after the first synchronisation point the bulk of every half-step's proposals is evaluated against
    # a narrower set chosen from the class histogram of the whole ensemble, the outliers against the primary lists
    # (padded, block-aligned batch).  The resident log-probs must still be those of the resident positions.
    for _ in range(3):
        smp.run(40, store_every=0)
        smp.sync()
    st = eng.stats()
    # (on this 22-channel grid both sets hold every (line, channel) pair; the batches are still split and padded)
    assert st["tight_builds"] >= 1 and 0 < st["tight_hv"] < st["hv_list"] and 0 < st["tight_pairs"] <= st["pairs"], st
    c, lp, nacc = smp.state()
    assert 0.1 < nacc / (4096 * 123) < 0.9
    direct = eng.log_prob(c)
    assert H.same_inf_pattern(lp, direct) and np.all(np.isfinite(lp))
    np.testing.assert_allclose(lp, direct, atol=2e-4, rtol=0)
    eng.close()


def test_spectralfitmcmc_end_to_end_matches_reference_mle_and_posterior(tmp_path):
    """The reference-shaped class on BASELINE config 1: data reduction, MLE column density (golden from the
    unmodified reference), short chain; posterior medians of the device-evaluated chain agree with the same
    sampler driven by the oracle's lnprob on the same seed (identical accept/reject decisions => identical chain)."""
    from cha1_mcmc_b200.inference import SpectralFitMCMC, posterior_summary
    from cha1_mcmc_b200.sampler import EnsembleSampler
    g = np.load(H.GOLD + "/hc5n_dsn_ref.npz")
    catdir = tmp_path / "catalog"; catdir.mkdir()
    import gzip, shutil
    with gzip.open(H.cat_path("hc5n_hfs"), "rb") as fi, open(catdir / "hc5n_hfs.cat", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    config = {
        'mol_name': 'hc5n_hfs', 'template_run': True, 'nruns': 60, 'nwalkers': 32,
        'bounds': H.HC5N_BOUNDS, 'template_means': np.array([3.4e10, 8.0, 4.3, 0.7575]),
        'template_stds': np.array([0.34e10, 3.0, 0.06, 0.22]), 'dish_size': 70, 'lower_limit': 18000,
        'upper_limit': 25000, 'aligned_velocity': 4.10, 'fixed_source_size': 52, 'MLE_for_Ncol': True,
        'block_interlopers': True, 'parallelize': False, 'fit_folder': str(tmp_path / "fit"),
        'cat_folder': str(catdir), 'prior_path': '', 'precision': 'fp64', 'seed': 123,
        'data_paths': {'hc5n_hfs': os.path.join(H.GOLD, "data", "cha_mms1_hc5n_example.npy")},
    }
    fit = SpectralFitMCMC(config)
    datafile, catfile = fit.init_setup()
    dg = np.load(datafile, allow_pickle=True)
    assert np.array_equal(np.asarray(dg[3], int), g["fixed/line_idx"])
    np.testing.assert_allclose(np.asarray(dg[2], float), g["fixed/grid_yerr"], rtol=1e-12)
    chain = fit.fit_multi_gaussian(datafile, catfile)
    assert chain.shape == (32, 60, 4)
    assert np.array_equal(np.load(os.path.join(config['fit_folder'], 'hc5n_hfs', 'chain_template.npy')), chain)
    # MLE initialisation: the walker ball is centred on the MLE column density
    mle = float(g["fixed/mle_ncol"])
    from cha1_mcmc_b200 import MolCat
    est = fit.estimate_Ncol_via_MLE(dg, MolCat("mol", catfile), (8.0, 4.3, 0.7575))
    assert abs(est / mle - 1) < 1e-6
    # row N2: one batched profile launch brackets the optimum, Brent then runs on that bracket only
    probes_bracketed = fit.mle_probes
    est_full = fit.estimate_Ncol_via_MLE(dg, MolCat("mol", catfile), (8.0, 4.3, 0.7575), n_profile=0)   # the reference's way
    assert abs(est_full / mle - 1) < 1e-6 and probes_bracketed < fit.mle_probes
    grid_n, prof = fit.ncol_profile(dg, MolCat("mol", catfile), (8.0, 4.3, 0.7575), n=64)
    k = int(np.argmax(prof))
    assert grid_n[k - 1] < mle < grid_n[k + 1] and np.all(np.isfinite(prof))
    # same seed, oracle-evaluated chain
    so, _ = H.specs_inference(52.0, H.HC5N_BOUNDS, 70, 4.10, 18000, 25000)
    co = _oracle_pair(so, [H.oracle_cat("hc5n_hfs")], (g["fixed/grid_freq"], g["fixed/grid_y"], g["fixed/grid_yerr"]),
                      [g["fixed/line_idx"]], (config['template_stds'], config['template_means']))
    np.random.seed(123)
    initial = config['template_means'].copy(); initial[0] = est
    pos = []
    for _ in range(32):
        t = None
        while t is None or not fit.is_within_bounds(t):
            t = initial + np.random.randn(4) * (config['template_stds'] / 10.0)
        pos.append(t)
    pos = np.array(pos)
    s = EnsembleSampler(32, 4, co.lnprob, vectorize=True)
    for _ in range(60):
        s.run_mcmc(pos, 1); pos = s.chain[:, -1, :]
    med, med_ref = posterior_summary(chain)[:, 0], posterior_summary(s.chain)[:, 0]
    np.testing.assert_allclose(med, med_ref, rtol=1e-6)


def test_optimistic_device_calls_are_rerun_when_the_pair_list_did_not_cover_them():
    """cha_log_prob_dev launches against the current pair list without a host round trip; batches whose dV / vlsr
    range the list did not cover must be detected at the next sync and re-evaluated (same values as the host path)."""
    import torch
    g = np.load(H.GOLD + "/benzonitrile_synth_ref.npz")
    _, spec = H.specs_inference(None, H.SYNTH_BOUNDS, 100, 5.8, 7000, 30000)
    cat = H.product_cat("benzonitrile")
    grid = (g["grid_freq"], g["grid_y"], g["grid_yerr"])
    mu, sd = g["free/prior_means"], g["free/prior_stds"]
    rng = np.random.default_rng(5)

    def batch(dv, dvl, n=256):
        th = np.tile(mu, (n, 1)) + rng.standard_normal((n, 5)) * sd * 0.1
        th[:, 4] = rng.uniform(0.9 * dv, dv, n)
        th[:, 3] = 5.8 + rng.uniform(-dvl, dvl, n)
        return th
    seq = [batch(0.10, 0.002), batch(0.11, 0.002), batch(0.28, 0.4), batch(0.12, 0.002), batch(0.29, 0.7), batch(0.06, 0.0)]
    with H.make_engine(spec, [cat], grid, [g["line_idx"]], prior=(sd, mu), precision="mixed") as ref_eng:
        want = [ref_eng.log_prob(t) for t in seq]
    with H.make_engine(spec, [cat], grid, [g["line_idx"]], prior=(sd, mu), precision="mixed") as eng:
        d_th = [torch.from_numpy(t).cuda() for t in seq]
        outs = [torch.empty(len(t), dtype=torch.float64, device="cuda") for t in seq]
        eng.log_prob_device(d_th[0], out=outs[0], sync=True)          # builds the list for the narrow batch
        r0 = eng.stat("rebuilds")
        for t, o in zip(d_th[1:], outs[1:]):                            # five calls in flight, two of them uncovered
            eng.log_prob_device(t, out=o, sync=False)
        eng.sync()
        assert eng.stat("rebuilds") > r0
        for o, w_ in zip(outs, want):
            got = o.cpu().numpy()
            assert np.all(np.isfinite(got))
            np.testing.assert_allclose(got, w_, atol=2e-4, rtol=0)     # different list extents: tails < 1.5e-8 of a peak
    # host-buffer calls take the same optimistic route (no scan of theta before the launch): an uncovered batch is
    # evaluated again after the rebuild, inside the call
    with H.make_engine(spec, [cat], grid, [g["line_idx"]], prior=(sd, mu), precision="mixed") as eng:
        got = [eng.log_prob(t) for t in seq]
        assert eng.stat("rebuilds") >= 3
        for o, w_ in zip(got, want):
            assert np.all(np.isfinite(o))
            np.testing.assert_allclose(o, w_, atol=2e-4, rtol=0)


@pytest.mark.parametrize("workload", ["benzonitrile_k4", "joint_k4"])
def test_full_size_four_component_and_joint_configs(workload):
    """BASELINE configs 2/4 at the full 2^20-channel size (K = 4; 1-cyanonapthalene + indene_hfs: 42 325 lines): the NumPy
    oracle (windowed evaluation of the reference loop, same masks) on two walkers, model spectra through the
    channel-stream kernel, and sharding independence of a 1024-walker batch."""
    from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder
    from bench import to_oracle_spec
    from oracle import lte_oracle as O
    prob = make_problem(workload, default_cat_folder(), n_chan=1 << 20, device=0, seed=0)
    th = prob.walkers(1024, seed=3)
    ospec = to_oracle_spec(prob.spec)
    ospec.guard_nonfinite = False
    ocats = [O.parse_catalog(c.catalog_file, name_for_q=os.path.basename(c.catalog_file).replace(".gz", "")) for c in prob.cats]
    lidx = [np.arange(c.trim_bounds(prob.spec.ll, prob.spec.ul)[1] - c.trim_bounds(prob.spec.ll, prob.spec.ul)[0]) for c in prob.cats]
    dg = (prob.freq, prob.y, prob.yerr, lidx)
    want = np.array([O.lnprob(ospec, ocats, dg, t, prob.prior_stds, prob.prior_means, windowed=True) for t in th[:2]])
    want_model = O.simulate(ospec, ocats, lidx, prob.freq, th[0], windowed=True)
    with prob.engine(precision="mixed") as eng:
        lp = eng.log_prob(th)
        assert np.all(np.isfinite(lp))
        np.testing.assert_allclose(lp[:2], want, atol=LL_ATOL, rtol=0)
        halves = np.concatenate([eng.log_prob(th[:300]), eng.log_prob(th[300:])])
        assert np.array_equal(halves, lp)
        mod = eng.simulate(th[:1])[0]
        assert np.max(np.abs(mod - want_model)) <= 1e-5 * np.max(np.abs(want_model))
    with prob.engine(precision="fp64") as eng64:
        np.testing.assert_allclose(eng64.log_prob(th[:2]), want, atol=0, rtol=1e-12)       # 1e6 fp64 terms, other order


def test_spectralfitmcmc_device_sampler_checkpoints_the_chain(tmp_path):
    """config['sampler'] = 'device': chains resident in HBM, chain file (reference layout) rewritten every save_every
    steps and complete at the end; the chain equals the resident sampler driven directly with the same seed."""
    import gzip, shutil
    from cha1_mcmc_b200.inference import SpectralFitMCMC
    catdir = tmp_path / "catalog"; catdir.mkdir()
    with gzip.open(H.cat_path("hc5n_hfs"), "rb") as fi, open(catdir / "hc5n_hfs.cat", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    config = {
        'mol_name': 'hc5n_hfs', 'template_run': True, 'nruns': 12, 'nwalkers': 32,
        'bounds': H.HC5N_BOUNDS, 'template_means': np.array([3.0e12, 8.0, 4.3, 0.7575]),
        'template_stds': np.array([0.34e12, 3.0, 0.06, 0.22]), 'dish_size': 70, 'lower_limit': 18000,
        'upper_limit': 25000, 'aligned_velocity': 4.10, 'fixed_source_size': 52, 'MLE_for_Ncol': False,
        'block_interlopers': True, 'parallelize': False, 'fit_folder': str(tmp_path / "fit"),
        'cat_folder': str(catdir), 'prior_path': '', 'precision': 'mixed', 'seed': 7, 'sampler': 'device',
        'save_every': 5,
        'data_paths': {'hc5n_hfs': os.path.join(H.GOLD, "data", "cha_mms1_hc5n_example.npy")},
    }
    fit = SpectralFitMCMC(config)
    datafile, catfile = fit.init_setup()
    chain = fit.fit_multi_gaussian(datafile, catfile)
    assert chain.shape == (32, 12, 4) and np.all(np.isfinite(chain))
    saved = np.load(os.path.join(config['fit_folder'], 'hc5n_hfs', 'chain_template.npy'))
    assert np.array_equal(saved, chain)
    assert np.any(chain[:, -1, :] != chain[:, 0, :])                 # walkers moved


def test_randomised_problems_mixed_vs_fp64_paths():
    """Seeded random fits (random line subsets, grid spacings from 1 kHz to 60 kHz, K in 1..3, one or two molecules,
    theta spread far beyond a walker ball incl. absorption (Tex < Tbg), negative column densities and masks that bite):
    the production path must follow the all-fp64 reference-order kernels within the error model on every row, and
    both must agree on which rows are -inf.  Exercises the general (masked / signed) path, dense tiles that cannot be
    staged, padding groups and the multi-molecule fast path."""
    from cha1_mcmc_b200 import synthetic as SY
    rng = np.random.default_rng(2024)
    names = ["hc5n_hfs", "hc7n_hfs", "hc9n_hfs", "benzonitrile", "phenol", "C8H-", "hc3n"]
    n_checked = 0
    for trial in range(14):
        K = int(rng.integers(1, 4)); M = int(rng.integers(1, 3))
        mols = list(rng.choice(names, size=M, replace=False))
        _, sp = H.specs_tmc1(K, M)
        cats = [H.product_cat(m) for m in mols]
        dnu = float(rng.choice([1.0e-3, 1.4e-3, 6.1e-3, 30.5e-3]))
        # the fitted lines are a random subset per molecule and the grid is built around exactly those, so every
        # line's core is on the grid (the mixed path drops terms beyond 6 sigma of a line: 1.5e-8 of THAT line's peak)
        lidx, pick = [], []
        for c in cats:
            i0, i1 = c.trim_bounds(sp.ll, sp.ul)
            sel = np.sort(rng.choice(i1 - i0, size=min(i1 - i0, int(rng.integers(3, 40))), replace=False))
            lidx.append(sel); pick.append(c.frequency[i0:i1][sel])
        freq = SY.window_grid(np.sort(np.concatenate(pick)), int(rng.integers(400, 6000)), dnu, 5.8)
        nd = sp.ndim
        th = np.empty((96, nd))
        th[:, sp.idx_ss] = rng.uniform(5, 150, (96, K))
        for row in sp.idx_ncol:
            th[:, row] = 10 ** rng.uniform(10.5, 13.5, (96, K))
        th[:, sp.idx_tex] = rng.uniform(2.9, 30, 96)
        th[:, sp.idx_vlsr] = np.sort(rng.uniform(5.2, 6.4, (96, K)), axis=1)
        th[:, sp.idx_dv] = 10 ** rng.uniform(-1.3, -0.55, 96)
        th[5, sp.idx_tex] = 2.2                                  # absorption
        th[6, sp.idx_ncol[0][0]] = -1e12                         # negative column density
        th[7, sp.idx_dv] = 0.02; th[7, sp.idx_vlsr] = 5.8 + 0.3 * np.arange(K)   # mask edge inside the line
        y = rng.normal(0, 0.01, freq.size); yerr = np.full(freq.size, 0.01) * rng.uniform(0.5, 2.0, freq.size)
        stds = np.full(nd, 1.0); means = th.mean(axis=0)
        with H.make_engine(sp, cats, (freq, y, yerr), lidx, prior=(stds, means), precision="fp64") as e64:
            want = e64.log_like(th); want_all = e64.simulate(th); want_m = want_all[:6]
        with H.make_engine(sp, cats, (freq, y, yerr), lidx, prior=(stds, means), precision="mixed") as emx:
            got = emx.log_like(th); got_m = emx.simulate(th[:6])
        # The data are pure noise and theta is random, so nothing here is a "fit": the rows are classified by the error
        # model itself (B = sum |m| |y - m| / sigma^2 from the fp64 spectra).  Lines down to 1/20 of a 30.5 kHz channel
        # wide: the fp32 velocity argument (ulp of the channel offset over sigma) puts the model at a few 1e-6 of its
        # peak -- inside BASELINE's 1e-5 -- hence 1e-5 * B beyond the 1e-3 regime instead of 1e-6 * B.
        _assert_ll(got, want, (freq, y, yerr), "mixed", rel=1e-5, tag=f"fuzz/{trial} {'+'.join(mols)} K={K}",
                   weight=H.model_error_weight(want_all, y, yerr))
        peak = np.abs(want_m).max(axis=1, keepdims=True)
        assert np.all(np.abs(got_m - want_m) <= 1e-5 * peak + 1e-12), (trial, mols, K)
        n_checked += 1
    assert n_checked == 14


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 5 (scaled): every shipped catalog, DSN-like and GOTHAM-like fits, against the C oracle
# ---------------------------------------------------------------------------------------------------------
def _all_molecules():
    from cha1_mcmc_b200 import survey as SV
    from cha1_mcmc_b200.synthetic import default_cat_folder
    return SV.list_molecules(default_cat_folder())


def test_survey_covers_all_35_shipped_catalogs():
    assert len(_all_molecules()) == 35


@pytest.mark.parametrize("kind", ["dsn", "gotham"])
def test_survey_fits_of_every_catalog_match_oracle(kind):
    """Each of the 35 molecules, lines chosen by the reference's 5 % rule (capped at the 40 strongest so the oracle
    stays in seconds), +-1.5 km/s windows, its own theta* and walker ball: log-likelihood and log-prob of the fp64
    and mixed paths against the C restatement of the reference."""
    from cha1_mcmc_b200 import survey as SV
    from cha1_mcmc_b200.synthetic import default_cat_folder
    from oracle import lte_oracle as O
    folder = default_cat_folder()
    n_fits = 0
    for mol in _all_molecules():
        p = SV.survey_problem(mol, kind, folder, device=0, seed=3, max_lines=40)
        t = SV.DSN_TEMPLATE
        so = (O.spec_inference(t["ss"], SV.DSN_BOUNDS, t["dish"], t["aligned"], t["ll"], t["ul"]) if kind == "dsn"
              else O.spec_tmc1(4, 1))
        co = _oracle_pair(so, [H.oracle_cat(mol)], (p.freq, p.y, p.yerr), p.line_idx, (p.prior_stds, p.prior_means))
        th = p.walkers(12, seed=5)
        th = np.vstack([p.theta_true, th, _ball(p.spec, p.theta_true, p.prior_stds, 3, 9, scale=2.0)])
        want_ll, want_lp = co.lnlike(th), co.lnprob(th)
        for prec in ("fp64", "mixed"):
            with p.engine(device=0, precision=prec) as eng:
                _assert_ll(eng.log_like(th), want_ll, (p.freq, p.y, p.yerr), prec)
                _assert_ll(eng.log_prob(th), want_lp, (p.freq, p.y, p.yerr), prec)
        n_fits += 1
    assert n_fits == 35


# ---------------------------------------------------------------------------------------------------------
# small batches: the launch sequence is replayed as one CUDA graph (BASELINE config 1: 128 walkers, 22 channels)
# ---------------------------------------------------------------------------------------------------------
def test_small_batches_replay_as_cuda_graphs_with_identical_results():
    import torch
    from cha1_mcmc_b200 import synthetic as SY
    p = SY.make_problem("hc5n_dsn", SY.default_cat_folder())
    th = [p.walkers(128, seed=s) for s in range(6)]
    th[4][:, p.spec.idx_dv] *= 1.9                  # needs a wider pair list: rebuild -> captured graphs are stale
    th[4][5] = np.nan; th[4][6, p.spec.idx_tex] = 1e9       # dead lanes inside a replay
    with p.engine(precision="mixed") as plain:      # every batch at a different size first: never the same key twice
        want = []
        for k, t in enumerate(th):
            plain.log_prob(np.vstack([t, t[:k + 1]]))
            want.append(plain.log_prob(t) if k % 2 else plain.log_prob(np.vstack([t, t[:1]]))[:128])
    # host-buffer entry point (cha_log_prob): same batch size every call, as emcee does
    with p.engine(precision="mixed") as eng:
        got = [eng.log_prob(t) for t in th]
        again = [eng.log_prob(t) for t in th]
        assert eng.stat("graph_launches") >= 3     # rebuilds (th[4] widens the list, th[5] shrinks it) restart the count
        for g_, a_, w_ in zip(got, again, want):
            assert H.same_inf_pattern(g_, w_) and H.same_inf_pattern(a_, w_)
            m = np.isfinite(w_)
            np.testing.assert_allclose(g_[m], w_[m], atol=2e-4, rtol=0)     # list extents differ between engines
            np.testing.assert_allclose(a_[m], w_[m], atol=2e-4, rtol=0)
        # a new prior is a new configuration: the replayed graph must not keep the old one
        eng.set_prior(p.prior_stds * 0.5, p.prior_means)
        plain_lp = eng.log_prior(th[0][:100])       # different batch size: plain launches
        lp = eng.log_prior(th[0]); lp = eng.log_prior(th[0]); lp = eng.log_prior(th[0])
        assert np.array_equal(lp[:100], plain_lp)
    # device-pointer entry point: theta rewritten in place between calls, same pointers -> replays
    with p.engine(precision="mixed") as eng:
        d_th = torch.empty((128, p.spec.ndim), dtype=torch.float64, device="cuda")
        d_out = torch.empty(128, dtype=torch.float64, device="cuda")
        for rep in range(2):
            for t, w_ in zip(th, want):
                d_th.copy_(torch.from_numpy(t)); torch.cuda.synchronize()
                eng.log_prob_device(d_th, out=d_out, sync=True)
                o = d_out.cpu().numpy()
                assert H.same_inf_pattern(o, w_)
                m = np.isfinite(w_)
                np.testing.assert_allclose(o[m], w_[m], atol=2e-4, rtol=0)
        assert eng.stat("graph_launches") >= 6


# ---------------------------------------------------------------------------------------------------------
# A walker's log-probability is a function of the walker and the resident lists alone -- never of its batch-mates
# ---------------------------------------------------------------------------------------------------------
def test_odd_rows_do_not_change_their_batch_mates():
    """Rows that cannot take the packed fast path (10 dV mask edge inside the walker's own 6 sigma, negative column
    density, Tex at the background temperature) used to send their whole 128-walker block down the general path,
    changing the other rows at the 1e-7 level.  The path is now chosen per walker: with the same resident lists the
    regular rows must come out BIT-identical whether or not odd rows share their block, and whatever their position."""
    g = np.load(H.GOLD + "/benzonitrile_synth_ref.npz")
    so, spec = H.specs_inference(None, H.SYNTH_BOUNDS, 100, 5.8, 7000, 30000)
    cat, ocat = H.product_cat("benzonitrile"), H.oracle_cat("benzonitrile")
    grid = (g["grid_freq"], g["grid_y"], g["grid_yerr"])
    mu, sd = g["free/prior_means"], g["free/prior_stds"]
    rng = np.random.default_rng(11)
    n = 384
    reg = np.tile(mu, (n, 1)) + rng.standard_normal((n, 5)) * sd * 0.1
    reg[0, 4] = 0.2; reg[0, 3] = 5.8 + 0.9                      # row 0 fixes the lists' extent in both batches
    odd = reg.copy()
    rows = {"mask_edge": 37, "negative_ncol": 130, "tex_at_tbg": 131, "mask_edge_2": 300}
    odd[rows["mask_edge"], 4] = 0.10; odd[rows["mask_edge"], 3] = 5.8 + 0.8      # |vlsr - al| = 0.8 > dV (10 - 6/2.355)
    odd[rows["mask_edge_2"], 4] = 0.06; odd[rows["mask_edge_2"], 3] = 5.8 - 0.5
    odd[rows["negative_ncol"], 1] = -2.0e11
    odd[rows["tex_at_tbg"], 2] = 2.7
    others = np.setdiff1d(np.arange(n), list(rows.values()))
    with H.make_engine(spec, [cat], grid, [g["line_idx"]], prior=(sd, mu), precision="mixed") as eng:
        a = eng.log_like(reg)
        r0 = eng.stat("rebuilds")
        b = eng.log_like(odd)
        assert eng.stat("rebuilds") == r0, "the two batches must be evaluated against the same lists"
        assert np.array_equal(a[others], b[others]), "odd rows changed their batch-mates"
        perm = rng.permutation(n)
        c = eng.log_like(odd[perm])
        assert eng.stat("rebuilds") == r0
        assert np.array_equal(c, b[perm]), "a row's value depends on its position in the batch"
        # the odd rows themselves are right (general path): against the C restatement of the reference
        co = _oracle_pair(so, [ocat], grid, [g["line_idx"]], (sd, mu))
        idx = np.array(list(rows.values()))
        want = co.lnlike(odd[idx])
        _assert_ll(b[idx], want, grid, "mixed", tag="odd rows (general path)")


# ---------------------------------------------------------------------------------------------------------
# north_star: "posterior medians must agree within MCMC error" -- device sampler vs the emcee-algorithm host sampler
# ---------------------------------------------------------------------------------------------------------
def test_device_sampler_posterior_agrees_with_the_emcee_algorithm_host_sampler():
    """The resident sampler deliberately differs from emcee in its red/blue split (parity of the walker id, not a
    shuffle) and its RNG (Philox, not MT19937), so its chains cannot be bit-compared with the host sampler, which
    restates emcee 3.1.6's move.  Both sample the same posterior: on BASELINE config 1 (hc5n_hfs on the DSN sample)
    the medians and the 16-84 widths of the two chains must agree within their Monte-Carlo error."""
    from cha1_mcmc_b200.inference import posterior_summary
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler, EnsembleSampler
    g, sp, eng, _ = _hc5n_setup("mixed")
    mu, sd = g["fixed/prior_means"].copy(), g["fixed/prior_stds"]
    mu[0] = float(g["fixed/mle_ncol"])
    nw, nsteps = 256, 600
    p0 = _ball(sp, mu, sd, nw, 21)
    dev = DeviceEnsembleSampler(eng, nw, p0, seed=99)
    dchain, _ = dev.run(nsteps)
    np.random.seed(4321)
    host = EnsembleSampler(nw, 4, eng.log_prob, vectorize=True)
    host.run_mcmc(p0, nsteps)
    hchain = host.chain
    sd_, sh_ = posterior_summary(dchain, burn_frac=0.5), posterior_summary(hchain, burn_frac=0.5)
    acc_d = dev.state()[2] / (nw * nsteps)
    acc_h = host.naccepted.sum() / (nw * nsteps)
    assert 0.2 < acc_d < 0.8 and abs(acc_d - acc_h) < 0.05, (acc_d, acc_h)

    def mc_error(chain):
        """standard error of a median from the scatter of 8 walker sub-ensembles (autocorrelation included)"""
        parts = np.array_split(np.arange(nw), 8)
        meds = np.array([np.median(chain[ix, nsteps // 2:, :].reshape(-1, 4), axis=0) for ix in parts])
        return meds.std(axis=0, ddof=1) / np.sqrt(8)
    err = np.hypot(mc_error(dchain), mc_error(hchain))
    width = 0.5 * (sd_[:, 1] + sd_[:, 2])
    dmed = np.abs(sd_[:, 0] - sh_[:, 0])
    print("[posterior] |d median| / MC error:", dmed / err, " / posterior sigma:", dmed / width, "acceptance", acc_d, acc_h)
    assert np.all(dmed < 5.0 * err + 0.02 * width), (dmed, err, width)
    # widths (16-84) agree to 10 %
    np.testing.assert_allclose(sd_[:, 1] + sd_[:, 2], sh_[:, 1] + sh_[:, 2], rtol=0.10)
    eng.close()


def test_survey_fits_cut_over_ranks_give_the_same_log_probs():
    """Config 5 sharding: expensive fits are cut into walker blocks that land on different ranks.  The pieces evaluated by
    two 'ranks' (two MoleculeSurvey objects on one GPU) must reproduce the whole-fit evaluation bit for bit, every walker
    exactly once."""
    from cha1_mcmc_b200 import survey as SV
    from cha1_mcmc_b200.synthetic import default_cat_folder
    folder = default_cat_folder()
    probs = [SV.survey_problem(m, k, folder, device=0, seed=3, max_lines=30) for m, k in
             (("benzonitrile", "gotham"), ("hc5n_hfs", "dsn"), ("phenol", "dsn"), ("hc3n", "dsn"))]
    costs = [SV.fit_cost(p) for p in probs]
    nw = 640
    whole = SV.MoleculeSurvey(probs, nw, device=0, seeds=[1 + i for i in range(len(probs))])
    ref = whole.log_prob()
    whole.close()
    parts = SV.shard_fit_walkers(costs, 2, nw)
    assert any(b - a < nw for p in parts for _, a, b in p), "no fit was cut: the test would not exercise the pieces"
    got = [np.full(nw, np.nan) for _ in probs]
    for mine in parts:
        sv = SV.MoleculeSurvey([probs[i] for i, _, _ in mine], nw, device=0, ranges=[(a, b) for _, a, b in mine],
                               seeds=[1 + i for i, _, _ in mine])
        for (i, a, b), lp in zip(mine, sv.log_prob()):
            assert np.all(np.isnan(got[i][a:b]))
            got[i][a:b] = lp
        sv.close()
    for g_, r_ in zip(got, ref):
        # different batch extents size the lists differently: tails beyond 6 sigma (< 1.5e-8 of a peak)
        assert H.same_inf_pattern(g_, r_)
        m = np.isfinite(r_)
        np.testing.assert_allclose(g_[m], r_[m], atol=2e-4, rtol=0)


# ---------------------------------------------------------------------------------------------------------
# channel-stream path: the one-pass span kernel (every output byte written once, TMA bulk stores) against the
# two-pass path (zero-fill + tiles) it replaces for spectra in ascending channel order
def _simulate_with(prob_engine, th, span):
    old = os.environ.get("CHALTE_SPAN_STREAM")
    os.environ["CHALTE_SPAN_STREAM"] = "2" if span else "0"
    try:
        with prob_engine() as eng:
            out = eng.simulate(th)
            launches = eng.stat("launches")
    finally:
        if old is None:
            os.environ.pop("CHALTE_SPAN_STREAM", None)
        else:
            os.environ["CHALTE_SPAN_STREAM"] = old
    return out, launches


def test_one_pass_channel_stream_equals_the_two_pass_path(full_size_problem):
    """Same lists, same arithmetic: the spectra must be bit-identical, at the full 2^20-channel size with a walker
    count that is neither a multiple of the 8-row sub-block nor of the 32-walker CTA, rows outside the bounds
    (all-zero spectra), and on small grids whose size is not a multiple of the span (K = 2 and a joint K = 4 fit)."""
    prob = full_size_problem
    th = prob.walkers(45, seed=5)
    th[7, -1] = -1.0                                   # dV < 0: out of bounds, the row must be exactly zero
    a, _ = _simulate_with(lambda: prob.engine(precision="mixed"), th, True)
    b, _ = _simulate_with(lambda: prob.engine(precision="mixed"), th, False)
    assert a.shape == (45, prob.freq.size) and np.array_equal(a, b)
    assert np.all(a[7] == 0.0) and np.count_nonzero(a[0]) > 1000
    from cha1_mcmc_b200 import LTEEngine
    for mols, K, n_chan in ((["benzonitrile"], 2, 1000), (["1-cyanonapthalene", "indene_hfs"], 4, 3 * 512 + 130), (["hc5n_hfs"], 1, 20), (["hc5n_hfs"], 1, 21)):        # 21: an odd channel count takes the two-pass path either way
        so, sp, ocats, pcats, grid, lidx, theta, stds = _small_problem(mols, K, n_chan, seed=11)
        thb = _ball(sp, theta, stds, 19, seed=2)

        def mk():
            eng = LTEEngine(device=0, precision="mixed")
            eng.set_model(sp)
            for m, c in enumerate(pcats):
                eng.set_molecule(m, c, line_idx=lidx[m])
            eng.set_spectrum(*grid)
            return eng
        a, _ = _simulate_with(mk, thb, True)
        b, _ = _simulate_with(mk, thb, False)
        assert np.array_equal(a, b), (mols, K, n_chan)
        from oracle import lte_oracle as O
        want = O.simulate(so, ocats, lidx, grid[0], thb[0], windowed=True)
        assert np.max(np.abs(a[0] - want)) <= 1e-5 * np.max(np.abs(want))


# ---------------------------------------------------------------------------------------------------------
# reach-ordered evaluation of plain log-prob batches: only the slot a row is evaluated at changes
def _with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


def test_reach_ordered_batches_give_bit_identical_log_probs(full_size_problem):
    """A spread-out ensemble in arbitrary order (dV over a factor of three, rows outside the bounds, a NaN row, a
    batch that is not a multiple of the block size, two chunks): forced on, forced off and adaptive must agree bit for
    bit, through device pointers and through host buffers; the adaptive engine must have switched itself on."""
    import torch
    prob = full_size_problem
    rng = np.random.default_rng(9)
    n = 16384 + 4500
    th = prob.walkers(n, seed=12)
    th[:, -1] *= rng.uniform(0.8, 2.4, n)                 # dV spread: reach classes differ inside every warp
    th[::97, -1] = -0.1                                   # out of bounds
    th[5, 1] = np.nan
    lo, hi = np.asarray(prob.spec.lo, dtype=float), np.asarray(prob.spec.hi, dtype=float)
    th[:, -1] = np.where(th[:, -1] >= hi[-1], 0.5 * (lo[-1] + hi[-1]), th[:, -1])
    d_th = torch.from_numpy(th).to("cuda:0")

    def run(via_device):
        def go():
            with prob.engine(precision="mixed") as eng:
                outs = []
                for _ in range(3):                            # the adaptive engine probes on its first eligible call
                    if via_device:
                        d_out = torch.empty(n, dtype=torch.float64, device="cuda:0")
                        eng.log_prob_device(d_th, out=d_out, with_prior=True, sync=True)
                        outs.append(d_out.cpu().numpy())
                    else:
                        outs.append(eng.log_prob(th))
                return outs, eng.stat("sorted_batches")
        return go
    for via_device in (True, False):
        off, n_off = _with_env("CHALTE_SORT_ROWS", "0", run(via_device))
        on, n_on = _with_env("CHALTE_SORT_ROWS", "1", run(via_device))
        auto, n_auto = _with_env("CHALTE_SORT_ROWS", "-1", run(via_device))
        assert n_off == 0 and n_on >= 3 and n_auto >= 2
        ref = off[0]
        assert np.isfinite(ref).sum() > 0.9 * n and not np.isfinite(ref[5]) and not np.isfinite(ref[0])
        for o in off[1:] + on + auto:
            assert np.array_equal(o, ref, equal_nan=True)


def test_reach_ordered_batches_with_a_state_sum_partition_function():
    """The state-sum partials are indexed by the caller's row, every other per-walker table by the evaluation slot:
    a molecule whose Q(T) is the explicit sum over states (1-cyanonapthalene), 5000 rows with a wide dV spread, forced
    reach ordering against plain order -- bit-identical -- and against the oracle on a few rows."""
    so, sp, ocats, pcats, grid, lidx, theta, stds = _small_problem(["1-cyanonapthalene"], 1, 2048, seed=11, ncol_scale=NCOL_SCALE["1-cyanonapthalene"])
    th = _ball(sp, theta, stds, 5000, seed=4, scale=0.3)
    rng = np.random.default_rng(3)
    th[:, -1] *= rng.uniform(0.7, 2.0, th.shape[0])
    th[:, sp.idx_tex] *= rng.uniform(0.9, 1.1, th.shape[0])       # Q(T) differs from row to row

    def run():
        with H.make_engine(sp, pcats, grid, None, prior=(stds, theta), precision="mixed") as eng:
            return eng.log_like(th), eng.stat("sorted_batches")
    plain, n0 = _with_env("CHALTE_SORT_ROWS", "0", run)
    ordered, n1 = _with_env("CHALTE_SORT_ROWS", "1", run)
    assert n0 == 0 and n1 >= 1
    assert np.array_equal(plain, ordered, equal_nan=True)
    co = _oracle_pair(so, ocats, grid, lidx, (stds, theta))
    assert np.isfinite(plain[:12]).any()
    _assert_ll(ordered[:12], co.lnlike(th[:12]), grid, "mixed", tag="reach-ordered state sum")
