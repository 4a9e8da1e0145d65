"""CPU suite: host-side logic of the product package (parser, Q dispatch, data reduction, samplers, grids)."""
import os

import numpy as np
import pytest

from tests import helpers as H


def test_constants_match_cuda_header_bit_for_bit():
    import re
    from cha1_mcmc_b200 import constants as K
    txt = open(os.path.join(H.ROOT, "cha1_mcmc_b200", "csrc", "lte_common.cuh")).read()
    hexes = dict(re.findall(r"constexpr double (k\w+)\s*=\s*(0x[0-9a-fA-F.]+p[+-]?\d+)", txt))
    assert float.fromhex(hexes["kKcm"]) == K.kcm
    assert float.fromhex(hexes["kCkm"]) == K.ckm
    assert float.fromhex(hexes["kCcm"]) == K.ccm == 29980000000.000004
    assert float.fromhex(hexes["kCm"]) == K.cm
    assert float.fromhex(hexes["kH"]) == K.h
    assert float.fromhex(hexes["kK"]) == K.k
    assert float.fromhex(hexes["kBeamConst"]) == 206265 * 1.22
    assert float.fromhex(hexes["kSijConst"]) == 4.16231 * 10 ** (-5)
    assert float.fromhex(hexes["kAijConst"]) == 1.16395 * 10 ** (-20)


def test_product_parser_against_reference_molcat():
    """cha1_mcmc_b200.MolCat vs the arrays the reference's MolCat produced (tests/golden/catalogs_ref.npz)."""
    g = np.load(H.GOLD + "/catalogs_ref.npz")
    from oracle import lte_oracle as O
    n_checked = 0
    for name in [str(n) for n in g["names"]]:
        if not os.path.exists(H.cat_path(name)):
            continue
        c = H.product_cat(name)
        assert c.frequency.size == int(g[f"{name}/N"]) and c.qns == int(g[f"{name}/qns"])
        assert np.array_equal(c.frequency, g[f"{name}/frequency"])
        assert np.array_equal(c.elower, g[f"{name}/elower"])
        assert np.array_equal(c.logint, g[f"{name}/logint"])
        assert np.array_equal(c.gup, g[f"{name}/gup"])
        assert np.array_equal(c.qn[:, 6:6 + c.qns].astype(np.int16), g[f"{name}/qn_lower"])
        qn = c.qn
        assert np.array_equal(np.array([qn.sum(), (qn * np.arange(1, 13)).sum(), (qn ** 2).sum()]), g[f"{name}/qn_digest"])
        # partition function data shipped to the device reproduces the reference's Q(T)
        Ts = g["Q_T"]
        if c.q_kind == 3:
            Q = [np.sum(c.state_g * np.exp(-c.state_E / (0.69503476 * T))) for T in Ts]
        elif c.q_kind == 1:
            a, b, s, d = c.q_params
            Q = [(a * T + b) / d if d else s * (a * T + b) for T in Ts]
        elif c.q_kind == 2:
            a, p, b, hb = c.q_params
            Q = [a * T ** p + (b if hb else 0.0) for T in Ts]
        else:
            Q = [sum(cf * T ** n for n, cf in enumerate(c.q_params)) for T in Ts]
        np.testing.assert_allclose(Q, g[f"{name}/Q"], rtol=1e-13)
        n_checked += 1
    assert n_checked == 35      # every shipped catalog travels with the tests


def test_q_dispatch_for_all_35_shipped_names():
    """Branch taken per shipped file name, incl. the three spelling traps."""
    from cha1_mcmc_b200 import resolve_q_mode
    g = np.load(H.GOLD + "/catalogs_ref.npz")
    state_sum = {"1-cyano-CPD", "1-cyanonapthalene", "2-cyano-CPD", "2-cyanonapthalene", "C10H-", "C8H-", "acenaphthylene",
                 "cyclopentadiene", "hc2nc", "indene", "indene_hfs"}
    for name in [str(n) for n in g["names"]]:
        kind, p = resolve_q_mode(f"/some/dir/{name}.cat")
        assert (kind == 3) == (name in state_sum), name
        if kind != 3:   # analytic Q agrees with the reference at every tabulated temperature
            Ts = g["Q_T"]
            if kind == 1:
                a, b, s, d = p
                Q = [(a * T + b) / d if d else s * (a * T + b) for T in Ts]
            else:
                a, pw, b, hb = p
                Q = [a * T ** pw + (b if hb else 0.0) for T in Ts]
            np.testing.assert_allclose(Q, g[f"{name}/Q"], rtol=1e-14)
    # dispatching on the full path reproduces the reference's directory-name trap
    assert resolve_q_mode("/x/hc3n_runs/benzonitrile.cat", dispatch_on="path")[0] == 1
    assert resolve_q_mode("/x/hc3n_runs/benzonitrile.cat")[0] == 2


def test_data_reduction_matches_reference_init_setup():
    """reduce_spectrum / calc_noise_std (row N3) on the DSN sample, with the reference's int_sim from the goldens."""
    from cha1_mcmc_b200.datagrid import reduce_spectrum
    from oracle import lte_oracle as O
    g = np.load(H.GOLD + "/hc5n_dsn_ref.npz")
    data = np.load(os.path.join(H.GOLD, "data", "cha_mms1_hc5n_example.npy"), allow_pickle=True)
    cat = H.oracle_cat("hc5n_hfs")
    f, t = O.line_taus(cat, 3.4e12, 7.0, 0.89, 18000, 25000)
    ints = O.stick_intensity(f, t, 7.0, 52.0, 70)
    rf, ri, re_, cov = reduce_spectrum(data[0], data[1], f, ints, 4.10)
    assert np.array_equal(cov, g["fixed/line_idx"])
    assert np.array_equal(rf, g["fixed/grid_freq"]) and np.array_equal(ri, g["fixed/grid_y"])
    np.testing.assert_allclose(re_, g["fixed/grid_yerr"], rtol=1e-13)
    # oracle's own restatement agrees too
    of, oi, oe, oc = O.reduce_spectrum(data[0], data[1], f, ints, 4.10)
    assert np.array_equal(oc, cov) and np.array_equal(of, rf)
    np.testing.assert_allclose(oe, re_, rtol=1e-13)


def test_window_grid_is_exact_size_sorted_and_covers_lines():
    from cha1_mcmc_b200.synthetic import window_grid
    rng = np.random.default_rng(0)
    lines = np.sort(rng.uniform(7000, 30000, 500))
    for n in (1 << 12, 50_000):
        x = window_grid(lines, n)
        assert x.size == n and np.all(np.diff(x) > 0)
    x = window_grid(lines, 1 << 14, v_centre=5.8)
    shifted = lines * (1 - 5.8 / 299800.0)
    near = np.abs(x[np.clip(np.searchsorted(x, shifted[:100]), 0, x.size - 1)] - shifted[:100])
    assert np.median(near) < 2e-3


def test_model_spec_layouts_and_prior_overrides():
    from cha1_mcmc_b200 import ModelSpec
    s = ModelSpec.inference(52.0, H.HC5N_BOUNDS, 70, 4.1, 18000, 25000)
    assert (s.ndim, s.idx_ss, s.idx_ncol, s.idx_tex, s.idx_vlsr, s.idx_dv) == (4, [-1], [[0]], 1, [2], 3)
    s5 = ModelSpec.inference(None, H.HC5N_BOUNDS, 70, 4.1, 18000, 25000)
    assert (s5.ndim, s5.idx_ss, s5.idx_ncol, s5.idx_tex, s5.idx_vlsr, s5.idx_dv) == (5, [0], [[1]], 2, [3], 4)
    mu, sd, gs = s5.effective_prior([6.5, 0.34e10, 3.0, 0.06, 0.22], [46.91, 3.4e10, 8.0, 4.3, 0.7575])
    assert sd[3] == 0.7575 * 0.8 and sd[4] == 0.7575 * 0.3 and list(gs) == [1, 0, 1, 1, 1]      # inference.py:221-222
    t = ModelSpec.tmc1(4, 1)
    assert t.ndim == 14 and t.idx_tex == 8 and t.idx_vlsr == [9, 10, 11, 12] and t.idx_dv == 13
    th = np.array([37, 25, 56, 22, 2.47e12, 11.19e12, 2.20e12, 5.64e12, 6.7, 5.624, 5.790, 5.910, 6.033, 0.117])
    assert t.within_bounds(th)
    bad = th.copy(); bad[9] = bad[10]            # vlsr ordering (TMC1:229)
    assert not t.within_bounds(bad)
    j = ModelSpec.tmc1(4, 2)
    assert j.ndim == 18 and j.idx_ncol == [[4, 5, 6, 7], [8, 9, 10, 11]] and j.idx_tex == 12


def test_host_ensemble_sampler_matches_emcee_restatement_in_oracle():
    """Product sampler vs the oracle's restatement of emcee's StretchMove: identical chains for identical RNG."""
    from cha1_mcmc_b200.sampler import EnsembleSampler
    from oracle import lte_oracle as O
    f = lambda x: -0.5 * np.sum((x / np.array([1.0, 2.0, 0.5])) ** 2, axis=1)      # noqa: E731
    np.random.seed(11)
    p0 = np.random.randn(16, 3)
    np.random.seed(5)
    s = EnsembleSampler(16, 3, f)
    rng = np.random.RandomState(); np.random.seed(5); rng.set_state(np.random.get_state())
    pos = p0.copy()
    c, lp = p0.copy(), f(p0)
    for _ in range(25):
        pos, _ = s.run_mcmc(pos, 1)
        c, lp, _ = O.stretch_move_step(c, f(c), f, rng)
        assert np.array_equal(pos, c)
    assert s.chain.shape == (16, 25, 3)                       # (nwalkers, nsteps, ndim) = chain.npy layout
    with pytest.raises(ValueError):
        EnsembleSampler(4, 3, f)                              # nwalkers < 2*ndim


def test_device_sampler_oracle_is_sharding_independent_and_samples_gaussian():
    from oracle import device_sampler_oracle as D
    f = lambda x: -0.5 * np.sum(x ** 2, axis=1)               # noqa: E731
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal((64, 4))
    c1, lp1, n1 = D.run(p0, f, 40, seed=123, shards=1)
    c4, lp4, n4 = D.run(p0, f, 40, seed=123, shards=4)
    assert np.array_equal(c1, c4) and np.array_equal(lp1, lp4) and n1 == n4
    c, _, nacc = D.run(p0, f, 600, seed=7)
    flat = c[200:].reshape(-1, 4)
    assert abs(flat.mean()) < 0.1 and abs(flat.std() - 1.0) < 0.1 and 0.2 < nacc / (600 * 64) < 0.9


def test_posterior_summary_matches_plot_results_table():
    from cha1_mcmc_b200.inference import posterior_summary
    rng = np.random.default_rng(1)
    chain = rng.normal(3.0, 0.5, size=(32, 500, 2))
    s = posterior_summary(chain)
    samples = chain[:, 100:, :].reshape(-1, 2)
    p = np.percentile(samples[:, 0], [16, 50, 84])
    assert s[0][0] == p[1] and s[0][1] == p[1] - p[0] and s[0][2] == p[2] - p[1]


def test_windowed_data_reduction_equals_the_full_scan():
    """reduce_spectrum brackets each line's +-1.5 km/s window by binary search on the sorted channel axis and builds
    the clipping mask by interval dilation; selection, noise and errors must equal the oracle's element-by-element
    restatement of the reference's full O(L*C) scan (inference.py:108-124, 256-303), on sorted and shuffled input,
    with interlopers strong enough to trigger the veto and the clipping."""
    from cha1_mcmc_b200.datagrid import calc_noise_std, reduce_spectrum
    from oracle import lte_oracle as O
    rng = np.random.default_rng(2)
    freqs = np.sort(rng.uniform(18000, 18200, 20000))
    lines = np.sort(rng.uniform(18005, 18195, 40))
    inten = rng.normal(0, 0.01, freqs.size)
    for f in lines[::2]:
        inten += 0.05 * np.exp(-0.5 * ((freqs - f * (1 - 4.1 / 299800.0)) / 0.02) ** 2)
    int_sim = rng.uniform(0.0, 1.0, lines.size)
    for block in (True, False):
        a = reduce_spectrum(freqs, inten, lines, int_sim, 4.1, block_interlopers=block)
        want = O.reduce_spectrum(freqs, inten, lines, int_sim, 4.1, block_interlopers=block)
        assert a[0].size > 0 and np.array_equal(a[3], want[3])
        assert np.array_equal(a[0], want[0]) and np.array_equal(a[1], want[1])
        np.testing.assert_allclose(a[2], want[2], rtol=1e-13)
        perm = rng.permutation(freqs.size)
        b = reduce_spectrum(freqs[perm], inten[perm], lines, int_sim, 4.1, block_interlopers=block)
        wb = O.reduce_spectrum(freqs[perm], inten[perm], lines, int_sim, 4.1, block_interlopers=block)
        assert np.array_equal(b[3], wb[3]) and np.array_equal(b[0], wb[0]) and np.array_equal(b[1], wb[1])
        np.testing.assert_allclose(b[2], wb[2], rtol=1e-13)
    # a shifted window (read_file's `shift` argument) selects the same channels in either input order
    a = reduce_spectrum(freqs, inten, lines, int_sim, 4.1, shift=4.5)
    perm = rng.permutation(freqs.size)
    b = reduce_spectrum(freqs[perm], inten[perm], lines, int_sim, 4.1, shift=4.5)
    o = np.argsort(b[0])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[0], b[0][o]) and np.array_equal(a[1], b[1][o])
    # clipped noise estimate: outliers at the edges, clusters, all-outlier and empty inputs
    for trial in range(40):
        v = rng.normal(0, 1, rng.integers(1, 60))
        k = rng.integers(0, 4)
        v[rng.integers(0, v.size, k)] += rng.choice([-30, 30], k)
        got, want = calc_noise_std(v), O.calc_noise_std(v)
        np.testing.assert_allclose(got, want, rtol=1e-13, equal_nan=True)
    assert all(np.isnan(calc_noise_std(np.array([]))))


def test_survey_helpers_grid_sharding_and_catalog_list():
    """Config 5 plumbing that needs no GPU: +-1.5 km/s window grid against a brute-force union, the cost-balanced
    assignment of whole fits to ranks (deterministic, complete, balanced), and the shipped catalog list."""
    from cha1_mcmc_b200 import survey as SV
    from cha1_mcmc_b200.constants import ckm
    from cha1_mcmc_b200.synthetic import default_cat_folder
    rng = np.random.default_rng(3)
    lines = np.sort(rng.uniform(18000, 25000, 40)); lines[5] = lines[4] + 0.01; lines[9] = lines[8]      # overlapping windows
    for dnu, vc in ((30.518e-3, 0.2), (1.4e-3, 5.8)):
        got = SV.velocity_window_grid(lines, 1.5, dnu, vc)
        want = set()
        for f in lines:
            c = f * (1 - vc / ckm); h = 1.5 / ckm * f
            want.update(range(int(np.ceil((c - h) / dnu)), int(np.floor((c + h) / dnu)) + 1))
        assert np.array_equal(got, np.array(sorted(want), dtype=float) * dnu)
        assert np.all(np.diff(got) > 0)
    assert SV.velocity_window_grid([], 1.5, 1e-3, 0.0).size == 0
    costs = list(rng.uniform(1, 100, 70)) + [1000.0]
    for world in (1, 2, 4, 8):
        parts = SV.shard_fits(costs, world)
        assert parts == SV.shard_fits(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        load = [sum(costs[i] for i in p) for p in parts]
        assert max(load) <= max(sum(costs) / world * 1.05, 1000.0 + 1e-9)
    # divisible fits: every walker of every fit on exactly one rank, pieces on block boundaries, loads within 10 %
    for world in (1, 2, 4, 8):
        parts = SV.shard_fit_walkers(costs, world, 3744)
        assert parts == SV.shard_fit_walkers(costs, world, 3744)
        seen = {}
        for r, p in enumerate(parts):
            for i, a, b in p:
                assert 0 <= a < b <= 3744 and (a % 128 == 0) and (b % 128 == 0 or b == 3744)
                seen.setdefault(i, []).append((a, b))
        assert sorted(seen) == list(range(len(costs)))
        for i, segs in seen.items():
            segs.sort()
            assert segs[0][0] == 0 and segs[-1][1] == 3744 and all(x[1] == y[0] for x, y in zip(segs[:-1], segs[1:]))
        load = [sum(costs[i] * (b - a) / 3744 for i, a, b in p) for p in parts]
        assert max(load) <= sum(costs) / world * 1.10
    mols = SV.list_molecules(default_cat_folder())
    assert len(mols) == 35 and "benzonitrile" in mols and "1-cyanonapthalene" in mols


def test_walker_ball_refuses_a_centre_outside_the_bounds():
    """inference.py:442-451 redraws walkers until they are inside the bounds -- which never ends when theta* itself
    is outside; the synthetic-problem helper must say so instead of spinning."""
    from cha1_mcmc_b200 import ModelSpec
    from cha1_mcmc_b200.synthetic import SyntheticProblem
    bounds = {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e14], 'Tex': [3.5, 12.0], 'vlsr': [3.0, 5.5], 'dV': [0.4, 1.5]}
    spec = ModelSpec.inference(52.0, bounds, 70, 4.10, 18000, 25000)
    theta = np.array([3e12, 8.0, 4.3, 0.75])
    ok = SyntheticProblem("t", spec, [], [], np.zeros(1), np.zeros(1), np.ones(1), theta, theta, np.array([1e11, 0.3, 0.01, 0.01]))
    w = ok.walkers(64, seed=3)
    assert w.shape == (64, 4) and all(spec.within_bounds(t) for t in w)
    bad = SyntheticProblem("t", spec, [], [], np.zeros(1), np.zeros(1), np.ones(1), np.array([3e14, 8.0, 4.3, 0.75]), theta,
                           np.array([1e11, 0.3, 0.01, 0.01]))
    with pytest.raises(ValueError):
        bad.walkers(4)
