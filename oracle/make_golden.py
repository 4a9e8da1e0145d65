"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build
container (python oracle/make_golden.py).  Test infrastructure only.

The reference ships no tests / golden vectors; these files
pin the oracle (oracle/lte_oracle.py, oracle/lte_oracle.c) and through it the
CUDA path.  /root/reference does not travel to the GPU box, the .npz do.

Also copies the catalog/data fixtures the tests and bench need (public CDMS
catalogs and the reference's sample spectra: input DATA, not source code) into
tests/golden/catalog/*.cat.gz and tests/golden/data/.
"""
import contextlib
import gzip
import io
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

REF = ref_shim.REF_ROOT
CATDIR = os.path.join(REF, "catalog")

# every shipped catalog travels (gzip, 1.9 MB in all): parser tests, every BASELINE config incl. config 5 (all 35 molecules)
SHIP_CATS = sorted(f[:-4] for f in os.listdir(CATDIR) if f.endswith(".cat"))


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ship_fixtures():
    os.makedirs(os.path.join(GOLD, "catalog"), exist_ok=True)
    os.makedirs(os.path.join(GOLD, "data"), exist_ok=True)
    for name in SHIP_CATS:
        src = os.path.join(CATDIR, name + ".cat")
        dst = os.path.join(GOLD, "catalog", name + ".cat.gz")
        with open(src, "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    shutil.copy(os.path.join(REF, "data", "DSN", "cha_mms1_hc5n_example.npy"), os.path.join(GOLD, "data"))


def golden_catalogs(ref):
    """Per shipped catalog (all 35): what the reference MolCat computes + MolSim taus."""
    out = {}
    names = sorted(f[:-4] for f in os.listdir(CATDIR) if f.endswith(".cat"))
    Ts = np.array([3.0, 5.0, 8.0, 12.0, 60.0, 300.0])
    for name in names:
        path = os.path.join(CATDIR, name + ".cat")
        if name in ("1-cyanonapthalene",):
            pass  # 36 s in the reference (O(N^2) glow); still do it once
        with quiet():
            cat = ref.classes.MolCat(name, path)
            q = np.array([ref.functions.calc_q(cat, float(T)) for T in Ts])
            obs = ref.classes.ObsParams("g", source_size=40)
            sim = ref.classes.MolSim("g", cat, obs, vlsr=[0.0], C=[1.0e12], dV=[0.3], T=[7.0],
                                     ll=[7000], ul=[30000], gauss=False)
        qn = np.stack([np.asarray(getattr(cat, f"qn{i}"), dtype=np.int64) for i in range(1, 13)], axis=1)
        out[f"{name}/N"] = np.int64(cat.frequency.size)
        out[f"{name}/qns"] = np.int64(cat.qns)
        out[f"{name}/Q"] = q
        out[f"{name}/frequency"] = np.asarray(cat.frequency, dtype=float)
        out[f"{name}/elower"] = np.asarray(cat.elower, dtype=float)
        out[f"{name}/logint"] = np.asarray(cat.logint, dtype=float)
        out[f"{name}/gup"] = np.asarray(cat.gup, dtype=np.int64)
        out[f"{name}/qn_digest"] = np.array([qn.sum(), (qn * np.arange(1, 13)).sum(), (qn ** 2).sum()], dtype=np.int64)
        out[f"{name}/qn_lower"] = qn[:, 6:6 + int(cat.qns)].astype(np.int16)
        out[f"{name}/sijmu"] = np.asarray(cat.sijmu, dtype=float)
        out[f"{name}/freq_sim"] = np.asarray(sim.freq_sim, dtype=float)
        out[f"{name}/tau_sim"] = np.asarray(sim.tau_sim, dtype=float)
        out[f"{name}/int_sim"] = np.asarray(sim.int_sim, dtype=float)
        print("catalog", name, cat.frequency.size, flush=True)
    out["names"] = np.array(names)
    out["Q_T"] = Ts
    np.savez_compressed(os.path.join(GOLD, "catalogs_ref.npz"), **out)


def hc5n_config(ref, tmp, fixed=52.0):
    cfg = {
        'mol_name': 'hc5n_hfs', 'template_run': True, 'nruns': 10, 'nwalkers': 128,
        'bounds': {'source_size': [30.0, 90.0], 'Ncol': [1e8, 1e14], 'Tex': [3.5, 12.0],
                   'vlsr': [3.0, 5.5], 'dV': [0.4, 1.5]},
        'template_means': np.array([46.91, 3.4e10, 8.0, 4.3, 0.7575]),
        'template_stds': np.array([6.5, 0.34e10, 3.0, 0.06, 0.22]),
        'dish_size': 70, 'lower_limit': 18000, 'upper_limit': 25000, 'aligned_velocity': 4.10,
        'fixed_source_size': fixed, 'MLE_for_Ncol': True, 'block_interlopers': True, 'parallelize': False,
        'fit_folder': tmp, 'cat_folder': CATDIR, 'prior_path': os.path.join(tmp, 'none.npy'),
        'data_paths': {'hc5n_hfs': os.path.join(REF, 'data', 'DSN', 'cha_mms1_hc5n_example.npy')},
    }
    if isinstance(fixed, (float, int)):
        cfg['template_means'] = cfg['template_means'][1:]
        cfg['template_stds'] = cfg['template_stds'][1:]
    return cfg


def draw_box(rng, lo, hi, n, frac_out=0.1):
    lo = np.asarray(lo, float); hi = np.asarray(hi, float)
    th = lo + (hi - lo) * rng.random((n, lo.size))
    # a few rows outside the box to pin the -inf prior
    nout = int(frac_out * n)
    for r in range(nout):
        j = rng.integers(lo.size)
        th[r, j] = hi[j] * 1.01 if rng.random() < 0.5 else lo[j] * 0.99
    return th


def golden_hc5n(ref):
    out = {}
    for tag, fixed in (("fixed", 52.0), ("free", None)):
        tmp = tempfile.mkdtemp()
        cfg = hc5n_config(ref, tmp, fixed)
        with quiet():
            fit = ref.inference.SpectralFitMCMC(cfg)
            datafile, catfile = fit.init_setup()
            dg = np.load(datafile, allow_pickle=True)
            cat = ref.classes.MolCat("mol", catfile)
        b = cfg['bounds']
        names = ["Ncol", "Tex", "vlsr", "dV"] if fixed is not None else ["source_size", "Ncol", "Tex", "vlsr", "dV"]
        lo = [b[n][0] for n in names]; hi = [b[n][1] for n in names]
        rng = np.random.default_rng(20260101 + (fixed is None))
        th = draw_box(rng, lo, hi, 200)
        # log-uniform Ncol over the realistic range instead of uniform over [1e8, 1e14]
        jn = names.index("Ncol")
        inb = (th[:, jn] > lo[jn]) & (th[:, jn] < hi[jn])
        th[inb, jn] = 10 ** rng.uniform(10.5, 13.5, inb.sum())
        th0 = np.array([3.4e12, 8.0, 4.3, 0.7575]) if fixed is not None else np.array([46.91, 3.4e12, 8.0, 4.3, 0.7575])
        th = np.vstack([th0, th])
        mu, sd = cfg['template_means'], cfg['template_stds']
        with quiet():
            ll = np.array([fit.lnlike(t, dg, cat) for t in th])
            lp = np.array([fit.lnprior(t, sd, mu) for t in th])
            lpr = np.array([fit.lnprob(t, dg, cat, sd, mu) for t in th])
            models = []
            for t in th[:16]:
                if fixed is not None:
                    N, T, v, d = t; ss = fixed
                else:
                    ss, N, T, v, d = t
                fr, it, ta = fit.predict_intensities(Ncol=N, Tex=T, dV=d, mol_cat=cat, source_size=ss)
                li = dg[3]
                models.append(fit.make_model(freqs=np.array(fr)[li], intensities=np.array(ta)[li], datagrid_freq=dg[0],
                                             datagrid_ints=dg[1], vlsr=v, dV=d, Tex=T, source_size=ss))
            if fixed is not None:
                mle = fit.estimate_Ncol_via_MLE(dg, cat, (mu[1], mu[2], mu[3]))
            else:
                mle = fit.estimate_Ncol_via_MLE(dg, cat, (mu[0], mu[2], mu[3], mu[4]))
        out[f"{tag}/grid_freq"] = np.asarray(dg[0], float); out[f"{tag}/grid_y"] = np.asarray(dg[1], float)
        out[f"{tag}/grid_yerr"] = np.asarray(dg[2], float); out[f"{tag}/line_idx"] = np.asarray(dg[3], np.int64)
        out[f"{tag}/theta"] = th; out[f"{tag}/lnlike"] = ll; out[f"{tag}/lnprior"] = lp; out[f"{tag}/lnprob"] = lpr
        out[f"{tag}/models"] = np.array(models); out[f"{tag}/mle_ncol"] = np.float64(mle)
        out[f"{tag}/prior_means"] = np.asarray(mu, float); out[f"{tag}/prior_stds"] = np.asarray(sd, float)
        out[f"{tag}/lo"] = np.array(lo); out[f"{tag}/hi"] = np.array(hi)
        print("hc5n", tag, "C", dg[0].size, "L", len(dg[3]), "lnlike0", repr(ll[0]), "lnprior0", repr(lp[0]),
              "mle", repr(mle), flush=True)
    np.savez_compressed(os.path.join(GOLD, "hc5n_dsn_ref.npz"), **out)


TMC1_INITIAL = np.array([37, 25, 56, 22, 2.47e12, 11.19e12, 2.20e12, 5.64e12, 6.7, 5.624, 5.790, 5.910, 6.033, 0.117])
TMC1_STDS = np.array([2.5, 2.0, 6.5, 2.0, 0.30e12, 1.75e12, 0.265e12, 1.185e12, 0.1, 0.0015, 0.001, 0.0035, 0.002, 0.002])


def golden_tmc1(ref):
    """4-component script on the shipped GOTHAM windows.  The script's own
    init_setup (its data reduction is OUT OF SCOPE) produces the datagrid; the
    datagrid itself is stored so the tests never re-run that reduction."""
    t4 = ref_shim.load_tmc1()
    out = {}
    for mol, scale in (("hc9n_hfs", 1.0), ("hc7n_hfs", 1.0), ("hc11n", 1.0), ("benzonitrile", 0.1)):
        tmp = tempfile.mkdtemp()
        with quiet():
            datafile, catfile = t4.init_setup(tmp, CATDIR, os.path.join(REF, "data", "GOTHAM", f"{mol}_chunks.npy"),
                                              mol, True)
            dg = np.load(datafile, allow_pickle=True)
            cat = ref.classes.MolCat(mol, catfile)
        th0 = TMC1_INITIAL.copy(); th0[4:8] *= scale
        rng = np.random.default_rng(7)
        n = 48
        th = th0 + rng.standard_normal((n, 14)) * (TMC1_STDS * np.r_[np.ones(4), np.full(4, scale), np.ones(6)]) * 0.5
        # wider excursions in dV/Tex/vlsr so the shifted mask (5.8 km/s) actually bites
        th[n // 2:, 13] = rng.uniform(0.02, 0.29, n - n // 2)
        th[n // 2:, 8] = rng.uniform(3.0, 15.0, n - n // 2)
        th[:4, 9] = th[:4, 10] + 0.01          # violates ordering -> -inf prior
        th = np.vstack([th0, th])
        with quiet(), np.errstate(all="ignore"):
            ll = np.array([t4.lnlike(t, dg, cat) for t in th])
            lp = np.array([t4.lnprior(t, TMC1_STDS, TMC1_INITIAL) for t in th])
            lpr = np.array([t4.lnprob(t, dg, cat, TMC1_STDS, TMC1_INITIAL) for t in th])
        out[f"{mol}/grid_freq"] = np.asarray(dg[0], float); out[f"{mol}/grid_y"] = np.asarray(dg[1], float)
        out[f"{mol}/grid_yerr"] = np.asarray(dg[2], float); out[f"{mol}/line_idx"] = np.asarray(dg[3], np.int64)
        out[f"{mol}/theta"] = th; out[f"{mol}/lnlike"] = ll; out[f"{mol}/lnprior"] = lp; out[f"{mol}/lnprob"] = lpr
        print("tmc1", mol, "C", dg[0].size, "L", len(dg[3]), "lnlike0", repr(ll[0]), "lnprior0", repr(lp[0]), flush=True)
    out["prior_means"] = TMC1_INITIAL; out["prior_stds"] = TMC1_STDS
    np.savez_compressed(os.path.join(GOLD, "tmc1_gotham_ref.npz"), **out)


def golden_synth(ref):
    """Small synthetic GOTHAM-like grid for benzonitrile (scaled down): the
    reference's own lnlike (inference.py layout, fixed and free source size) on it."""
    tmp = tempfile.mkdtemp()
    cfg = hc5n_config(ref, tmp, 40.0)
    cfg.update({'mol_name': 'benzonitrile', 'dish_size': 100, 'lower_limit': 7000, 'upper_limit': 30000,
                'aligned_velocity': 5.8,
                'bounds': {'source_size': [0.0, 200.0], 'Ncol': [1e8, 1e14], 'Tex': [2.7, 15.0],
                           'vlsr': [5.0, 6.6], 'dV': [0.05, 0.3]}})
    out = {}
    with quiet():
        fit = ref.inference.SpectralFitMCMC(cfg)
        cat = ref.classes.MolCat("benzonitrile", os.path.join(CATDIR, "benzonitrile.cat"))
        fr, it, ta = fit.predict_intensities(Ncol=1.7e11, Tex=6.7, dV=0.117, mol_cat=cat, source_size=40.0)
    fr = np.array(fr)
    rng = np.random.default_rng(0)
    # windows of +-24 channels at 1.4 kHz around 160 randomly chosen lines (shifted to vlsr 5.8)
    pick = np.sort(rng.choice(fr.size, 160, replace=False))
    dnu = 1.4e-3
    chans = []
    for f0 in fr[pick]:
        fc = f0 * (1 - 0.0 / 299800.0)
        chans.append(fc + dnu * np.arange(-24, 25))
    x = np.unique(np.round(np.concatenate(chans) / dnu) * dnu)
    line_idx = np.arange(fr.size)            # every line in (ll, ul] participates
    theta_true = np.array([1.7e11, 6.7, 5.8 + 0.02, 0.117])
    dg0 = np.array([x, np.zeros_like(x), np.ones_like(x), line_idx], dtype=object)
    with quiet():
        m_true = fit.make_model(freqs=fr[line_idx], intensities=np.array(ta)[line_idx], datagrid_freq=x,
                                datagrid_ints=dg0[1], vlsr=theta_true[2], dV=0.117, Tex=6.7, source_size=40.0)
    y = m_true + rng.normal(0, 0.005, x.size)
    yerr = np.sqrt(0.005 ** 2 + (0.1 * y) ** 2)
    dg = np.array([x, y, yerr, line_idx], dtype=object)
    for tag, fixed in (("fixed", 40.0), ("free", None)):
        cfg2 = dict(cfg); cfg2['fixed_source_size'] = fixed
        with quiet():
            fit = ref.inference.SpectralFitMCMC(cfg2)
        names = ["Ncol", "Tex", "vlsr", "dV"] if fixed is not None else ["source_size", "Ncol", "Tex", "vlsr", "dV"]
        b = cfg['bounds']
        lo = [b[n][0] for n in names]; hi = [b[n][1] for n in names]
        th = draw_box(rng, lo, hi, 40)
        jn = names.index("Ncol")
        inb = (th[:, jn] > lo[jn]) & (th[:, jn] < hi[jn])
        th[inb, jn] = 10 ** rng.uniform(10.5, 12.5, inb.sum())
        t0 = theta_true if fixed is not None else np.r_[40.0, theta_true]
        ball = t0 + rng.standard_normal((24, len(t0))) * np.array(([4.0] if fixed is None else []) + [2e10, 0.3, 0.01, 0.01])
        th = np.vstack([t0, ball, th])
        mu = t0; sd = np.array(([4.0] if fixed is None else []) + [2e10, 0.3, 0.01, 0.01])
        with quiet():
            ll = np.array([fit.lnlike(t, dg, cat) for t in th])
            lpr = np.array([fit.lnprob(t, dg, cat, sd, mu) for t in th])
        out[f"{tag}/theta"] = th; out[f"{tag}/lnlike"] = ll; out[f"{tag}/lnprob"] = lpr
        out[f"{tag}/prior_means"] = mu; out[f"{tag}/prior_stds"] = sd
        out[f"{tag}/lo"] = np.array(lo); out[f"{tag}/hi"] = np.array(hi)
        print("synth", tag, "C", x.size, "lnlike0", repr(ll[0]), flush=True)
    out["grid_freq"] = x; out["grid_y"] = y; out["grid_yerr"] = yerr; out["line_idx"] = line_idx
    out["model_true"] = m_true
    np.savez_compressed(os.path.join(GOLD, "benzonitrile_synth_ref.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fixtures", "hc5n", "tmc1", "synth", "catalogs"]
    ref = ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    if "fixtures" in which:
        ship_fixtures()
    if "hc5n" in which:
        golden_hc5n(ref)
    if "tmc1" in which:
        golden_tmc1(ref)
    if "synth" in which:
        golden_synth(ref)
    if "catalogs" in which:
        golden_catalogs(ref)
