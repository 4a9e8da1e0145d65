"""Host side with the reference's API: ``SpectralFitMCMC(config).run()`` (inference.py:63-488).

Same ``config`` dict (keys of inference.py:585-631), same method names and argument meaning
(``lnlike(theta, datagrid, mol_cat)``, ``lnprior(theta, prior_stds, prior_means)``, ``lnprob(...)``,
``is_within_bounds``, ``predict_intensities``, ``make_model``, ``init_setup``, ``estimate_Ncol_via_MLE``,
``fit_multi_gaussian``, ``run``), same on-disk artefacts (the reduced datagrid ``.npy`` and
``chain[_template].npy`` of shape (nwalkers, nsteps, ndim)).  New: ``log_prob(theta[nwalkers, ndim])``,
the emcee-compatible *vectorised* entry the GPU is built for.

Every number on the log-probability path is computed by libchalte.so on the B200; this module only
orchestrates (config, priors, walker ball, optimiser and sampler loops, files).  Optional config keys that the
reference does not have: 'device' (0), 'precision' ('mixed'|'fp64'), 'sampler' ('host'|'device'),
'save_every' (1 = the reference's save-after-every-step), 'seed' (None).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.optimize as opt

from .catalog import MolCat, find_catalog
from .constants import CYAN, GRAY, GREEN, RED, RESET
from .datagrid import calc_noise_std as _calc_noise_std
from .datagrid import reduce_spectrum, save_datagrid
from .engine import LTEEngine, ModelSpec
from .sampler import DeviceEnsembleSampler, EnsembleSampler

try:  # progress bar / table are cosmetic; both packages are in the image but not required
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **_):
        return it
try:
    from tabulate import tabulate
except Exception:  # pragma: no cover
    tabulate = None


class SpectralFitMCMC:
    def __init__(self, config):
        self.config = config
        self.mol_name = config['mol_name']
        self.template_run = config['template_run']
        self.fit_folder = config['fit_folder']
        self.cat_folder = config['cat_folder']
        self.data_path = config['data_paths'].get(self.mol_name)
        self.prior_path = config['prior_path']
        self.block_interlopers = config['block_interlopers']
        self.lower_limit = config['lower_limit']
        self.upper_limit = config['upper_limit']
        self.aligned_velocity = config['aligned_velocity']
        self.dish_size = config['dish_size']
        self.nwalkers = config['nwalkers']
        self.nruns = config['nruns']
        self.template_means = config['template_means']
        self.template_stds = config['template_stds']
        self.parallelize = config['parallelize']          # accepted for compatibility; the GPU replaces the Pool
        self.fixed_source_size = config['fixed_source_size']
        self.bounds = config['bounds']
        self.MLE_for_Ncol = config.get('MLE_for_Ncol', False)
        # extensions
        self.device = config.get('device', 0)
        self.precision = config.get('precision', 'mixed')
        self.sampler_kind = config.get('sampler', 'host')
        self.save_every = int(config.get('save_every', 1))
        self.seed = config.get('seed', None)

        if isinstance(self.fixed_source_size, (float, int)) and not isinstance(self.fixed_source_size, bool):
            self.source_size = self.fixed_source_size
            self.ndim = 4
            self.param_labels = ['Ncol [cm⁻²]', 'Tex [K]', 'vlsr [km s⁻¹]', 'dV [km s⁻¹]']
        else:
            self.source_size = None
            self.ndim = 5
            self.param_labels = ['Source Size [″]', 'Ncol [cm⁻²]', 'Tex [K]', 'vlsr [km s⁻¹]', 'dV [km s⁻¹]']
        self.param_labels_latex = list(self.param_labels)
        self.spec = ModelSpec.inference(self.source_size, self.bounds, self.dish_size, self.aligned_velocity,
                                        self.lower_limit, self.upper_limit)
        self._eng = None
        self._bound_grid = None
        self._bound_cat = None
        self._bound_prior = None

    # ------------------------------------------------------------------------------------------------
    # engine binding: the reference passes (datagrid, mol_cat, prior_stds, prior_means) on every call;
    # they are uploaded once and re-uploaded only when a different object is passed
    # ------------------------------------------------------------------------------------------------
    @property
    def engine(self) -> LTEEngine:
        if self._eng is None:
            self._eng = LTEEngine(device=self.device, precision=self.precision)
            self._eng.set_model(self.spec)
        return self._eng

    def bind(self, datagrid, mol_cat, prior_stds=None, prior_means=None):
        eng = self.engine
        if self._bound_cat is not mol_cat or self._bound_grid is not datagrid:
            eng.set_molecule(0, mol_cat, self.lower_limit, self.upper_limit,
                             line_idx=np.asarray(datagrid[3], dtype=np.int64))
            eng.set_spectrum(np.asarray(datagrid[0], float), np.asarray(datagrid[1], float), np.asarray(datagrid[2], float))
            self._bound_cat, self._bound_grid = mol_cat, datagrid
        if prior_stds is not None:
            key = (tuple(np.asarray(prior_stds, float)), tuple(np.asarray(prior_means, float)))
            if self._bound_prior != key:
                eng.set_prior(prior_stds, prior_means)
                self._bound_prior = key
        return eng

    def log_prob(self, theta, datagrid=None, mol_cat=None, prior_stds=None, prior_means=None):
        """Vectorised lnprob: theta[nwalkers, ndim] -> float64[nwalkers] (emcee ``vectorize=True`` contract)."""
        if datagrid is not None:
            self.bind(datagrid, mol_cat, prior_stds, prior_means)
        return self.engine.log_prob(theta)

    # ------------------------------------------------------------------------------------------------
    # reference-shaped scalar API
    # ------------------------------------------------------------------------------------------------
    def make_model(self, freqs, intensities, datagrid_freq, datagrid_ints, vlsr, dV, Tex, source_size):
        """inference.py:99-102 -> make_model_numba (44-61), evaluated on the device."""
        return self.engine.make_model(freqs, intensities, datagrid_freq, vlsr, dV, Tex, source_size,
                                      self.aligned_velocity, self.dish_size, mask_centre=0.0, planck_eps=1e-10)

    def calc_noise_std(self, intensity, threshold=3.5):
        return _calc_noise_std(intensity, threshold)

    def lnlike(self, theta, datagrid, mol_cat):
        """inference.py:127-166.  Exceptions are mapped to -inf like the reference (140-155)."""
        try:
            eng = self.bind(datagrid, mol_cat)
            return float(eng.log_like(np.asarray(theta, dtype=float)[None, :])[0])
        except Exception as e:  # noqa: BLE001 - mirror of the reference's catch-all
            print(f"{RED}Error in lnlike: {e}{RESET}")
            return -np.inf

    def is_within_bounds(self, theta):
        return self.spec.within_bounds(theta)

    def lnprior(self, theta, prior_stds, prior_means, weight=1.0):
        """inference.py:193-236 (evaluated by the device's prior kernel)."""
        self.engine.set_prior(prior_stds, prior_means)
        self._bound_prior = (tuple(np.asarray(prior_stds, float)), tuple(np.asarray(prior_means, float)))
        lp = float(self.engine.log_prior(np.asarray(theta, dtype=float)[None, :])[0])
        return lp if not np.isfinite(lp) else weight * lp

    def lnprob(self, theta, datagrid, mol_cat, prior_stds, prior_means):
        """inference.py:239-246."""
        eng = self.bind(datagrid, mol_cat, prior_stds, prior_means)
        return float(eng.log_prob(np.asarray(theta, dtype=float)[None, :])[0])

    def predict_intensities(self, Ncol, Tex, dV, mol_cat, source_size):
        """inference.py:249-253: (freq_sim, int_sim, tau_sim) of MolSim(gauss=False).  ObsParams there keeps its
        default dish_size=100 (classes.py:492), reproduced here."""
        eng = self.engine
        if self._bound_cat is not mol_cat:
            eng.set_molecule(0, mol_cat, self.lower_limit, self.upper_limit, line_idx=None)
            self._bound_cat, self._bound_grid = mol_cat, None
        return eng.stick_spectrum(0, mol_cat.frequency.size, Ncol, Tex, dV, source_size, 100)

    def read_file(self, filename, restfreqs, int_sim, shift=None, GHz=False, plot=False, block_interlopers=True):
        data = np.load(filename, allow_pickle=True)
        return reduce_spectrum(data[0], data[1], restfreqs, int_sim, self.aligned_velocity, shift=shift, GHz=GHz,
                               block_interlopers=block_interlopers, log=lambda m: print(f"{GRAY}{m}{RESET}"))

    def init_setup(self):
        """inference.py:305-342."""
        print(f"\n{CYAN}Reducing spectral data for {self.mol_name}.{RESET}")
        catfile_path = find_catalog(self.cat_folder, self.mol_name)
        os.makedirs(os.path.join(self.fit_folder, self.mol_name), exist_ok=True)
        if catfile_path is None:
            raise FileNotFoundError(f"{RED}No catalog file found at {os.path.join(self.cat_folder, self.mol_name + '.cat')}.{RESET}")
        source_size = self.source_size if self.source_size is not None else self.template_means[0]
        mol_cat = MolCat(self.mol_name, catfile_path)
        eng = self.engine
        eng.set_molecule(0, mol_cat, self.lower_limit, self.upper_limit, line_idx=None)
        self._bound_cat, self._bound_grid = mol_cat, None
        # MolSim(C=3.4e12, dV=0.89, T=7.0) with ObsParams(dish_size=self.dish_size): inference.py:322-327
        freq_sim, int_sim, _ = eng.stick_spectrum(0, mol_cat.frequency.size, 3.4e12, 7.0, 0.89, source_size, self.dish_size)
        print(f"{GRAY}Reading in spectral data from: {self.data_path}{RESET}")
        f, i, e, cov = self.read_file(self.data_path, freq_sim, int_sim, block_interlopers=self.block_interlopers)
        datafile_path = os.path.join(self.fit_folder, self.mol_name, "all_" + self.mol_name + "_lines_DSN_freq_space.npy")
        print(f"{GRAY}Saving reduced spectrum to: {datafile_path}{RESET}\n")
        save_datagrid(datafile_path, f, i, e, cov)
        return datafile_path, catfile_path

    def _ncol_theta(self, fixed_params):
        """theta rows with everything but the column density fixed (inference.py:348-352, 360-364)."""
        if self.source_size is not None:
            Tex, vlsr, dV = fixed_params
            return lambda N: np.column_stack([N, np.full_like(N, Tex), np.full_like(N, vlsr), np.full_like(N, dV)])
        source_size, Tex, vlsr, dV = fixed_params
        return lambda N: np.column_stack([np.full_like(N, source_size), N, np.full_like(N, Tex), np.full_like(N, vlsr),
                                          np.full_like(N, dV)])

    def ncol_profile(self, datagrid, mol_cat, fixed_params, n=256, bounds=None):
        """lnlike on a log-grid of n column densities inside `bounds` (default: the config's Ncol bounds) in ONE
        batched launch.  Returns (grid, lnlike[n])."""
        lo, hi = bounds if bounds is not None else (self.bounds['Ncol'][0], self.bounds['Ncol'][1])
        grid = np.geomspace(lo, hi, n + 2)[1:-1]
        return grid, self.bind(datagrid, mol_cat).log_like(self._ncol_theta(fixed_params)(grid))

    def estimate_Ncol_via_MLE(self, datagrid, mol_cat, fixed_params, n_profile=256):
        """inference.py:345-376: the column density that maximises lnlike with every other parameter fixed.

        The reference hands the whole box (1e8, 1e14) to scipy's bounded Brent, one lnlike call per probe.  Here ONE
        batched launch evaluates lnlike on a log-grid of `n_profile` column densities (`ncol_profile`); the grid
        maximum and its two neighbours bracket the optimum, and the same bounded Brent (same xatol) then runs on
        that bracket only.  The optimum is the reference's to the optimiser's own tolerance
        (tests/test_gpu_extended.py: 3.0209937550443867e12 on the HC5N fixture, 1e-6 relative)."""
        eng = self.bind(datagrid, mol_cat)
        mk = self._ncol_theta(fixed_params)

        def nll(Ncol):
            return -float(eng.log_like(mk(np.array([Ncol], dtype=float)))[0])

        Ncol_bounds = (self.bounds['Ncol'][0], self.bounds['Ncol'][1])
        bracket = Ncol_bounds
        self.mle_probes = 0
        try:
            if n_profile and n_profile >= 3:
                grid, ll = self.ncol_profile(datagrid, mol_cat, fixed_params, n=n_profile)
                self.mle_probes += 1
                if np.any(np.isfinite(ll)):
                    k = int(np.argmax(np.where(np.isfinite(ll), ll, -np.inf)))
                    bracket = (grid[k - 1] if k > 0 else Ncol_bounds[0],
                               grid[k + 1] if k + 1 < grid.size else Ncol_bounds[1])

            def counted(N):
                self.mle_probes += 1
                return nll(N)

            result = opt.minimize_scalar(counted, bounds=bracket, method='bounded', options={'xatol': 1e-6})
            if result.success:
                print(f"{GREEN}Succesful MLE fit for column density. Prior Ncol: {result.x:.3e}{RESET}")
                return result.x
            print(f"{RED}MLE for Ncol failed to converge.{RESET}")
            raise RuntimeError("MLE for Ncol did not converge.")
        except Exception as e:
            print(f"{RED}MLE for Ncol encountered an error: {e}{RESET}")
            raise

    # ------------------------------------------------------------------------------------------------
    def load_priors(self):
        """inference.py:388-419 incl. the (p16 + p84 - 2 p50)/2 'width' kept bug-for-bug (408)."""
        if self.template_run:
            initial = np.array(self.template_means, dtype=float)
            prior_means = initial
            prior_stds = np.array(self.template_stds, dtype=float)
            print(f"{GRAY}Using template priors and initial positions for {self.mol_name}.{RESET}")
            file_name = os.path.join(self.fit_folder, self.mol_name, "chain_template.npy")
        else:
            if not os.path.exists(self.prior_path):
                raise FileNotFoundError(f"{RED}The prior path {self.prior_path} could not be found.{RESET}")
            print(f"{GRAY}Loading previous chain data from: {self.prior_path}{RESET}")
            psamples = np.load(self.prior_path).T
            print(f"{GRAY}Dimensions of samples loaded from chain: {psamples.shape}{RESET}")
            prior_means = np.mean(np.percentile(psamples, 50, axis=1), axis=1)
            percentile_16 = np.percentile(psamples, 16, axis=1).mean(axis=1)
            percentile_84 = np.percentile(psamples, 84, axis=1).mean(axis=1)
            prior_stds = np.abs((percentile_16 - prior_means + percentile_84 - prior_means) / 2.0)
            initial = prior_means.copy()
            file_name = os.path.join(self.fit_folder, self.mol_name, "chain.npy")
        return initial, prior_means, prior_stds, file_name

    def fit_multi_gaussian(self, datafile, catalogue):
        """inference.py:379-473."""
        print(f"{CYAN}Estimating free parameters for {self.mol_name}.{RESET}")
        ndim = self.ndim
        if not os.path.exists(datafile):
            raise FileNotFoundError(f"{RED}The data file {datafile} could not be found.{RESET}")
        datagrid = np.load(datafile, allow_pickle=True)
        mol_cat = MolCat("mol", catalogue)
        if self.seed is not None:
            np.random.seed(self.seed)
        initial, prior_means, prior_stds, file_name = self.load_priors()
        initial = np.array(initial, dtype=float)

        if self.MLE_for_Ncol:
            print(f"{GRAY}Initializing Ncol via MLE.{RESET}")
            if self.source_size is not None:
                fixed_params = (prior_means[1], prior_means[2], prior_means[3])
            else:
                fixed_params = (prior_means[0], prior_means[2], prior_means[3], prior_means[4])
            try:
                estimated_Ncol = self.estimate_Ncol_via_MLE(datagrid, mol_cat, fixed_params)
                initial[0 if self.source_size is not None else 1] = estimated_Ncol
            except Exception:
                print(f"{RED}Failed to initialize Ncol via MLE. Exiting.{RESET}")
                return None

        # walker ball: inference.py:442-451 (global NumPy RNG, redraw until inside the bounds)
        pos, count = [], 0
        for _ in range(self.nwalkers):
            trial = None
            while trial is None or not self.is_within_bounds(trial):
                trial = initial + np.random.randn(ndim) * (prior_stds / 10.0)
                count += 1
            pos.append(trial)
        pos = np.array(pos)
        print(f"{GRAY}Failed walker initalizations: {self.nwalkers - count}{RESET}\n")

        self.bind(datagrid, mol_cat, prior_stds, prior_means)
        if self.sampler_kind == 'device':
            smp = DeviceEnsembleSampler(self.engine, self.nwalkers, pos, seed=int(self.seed or 0))
            # chains stay in HBM; the reference's checkpoint (np.save of the whole chain, inference.py:462/471) is
            # written every `save_every` steps instead of every step
            parts, done = [], 0
            every = max(1, self.save_every)
            while done < self.nruns:
                n = min(every, self.nruns - done)
                c, _ = smp.run(n, store_every=1)
                parts.append(c); done += n
                chain = np.concatenate(parts, axis=1)
                np.save(file_name, chain)
            return chain
        sampler = EnsembleSampler(self.nwalkers, ndim, self.engine.log_prob, vectorize=True)
        sampler.reserve(self.nruns)
        # The reference re-enters run_mcmc with a bare ndarray every step (inference.py:461-463), so emcee recomputes
        # the log-prob of every walker each step.  That call draws no random numbers and returns what the previous
        # step already holds, so passing log_prob0 gives the identical chain with half the evaluations.
        lp = None
        for step in tqdm(range(self.nruns), desc=f"MCMC Sampling for {self.mol_name}", colour='white'):
            pos, lp = sampler.run_mcmc(pos, 1, log_prob0=lp)
            if (step + 1) % self.save_every == 0 or step + 1 == self.nruns:
                np.save(file_name, sampler.chain)                     # the reference's checkpoint (462/471)
        return sampler.chain

    def run(self):
        datafile_path, catalogue_path = self.init_setup()
        chain = self.fit_multi_gaussian(datafile_path, catalogue_path)
        chain_path = os.path.join(self.fit_folder, self.mol_name, "chain_template.npy" if self.template_run else "chain.npy")
        if os.path.exists(chain_path):
            print_summary(chain_path, self.param_labels)
        else:
            print(f"{RED}Chain file not found at {chain_path}. Exiting.{RESET}")
        return chain


def posterior_summary(chain, burn_frac=0.2):
    """plot_results' table (inference.py:501-503, 565-581): drop the first 20 % of steps, flatten, 16/50/84
    percentiles per parameter -> rows of (median, lower, upper)."""
    chain = np.asarray(chain)
    burn_in = int(burn_frac * chain.shape[1])
    samples = chain[:, burn_in:, :].reshape((-1, chain.shape[-1]))
    out = []
    for i in range(samples.shape[1]):
        p = np.percentile(samples[:, i], [16, 50, 84])
        q = np.diff(p)
        out.append((p[1], q[0], q[1]))
    return np.array(out)


def print_summary(chain_path, param_labels):
    summ = posterior_summary(np.load(chain_path))
    table = []
    for label, (med, lo, hi) in zip(param_labels, summ):
        if abs(med) < 1e-3 or abs(med) > 1e3:
            table.append([label, f"{med:.2e}", f"{lo:.2e}", f"{hi:.2e}"])
        else:
            table.append([label, f"{med:.5f}", f"{lo:.5f}", f"{hi:.5f}"])
    headers = ["Parameter", "Median Estimate", "Lower Uncertainty", "Upper Uncertainty"]
    if tabulate is not None:
        print("\n" + tabulate(table, headers=headers, tablefmt="grid", colalign=["center"] * 4) + "\n")
    else:  # pragma: no cover
        for row in table:
            print(row)
    return summ
