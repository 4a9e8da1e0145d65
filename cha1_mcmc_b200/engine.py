"""ctypes binding of libchalte.so (include/chalte.h) -- the thin host layer between the
reference-shaped Python API and the sm_100a kernels.

There is no CPU implementation behind this module: if the shared library is missing or no
B200 is visible, construction raises.  PyTorch is used only as plumbing for device-resident
tensors (``log_prob_device``); NumPy buffers go through the library's own pinned staging.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .catalog import MolCat, Q_SUM

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libchalte.so")

PREC_FP64, PREC_MIXED = 0, 1
_PREC = {"fp64": PREC_FP64, "f64": PREC_FP64, "mixed": PREC_MIXED, PREC_FP64: PREC_FP64, PREC_MIXED: PREC_MIXED}

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)

# every exported symbol with its signature (also used by tests/test_abi.py)
SIGNATURES = {
    "cha_version": (C.c_int, []),
    "cha_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "cha_destroy": (C.c_int, [C.c_void_p]),
    "cha_last_error": (C.c_char_p, [C.c_void_p]),
    "cha_set_molecule": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _dp, _dp, _dp, C.c_int, _dp, C.c_int,
                                   C.c_int64, _dp, _dp, C.c_double, C.c_double, _lp, C.c_int64]),
    "cha_set_spectrum": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp]),
    "cha_set_model": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, C.c_int,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "cha_set_prior": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _ip, C.c_double, C.c_double]),
    "cha_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "cha_log_prob": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp]),
    "cha_log_like": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp]),
    "cha_log_prior": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp]),
    "cha_simulate": (C.c_int, [C.c_void_p, _dp, C.c_int64, _dp]),
    "cha_log_prob_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "cha_simulate_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cha_sync": (C.c_int, [C.c_void_p]),
    "cha_stream": (C.c_void_p, [C.c_void_p]),
    "cha_sampler_init": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_uint64, C.c_double]),
    "cha_sampler_half_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "cha_sampler_coords_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cha_sampler_get": (C.c_int, [C.c_void_p, _dp, _dp, _lp]),
    "cha_comm_unique_id": (C.c_int, [C.c_char_p]),
    "cha_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "cha_comm_destroy": (C.c_int, [C.c_void_p]),
    "cha_sampler_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "cha_sampler_chain_len": (C.c_int64, [C.c_void_p]),
    "cha_sampler_chain_read": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _dp, _dp]),
    "cha_sampler_chain_clear": (C.c_int, [C.c_void_p]),
    "cha_stat": (C.c_int64, [C.c_void_p, C.c_int]),
    "cha_stick_spectrum": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                     _dp, _dp, _dp, _lp]),
    "cha_make_model": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, C.c_int64, _dp, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp]),
    "cha_count_window_pairs": (C.c_int, [C.c_void_p, _dp, C.c_int64, _lp]),
}


def load_library(path: str = None):
    """dlopen libchalte.so and declare every signature.  Fails loudly when it is missing.
    CHALTE_LIB names an alternative build of the same sources (kernel experiments)."""
    path = path or os.environ.get("CHALTE_LIB", LIB_PATH)
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension has not been built "
            "(python -m cha1_mcmc_b200.build).  This engine has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray, typ=_dp):
    return a.ctypes.data_as(typ)


@dataclass
class ModelSpec:
    """theta layout + telescope + masks of one fit (K velocity components, M molecules).

    inference.py:133-137 (4-/5-dim) and scripts/MCMC/TMC1_four_component.py:189 (14-dim) are the two
    layouts the reference ships; the joint M-molecule layout is the composition defined in
    SURVEY.md 8(d) config 4."""
    ndim: int
    K: int
    idx_ss: Sequence[int]
    idx_ncol: Sequence[Sequence[int]]     # [M][K]
    idx_tex: int
    idx_vlsr: Sequence[int]
    idx_dv: int
    fixed_ss: float = float("nan")
    dish_size: float = 100.0
    aligned_velocity: float = 0.0
    mask_centre: float = 0.0              # 0: inference.py:52 ; 5.8: TMC1_four_component.py:160
    planck_eps: float = 1e-10             # inference.py:56-57 ; 0 for TMC1_four_component.py:168-169
    ll: float = 7000.0
    ul: float = 30000.0
    lo: Optional[np.ndarray] = None
    hi: Optional[np.ndarray] = None
    vlsr_min_sep: float = float("nan")
    vlsr_max_sep: float = float("nan")
    param_names: List[str] = field(default_factory=list)

    @property
    def M(self) -> int:
        return len(self.idx_ncol)

    @property
    def ncol_indices(self):
        return sorted({i for row in self.idx_ncol for i in row})

    @staticmethod
    def inference(fixed_source_size, bounds, dish_size, aligned_velocity, lower_limit, upper_limit) -> "ModelSpec":
        """Layouts and strict box bounds of inference.py:133-137, 169-190."""
        if isinstance(fixed_source_size, (float, int)) and not isinstance(fixed_source_size, bool):
            names = ["Ncol", "Tex", "vlsr", "dV"]
            s = ModelSpec(ndim=4, K=1, idx_ss=[-1], idx_ncol=[[0]], idx_tex=1, idx_vlsr=[2], idx_dv=3,
                          fixed_ss=float(fixed_source_size))
        else:
            names = ["source_size", "Ncol", "Tex", "vlsr", "dV"]
            s = ModelSpec(ndim=5, K=1, idx_ss=[0], idx_ncol=[[1]], idx_tex=2, idx_vlsr=[3], idx_dv=4)
        s.param_names = names
        s.dish_size = float(dish_size); s.aligned_velocity = float(aligned_velocity)
        s.mask_centre = 0.0; s.planck_eps = 1e-10
        s.ll = float(lower_limit); s.ul = float(upper_limit)
        s.lo = np.array([bounds[n][0] for n in names], dtype=float)
        s.hi = np.array([bounds[n][1] for n in names], dtype=float)
        return s

    @staticmethod
    def tmc1(K: int = 4, n_mol: int = 1) -> "ModelSpec":
        """[ss_1..K, Ncol(mol 0)_1..K, ..., Tex, vlsr_1..K, dV] with the bounds of
        TMC1_four_component.py:224-233 (dish 100 m, window 7-30 GHz, mask centre 5.8 km/s)."""
        ndim = K + n_mol * K + 1 + K + 1
        s = ModelSpec(ndim=ndim, K=K, idx_ss=list(range(K)),
                      idx_ncol=[[K + m * K + c for c in range(K)] for m in range(n_mol)],
                      idx_tex=K + n_mol * K, idx_vlsr=[K + n_mol * K + 1 + c for c in range(K)],
                      idx_dv=K + n_mol * K + 1 + K)
        s.dish_size = 100.0; s.aligned_velocity = 0.0; s.mask_centre = 5.8; s.planck_eps = 0.0
        s.ll = 7000.0; s.ul = 30000.0
        lo = np.full(ndim, -np.inf); hi = np.full(ndim, np.inf)
        lo[:K] = 0.0; hi[:K] = 200.0
        lo[K:K + n_mol * K] = 0.0; hi[K:K + n_mol * K] = 10 ** 16.
        lo[s.idx_tex] = 2.7
        hi[s.idx_dv] = 0.3
        s.lo, s.hi = lo, hi
        s.vlsr_min_sep, s.vlsr_max_sep = 0.05, 0.3
        s.param_names = ([f"source_size{c + 1}" for c in range(K)]
                         + [f"Ncol{c + 1}" + (f"_m{m}" if n_mol > 1 else "") for m in range(n_mol) for c in range(K)]
                         + ["Tex"] + [f"vlsr{c + 1}" for c in range(K)] + ["dV"])
        return s

    def effective_prior(self, prior_stds, prior_means):
        """(mu, sigma, gauss flags) with the overrides of inference.py:200-201 / TMC1:243-247:
        sigma_vlsr := 0.8*mean_dV, sigma_dV := 0.3*mean_dV; flat in every column density."""
        mu = np.array(prior_means, dtype=float)
        sd = np.array(prior_stds, dtype=float)
        for i in self.idx_vlsr:
            sd[i] = mu[self.idx_dv] * 0.8
        sd[self.idx_dv] = mu[self.idx_dv] * 0.3
        gauss = np.ones(self.ndim, dtype=np.int32)
        gauss[self.ncol_indices] = 0
        return mu, sd, gauss

    def within_bounds(self, theta) -> bool:
        """Host mirror of is_within_bounds (inference.py:169-190; TMC1:224-233), used by the walker-ball
        initialisation loop (inference.py:444-450), not by the log-prob path (the device checks bounds)."""
        t = np.asarray(theta, dtype=float)
        if not (np.all(self.lo < t) and np.all(t < self.hi)):
            return False
        v = t[list(self.idx_vlsr)]
        if not math.isnan(self.vlsr_min_sep) and not np.all(v[:-1] < v[1:] - self.vlsr_min_sep):
            return False
        if not math.isnan(self.vlsr_max_sep) and not np.all(v[1:] < v[:-1] + self.vlsr_max_sep):
            return False
        return True


class EngineError(RuntimeError):
    pass


class LTEEngine:
    """One GPU, one stream, one fit resident in HBM."""

    STAT = {"launches": 0, "lines": 1, "active_channels": 2, "pairs": 3, "tiles": 4, "dv_list_e9": 5,
            "rebuilds": 6, "fused_ns": 7, "groups": 8, "records": 9, "hv_list_e9": 10, "build_us": 11,
            "graph_launches": 12, "collectives": 13, "collective_bytes": 14, "reruns": 15,
            "tight_tiles": 16, "tight_pairs": 17, "tight_builds": 18, "tight_hv_e9": 19,
            "sync_us": 20, "syncs": 21, "uncovered_events": 22, "sorted_batches": 23}

    def __init__(self, device: int = 0, precision="mixed"):
        self._lib = load_library()
        h = C.c_void_p()
        if self._lib.cha_create(int(device), C.byref(h)) != 0:
            raise EngineError(self._lib.cha_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.spec: Optional[ModelSpec] = None
        self.n_chan = 0
        self.set_precision(precision)

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self._lib.cha_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cha_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration ----------------------------------------------------------------------------
    def set_precision(self, precision):
        self.precision = _PREC[precision]
        self._ck(self._lib.cha_set_precision(self._h, self.precision))

    def set_model(self, spec: ModelSpec):
        K, M = spec.K, spec.M
        ss = np.ascontiguousarray(spec.idx_ss, dtype=np.int32)
        nc = np.ascontiguousarray(np.array(spec.idx_ncol, dtype=np.int32).reshape(M * K))
        vl = np.ascontiguousarray(spec.idx_vlsr, dtype=np.int32)
        self._ck(self._lib.cha_set_model(self._h, spec.ndim, K, M, _ptr(ss, _ip), _ptr(nc, _ip), spec.idx_tex,
                                         _ptr(vl, _ip), spec.idx_dv, float(spec.fixed_ss), float(spec.dish_size),
                                         float(spec.aligned_velocity), float(spec.mask_centre), float(spec.planck_eps)))
        self.spec = spec

    def set_molecule(self, mol_id: int, cat: MolCat, ll: Optional[float] = None, ul: Optional[float] = None,
                     line_idx=None):
        ll = self.spec.ll if ll is None else ll
        ul = self.spec.ul if ul is None else ul
        nu, li, el = _f64(cat.frequency), _f64(cat.logint), _f64(cat.elower)
        qp = _f64(list(cat.q_params) + [0.0] * (8 - len(cat.q_params)))
        sg, sE = _f64(cat.state_g), _f64(cat.state_E)
        if line_idx is None:
            lptr, nsel = None, 0
        else:
            lidx = np.ascontiguousarray(line_idx, dtype=np.int64)
            lptr, nsel = _ptr(lidx, _lp), lidx.size
        self._ck(self._lib.cha_set_molecule(self._h, mol_id, nu.size, _ptr(nu), _ptr(li), _ptr(el), cat.q_kind,
                                            _ptr(qp), len(cat.q_params), sg.size,
                                            _ptr(sg) if cat.q_kind == Q_SUM else None,
                                            _ptr(sE) if cat.q_kind == Q_SUM else None,
                                            float(ll), float(ul), lptr, nsel))

    def set_spectrum(self, freq, y, yerr):
        f, yy, ye = _f64(freq), _f64(y), _f64(yerr)
        if not (f.shape == yy.shape == ye.shape and f.ndim == 1):
            raise ValueError("freq, y, yerr must be 1-D arrays of one length")
        self._ck(self._lib.cha_set_spectrum(self._h, f.size, _ptr(f), _ptr(yy), _ptr(ye)))
        self.n_chan = f.size

    def set_prior(self, prior_stds, prior_means):
        """lnprior of inference.py:193-236 with the spec's bounds."""
        mu, sd, gauss = self.spec.effective_prior(prior_stds, prior_means)
        lo, hi = _f64(self.spec.lo), _f64(self.spec.hi)
        g = np.ascontiguousarray(gauss, dtype=np.int32)
        self._ck(self._lib.cha_set_prior(self._h, _ptr(lo), _ptr(hi), _ptr(_f64(mu)), _ptr(_f64(sd)), _ptr(g, _ip),
                                         float(self.spec.vlsr_min_sep), float(self.spec.vlsr_max_sep)))

    # -- evaluation (host buffers) ----------------------------------------------------------------
    def _theta(self, theta):
        if self.spec is None:
            raise EngineError("set_model has not been called")
        t = _f64(theta)
        if t.ndim == 1:
            t = t[None, :]
        if t.ndim != 2 or t.shape[1] != self.spec.ndim:
            raise ValueError(f"theta must be (nwalkers, {self.spec.ndim})")
        return np.ascontiguousarray(t)

    def _eval(self, fn, theta, width=1):
        t = self._theta(theta)
        out = np.empty((t.shape[0], width) if width > 1 else t.shape[0], dtype=np.float64)
        self._ck(fn(self._h, _ptr(t), t.shape[0], _ptr(out)))
        return out

    def log_prob(self, theta) -> np.ndarray:
        """Vectorised lnprob (inference.py:239-246): theta[nw, ndim] -> float64[nw]."""
        return self._eval(self._lib.cha_log_prob, theta)

    def log_like(self, theta) -> np.ndarray:
        return self._eval(self._lib.cha_log_like, theta)

    def log_prior(self, theta) -> np.ndarray:
        return self._eval(self._lib.cha_log_prior, theta)

    def simulate(self, theta) -> np.ndarray:
        """Model spectra make_model returns (inference.py:44-61): float64[nw, n_chan]."""
        out = self._eval(self._lib.cha_simulate, theta, width=max(self.n_chan, 1))
        return out.reshape(-1, max(self.n_chan, 1))[:, :self.n_chan]

    def stick_spectrum(self, mol_id, n_lines, Ncol, Tex, dV, source_size, dish_size):
        """MolSim(gauss=False) for one component (classes.py:336-397): (freq_sim, int_sim, tau_sim) inside (ll, ul]."""
        f = np.empty(n_lines); t = np.empty(n_lines); i = np.empty(n_lines)
        n = C.c_int64(0)
        self._ck(self._lib.cha_stick_spectrum(self._h, int(mol_id), float(Ncol), float(Tex), float(dV), float(source_size),
                                              float(dish_size), _ptr(f), _ptr(t), _ptr(i), C.byref(n)))
        k = int(n.value)
        return f[:k].copy(), i[:k].copy(), t[:k].copy()

    def make_model(self, freqs, taus, x, vlsr, dV, Tex, source_size, aligned_velocity, dish_size,
                   mask_centre=0.0, planck_eps=1e-10):
        """make_model_numba (inference.py:44-61) on explicit line lists."""
        fr, ta, xx = _f64(freqs), _f64(taus), _f64(x)
        out = np.empty(xx.size)
        self._ck(self._lib.cha_make_model(self._h, fr.size, _ptr(fr), _ptr(ta), xx.size, _ptr(xx), float(vlsr), float(dV),
                                          float(Tex), float(source_size), float(aligned_velocity), float(dish_size),
                                          float(mask_centre), float(planck_eps), _ptr(out)))
        return out

    def count_window_pairs(self, theta) -> np.ndarray:
        t = self._theta(theta)
        out = np.zeros(t.shape[0], dtype=np.int64)
        self._ck(self._lib.cha_count_window_pairs(self._h, _ptr(t), t.shape[0], _ptr(out, _lp)))
        return out

    # -- evaluation (device-resident torch tensors) -----------------------------------------------
    def log_prob_device(self, theta, out=None, with_prior=True, sync=True, wait_torch=True):
        """theta: CUDA float64 tensor [nw, ndim] on this engine's device; returns a CUDA tensor [nw].
        Runs on the engine's own stream; ``sync`` waits for it (required before reading ``out``
        from another stream).  ``wait_torch=False`` skips the wait on torch's current stream (the caller
        guarantees theta is already complete, e.g. it waited once for a whole batch of engines)."""
        import torch
        if not (theta.is_cuda and theta.dtype == torch.float64 and theta.is_contiguous()):
            raise ValueError("theta must be a contiguous CUDA float64 tensor")
        nw = theta.shape[0]
        if out is None:
            out = torch.empty(nw, dtype=torch.float64, device=theta.device)
        if wait_torch:
            torch.cuda.current_stream(theta.device).synchronize()
        self._ck(self._lib.cha_log_prob_dev(self._h, C.c_void_p(theta.data_ptr()), nw, C.c_void_p(out.data_ptr()),
                                            1 if with_prior else 0))
        if sync:
            self.sync()
        return out

    def simulate_device(self, theta, out=None, sync=True):
        """Channel-stream kernel on device tensors: theta CUDA float64 [nw, ndim] -> model spectra [nw, n_chan]
        (what make_model returns, inference.py:44-61), written straight to HBM in the caller's channel order."""
        import torch
        if not (theta.is_cuda and theta.dtype == torch.float64 and theta.is_contiguous()):
            raise ValueError("theta must be a contiguous CUDA float64 tensor")
        nw = theta.shape[0]
        if out is None:
            out = torch.empty((nw, self.n_chan), dtype=torch.float64, device=theta.device)
        torch.cuda.current_stream(theta.device).synchronize()
        self._ck(self._lib.cha_simulate_dev(self._h, C.c_void_p(theta.data_ptr()), nw, C.c_void_p(out.data_ptr())))
        if sync:
            self.sync()
        return out

    def sync(self):
        self._ck(self._lib.cha_sync(self._h))

    def stat(self, name: str) -> int:
        return int(self._lib.cha_stat(self._h, self.STAT[name]))

    def stats(self) -> dict:
        d = {k: self.stat(k) for k in self.STAT}
        d["dv_list"] = d.pop("dv_list_e9") * 1e-9
        d["hv_list"] = d.pop("hv_list_e9") * 1e-9
        d["tight_hv"] = d.pop("tight_hv_e9") * 1e-9
        return d

    # -- on-device sampler ------------------------------------------------------------------------
    def sampler_init(self, coords_local, nw_global=None, w0=0, seed=0, a=2.0):
        c = self._theta(coords_local)
        nwg = c.shape[0] if nw_global is None else int(nw_global)
        self._ck(self._lib.cha_sampler_init(self._h, nwg, int(w0), c.shape[0], _ptr(c), C.c_uint64(seed), float(a)))
        self._s_local = c.shape[0]

    def sampler_half_step(self, step: int, split: int, all_coords_ptr: int):
        self._ck(self._lib.cha_sampler_half_step(self._h, int(step), int(split), C.c_void_p(all_coords_ptr)))

    def sampler_device_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self._lib.cha_sampler_coords_dev(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- walkers sharded over ranks (one process per GPU) --------------------------------------------
    COMM_ID_BYTES = 128

    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0: the 128-byte id every rank hands to ``comm_init`` (the caller broadcasts it)."""
        lib = load_library()
        buf = C.create_string_buffer(LTEEngine.COMM_ID_BYTES)
        if lib.cha_comm_unique_id(buf) != 0:
            raise EngineError(lib.cha_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank: int, world: int, comm_id: bytes):
        """Collective over all ranks: attaches the communicator the sampler's per-half-step all-gather runs on."""
        if len(comm_id) != self.COMM_ID_BYTES:
            raise ValueError("comm_id must be the 128 bytes comm_unique_id() returned on rank 0")
        self._ck(self._lib.cha_comm_init(self._h, int(rank), int(world), C.create_string_buffer(comm_id, self.COMM_ID_BYTES)))

    def comm_destroy(self):
        self._ck(self._lib.cha_comm_destroy(self._h))

    def sampler_run(self, step0: int, n_steps: int, store_every: int = 0):
        """Queue n_steps stretch-move steps on the engine's stream (all-gather inside when sharded); no host sync."""
        self._ck(self._lib.cha_sampler_run(self._h, int(step0), int(n_steps), int(store_every)))

    def sampler_chain_len(self) -> int:
        return int(self._lib.cha_sampler_chain_len(self._h))

    def sampler_chain_read(self, slot0=0, n_slots=None):
        """(coords[n_slots, n_local, ndim], logp[n_slots, n_local]) of the chain kept in HBM."""
        n = self.sampler_chain_len() - slot0 if n_slots is None else int(n_slots)
        coords = np.empty((n, self._s_local, self.spec.ndim)); logp = np.empty((n, self._s_local))
        if n:
            self._ck(self._lib.cha_sampler_chain_read(self._h, int(slot0), n, _ptr(coords), _ptr(logp)))
        return coords, logp

    def sampler_chain_clear(self):
        self._ck(self._lib.cha_sampler_chain_clear(self._h))

    def sampler_get(self):
        coords = np.empty((self._s_local, self.spec.ndim))
        logp = np.empty(self._s_local)
        nacc = C.c_int64(0)
        self._ck(self._lib.cha_sampler_get(self._h, _ptr(coords), _ptr(logp), C.byref(nacc)))
        return coords, logp, int(nacc.value)
