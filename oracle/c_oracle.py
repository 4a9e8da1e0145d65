"""ctypes wrapper of oracle/_ref/liblte_oracle.so (C restatement of the reference path).
Test infrastructure / CPU baseline only -- never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import lte_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "liblte_oracle.so")
_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_int64)


class _Mol(C.Structure):
    _fields_ = [("n", C.c_int64), ("nu", _dp), ("logint", _dp), ("elower", _dp), ("q_kind", C.c_int),
                ("n_qp", C.c_int), ("qp", C.c_double * 8), ("n_states", C.c_int64), ("sg", _dp), ("sE", _dp),
                ("i0", C.c_int64), ("i1", C.c_int64), ("n_sel", C.c_int64), ("sel", _lp), ("aij_gup", _dp)]


class _Spec(C.Structure):
    _fields_ = [("ndim", C.c_int), ("K", C.c_int), ("M", C.c_int), ("idx_ss", C.c_int * 8),
                ("idx_ncol", C.c_int * 32), ("idx_tex", C.c_int), ("idx_vlsr", C.c_int * 8), ("idx_dv", C.c_int),
                ("fixed_ss", C.c_double), ("dish", C.c_double), ("al", C.c_double), ("mc", C.c_double),
                ("eps", C.c_double), ("guard", C.c_int), ("has_prior", C.c_int),
                ("lo", C.c_double * 64), ("hi", C.c_double * 64), ("mu", C.c_double * 64), ("sg", C.c_double * 64),
                ("gauss", C.c_int * 64), ("vmin_sep", C.c_double), ("vmax_sep", C.c_double)]


def build(force=False):
    src = os.path.join(HERE, "lte_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_eval.restype = C.c_int
        _lib.oracle_eval.argtypes = [C.POINTER(_Spec), C.POINTER(_Mol), C.c_int64, _dp, _dp, _dp, _dp, C.c_int64,
                                     C.c_int, C.c_int, _dp]
        _lib.oracle_mol_prepare.restype = C.c_int
        _lib.oracle_mol_prepare.argtypes = [C.POINTER(_Mol), C.c_double, C.c_double]
        _lib.oracle_mol_free.argtypes = [C.POINTER(_Mol)]
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _q_encoding(cat: O.OracleCatalog):
    kind, p = O.q_mode_for(cat.catalog_file)
    if kind == "poly":
        return 0, list(p)
    if kind == "lin":
        a, b, s = p
        return 1, [a, b, 3.0 if s == "3" else 1.0, 3.0 if s == "/3" else 0.0]
    if kind == "pow":
        a, pw, b = p
        return 2, [a, pw, b, 1.0 if b != 0.0 else 0.0]
    return 3, []


class COracle:
    """Same problem description as the NumPy oracle: (spec, cats, datagrid) + optional prior."""

    def __init__(self, spec: O.ModelSpec, cats, datagrid, prior=None):
        L = lib()
        self._keep = []
        x, y, yerr, lidx = datagrid
        self.x = np.ascontiguousarray(x, float); self.y = np.ascontiguousarray(y, float)
        self.yerr = np.ascontiguousarray(yerr, float)
        self.spec = spec
        s = _Spec()
        s.ndim, s.K, s.M = spec.ndim, spec.K, len(cats)
        for c in range(spec.K):
            s.idx_ss[c] = spec.idx_ss[c]; s.idx_vlsr[c] = spec.idx_vlsr[c]
            for m in range(len(cats)):
                s.idx_ncol[m * spec.K + c] = spec.idx_ncol[m][c]
        s.idx_tex, s.idx_dv = spec.idx_tex, spec.idx_dv
        s.fixed_ss = spec.fixed_ss; s.dish = spec.dish_size; s.al = spec.aligned_velocity
        s.mc = spec.mask_centre; s.eps = spec.planck_eps; s.guard = 1 if spec.guard_nonfinite else 0
        s.vmin_sep, s.vmax_sep = spec.vlsr_min_sep, spec.vlsr_max_sep
        if prior is not None:
            stds, means = prior
            mu = np.asarray(means, float); sd = np.array(stds, float)
            for i in spec.idx_vlsr:
                sd[i] = mu[spec.idx_dv] * 0.8
            sd[spec.idx_dv] = mu[spec.idx_dv] * 0.3
            ncol = {i for row in spec.idx_ncol for i in row}
            for p in range(spec.ndim):
                s.lo[p] = spec.lo[p]; s.hi[p] = spec.hi[p]; s.mu[p] = mu[p]; s.sg[p] = sd[p]
                s.gauss[p] = 0 if p in ncol else 1
            s.has_prior = 1
        self._spec = s
        Mols = _Mol * len(cats)
        self._mols = Mols()
        for m, cat in enumerate(cats):
            mm = self._mols[m]
            nu = np.ascontiguousarray(cat.frequency, float); li = np.ascontiguousarray(cat.logint, float)
            el = np.ascontiguousarray(cat.elower, float)
            sel = np.ascontiguousarray(lidx[m], np.int64)
            kind, qp = _q_encoding(cat)
            mm.n = nu.size; mm.nu = nu.ctypes.data_as(_dp); mm.logint = li.ctypes.data_as(_dp)
            mm.elower = el.ctypes.data_as(_dp); mm.q_kind = kind; mm.n_qp = len(qp)
            for i, v in enumerate(qp):
                mm.qp[i] = v
            self._keep += [nu, li, el, sel]
            if kind == 3:
                J, E = O.unique_states(cat)
                g = np.ascontiguousarray(2 * J + 1); E = np.ascontiguousarray(E)
                mm.n_states = g.size; mm.sg = g.ctypes.data_as(_dp); mm.sE = E.ctypes.data_as(_dp)
                self._keep += [g, E]
            if L.oracle_mol_prepare(C.byref(mm), float(spec.ll), float(spec.ul)):
                raise MemoryError
            ntrim = mm.i1 - mm.i0
            sel = np.where(sel < 0, sel + ntrim, sel)
            if sel.size and (sel.min() < 0 or sel.max() >= ntrim):
                raise IndexError("line index outside the trimmed catalog")
            sel = np.ascontiguousarray(sel, np.int64); self._keep.append(sel)
            mm.n_sel = sel.size; mm.sel = sel.ctypes.data_as(_lp)

    def _run(self, theta, mode, nthreads=0):
        t = np.ascontiguousarray(np.atleast_2d(theta), float)
        nw = t.shape[0]
        out = np.empty(nw * (self.x.size if mode == 3 else 1))
        rc = lib().oracle_eval(C.byref(self._spec), self._mols, self.x.size, self.x.ctypes.data_as(_dp),
                               self.y.ctypes.data_as(_dp), self.yerr.ctypes.data_as(_dp), t.ctypes.data_as(_dp), nw,
                               mode, int(nthreads), out.ctypes.data_as(_dp))
        if rc:
            raise MemoryError
        return out.reshape(nw, -1) if mode == 3 else out

    def lnlike(self, theta, nthreads=0): return self._run(theta, 0, nthreads)
    def lnprob(self, theta, nthreads=0): return self._run(theta, 1, nthreads)
    def lnprior(self, theta, nthreads=0): return self._run(theta, 2, nthreads)
    def simulate(self, theta, nthreads=0): return self._run(theta, 3, nthreads)

    @staticmethod
    def max_threads():
        return int(lib().oracle_max_threads())

    def __del__(self):
        try:
            for m in range(len(self._mols)):
                lib().oracle_mol_free(C.byref(self._mols[m]))
        except Exception:
            pass
