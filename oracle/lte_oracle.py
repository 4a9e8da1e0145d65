"""CPU ORACLE (test infrastructure, NOT product code).

A plain NumPy restatement of the reference's walker log-probability hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``cha1_mcmc_b200`` never does.

Every function cites the reference file:line (relative to the reference
repository root) whose arithmetic it restates.  The restatement is pinned
against the *unmodified* reference executed in the build container (see
``oracle/ref_shim.py`` + ``oracle/make_golden.py`` -> ``tests/golden/*.npz``);
the reference itself ships no tests or golden vectors, so
"parity pinned by executing the reference", not by reference-owned KATs.

emcee 3.1.6 (requirements.txt:9) is a third-party dependency that is absent
from the reference tree and from this image: ``stretch_move_step`` restates
its published StretchMove/RedBlueMove algorithm; parity at that boundary is
UNPINNED (no golden output of emcee exists anywhere in the reference).
"""
from __future__ import annotations

import gzip
import math
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

# --------------------------------------------------------------------------
# constants -- spectral_simulator/constants.py:2-7 (the *expressions* matter:
# 2.998 * 10**10 is 29980000000.000004, not 2.998e10)
# --------------------------------------------------------------------------
kcm = 0.69503476
ckm = 2.998 * 10**5
ccm = 2.998 * 10**10
cm = 2.998 * 10**8
h = 6.626 * 10 ** (-34)
k = 1.381 * 10 ** (-23)

TBG = 2.7                    # classes.py:492, inference.py:57
CT = 300                     # classes.py:19
FWHM_TO_SIGMA = 2.355        # inference.py:53
BEAM_CONST = 206265 * 1.22   # inference.py:38


# --------------------------------------------------------------------------
# catalog parsing -- classes.py:132-288 (+ functions.py:330-335, 340-501)
# --------------------------------------------------------------------------
def _open_text(path):
    if path.endswith(".gz"):
        return gzip.open(path, "rt")
    return open(path)


def _letter_qn(text):
    """functions.py:340-501: a letter anywhere in the field selects the
    hundreds/tens offset (A/a->100, B/b->110 ... Z/z->350); the units digit is
    the *second character* of the field as passed.  No letter -> 0."""
    val = 0
    for ch in "ABCDEFGHIJKLMNOPQRSTUVWXYZ":          # sequence of ifs, last match wins
        if ch in text:
            val = 100 + 10 * (ord(ch) - ord("A")) + int(text[1])
    for ch in "abcdefghijklmnopqrstuvwxyz":
        if ch in text:
            val = 100 + 10 * (ord(ch) - ord("a")) + int(text[1])
    return int(val)


@dataclass
class OracleCatalog:
    catalog_file: str
    frequency: np.ndarray
    error: np.ndarray
    logint: np.ndarray
    dof: np.ndarray
    elower: np.ndarray
    gup: np.ndarray
    tag: np.ndarray
    qnformat: np.ndarray
    qn: np.ndarray              # (N, 12) int64, columns qn1..qn12
    qns: int = 0
    eupper: np.ndarray = field(default=None)
    sijmu: np.ndarray = field(default=None)
    aij: np.ndarray = field(default=None)


def parse_catalog(path: str, name_for_q: Optional[str] = None) -> OracleCatalog:
    """classes.py:56-112 minus the O(N^2) ``glow`` match (100-110): glow
    multiplies both numerator (349, 351) and denominator (353) of tau and
    cancels, so it is never needed on this path.

    ``name_for_q`` overrides the string the partition-function dispatch
    (functions.py:139-257) sees; the reference uses the full path."""
    with _open_text(path) as fh:
        rows = [ln for ln in fh]
    n = len(rows)
    freq = np.empty(n); err = np.empty(n); logint = np.empty(n)
    dof = np.empty(n, dtype=np.int64); elo = np.empty(n)
    gup = np.empty(n, dtype=np.int64); tag = np.empty(n, dtype=np.int64)
    fmt = np.empty(n, dtype=np.int64)
    qtxt = np.empty((n, 12), dtype=object)
    for r, x in enumerate(rows):                       # classes.py:154-178
        freq[r] = float(x[:13].strip())
        err[r] = float(x[13:21].strip())
        logint[r] = float(x[21:29].strip())
        dof[r] = int(x[29:31].strip())
        elo[r] = float(x[31:41].strip())
        g = x[41:44]
        try:
            gup[r] = int(g.strip())
        except ValueError:                             # classes.py:160-163
            gup[r] = _letter_qn(g)
        tag[r] = int(x[44:51].strip())
        fmt[r] = int(x[51:55].strip())
        for q in range(11):
            qtxt[r, q] = x[55 + 2 * q:57 + 2 * q].strip()
        qtxt[r, 11] = x[77:].strip()
    qn = np.zeros((n, 12), dtype=np.int64)
    for q in range(12):
        col = qtxt[:, q]
        if np.any(col == "+") or np.any(col == "-"):   # classes.py:180-214 -> fix_pm
            col = col.copy()
            col[col == ""] = "0"
            col[col == "+"] = "1"
            col[col == "-"] = "2"
        for r in range(n):                             # classes.py:216-286
            try:
                qn[r, q] = int(col[r])
            except ValueError:
                qn[r, q] = _letter_qn(col[r])
    cat = OracleCatalog(catalog_file=name_for_q if name_for_q is not None else path,
                        frequency=freq, error=err, logint=logint, dof=dof, elower=elo,
                        gup=gup, tag=tag, qnformat=fmt, qn=qn)
    cat.qns = min(int(str(int(fmt[0]))[-1:]), 6)       # classes.py:116-122
    cat.eupper = elo + freq / 29979.2458               # classes.py:90
    q_ct = calc_q(cat, CT)                             # classes.py:94
    cat.sijmu = ((np.exp(-(elo / 0.695) / CT) - np.exp(-(cat.eupper / 0.695) / CT)) ** (-1)
                 * ((10 ** logint) / freq) * ((4.16231 * 10 ** (-5)) ** (-1)) * q_ct)   # classes.py:95
    with np.errstate(divide="ignore", invalid="ignore"):
        cat.aij = 1.16395 * 10 ** (-20) * freq ** 3 * cat.sijmu / gup                    # classes.py:98
    return cat


# --------------------------------------------------------------------------
# partition function -- functions.py:136-325
# --------------------------------------------------------------------------
# (substring tests, in order; "hfs" handling as in the reference). Each entry:
# (needle, hfs_requirement, kind, params) with kind "lin": a*T+b scaled by s,
# "pow": a*T**p + b, "poly": sum c_n T^n.
_Q_DISPATCH = [
    ("n2h+_hfs.cat", None, "poly", (3.32018827e+00, 4.01951955e+00, 3.28722820e-05, -3.13420474e-08)),
    ("acetone.cat", None, "poly", (16431, -2728.3, 245.28, -5.5477, 0.05471337, -0.00021050085, 2.91296 * 10 ** (-7))),
    ("sh.cat", None, "poly", (15.357239728157400, 0.069272946237033, 0.002288160909445, -0.000008528126823, 0.000000012549467)),
    ("h2s.cat", None, "poly", (-1.764494755639740, 0.507648423477309, 0.005498622332982, -0.000004859941547)),
    ("hcn.cat", None, "poly", (.386550361, 1.48629408, -1.15188755 * 10 ** -3, 4.62476813 * 10 ** -6, -1.64946939 * 10 ** -9)),
]


def q_mode_for(catalog_file: str):
    """Resolve which branch of functions.py:139-257 a catalog *path* takes.
    Returns (kind, params): kind in {"poly","lin","pow","sum"}.
      lin : Q = s*(a*T + b)       params (a, b, s)   [s = 3, 1 or 1/3 written as the reference writes it]
      pow : Q = a*T**p + b        params (a, p, b)
      poly: Q = sum c[n]*T**n
      sum : explicit state sum (functions.py:263-323)
    Only the branches reachable from the shipped catalogs + the generic ones
    are listed; the methanol/c2n/ch2nh/13ch3oh families (155-168) are included
    for completeness of the dispatch order."""
    f = catalog_file.lower()
    for needle, _, kind, params in _Q_DISPATCH:
        if needle in f:
            return kind, params
    if any(s in f for s in ("methanol.cat", "ch3oh.cat", "ch3oh_v0.cat", "ch3oh_v1.cat", "ch3oh_v2.cat", "ch3oh_vt.cat")):
        return "poly", (-1.25670, 4.39632 * 10 ** -1, 2.05911 * 10 ** -1, -1.83807 * 10 ** -3,
                        1.27624 * 10 ** -5, -4.04024 * 10 ** -8, 4.83410 * 10 ** -11)
    if "13methanol.cat" in f or "13ch3oh.cat" in f:
        return "poly", (-31.876881967, 4.317920731, 0.076540934, 0.000050130)
    if "c2n.cat" in f or "ccn.cat" in f:
        return "poly", (22.55770, 7.135161, 0.1837397, -1.40473 * 10 ** (-3), 5.99936 * 10 ** (-6),
                        -1.324086 * 10 ** (-8), 1.173755 * 10 ** (-11))
    if "ch2nh.cat" in f:
        return "pow", (1.2152, 1.4863, 0.0)
    if "13ch3oh.cat" in f or "c033502.cat" in f:
        return "pow", (0.399272, 1.756329, 0.0)
    hfs = "hfs" in f
    lin = [("hc3n", 4.581898, 0.2833, "x3"), ("hc2nc_hfs", 12.58340, 1.0604, "plain"),
           ("hc5n", 15.65419, 0.2214, "x3"), ("hc4nc", 44.62171, 0.6734, "div3"),
           ("hc7n", 36.94999, 0.1356, "x3"), ("hc6nc", 107.3126, 1.2714, "div3"),
           ("hc9n", 71.7308577, 0.02203968, "x3")]
    for needle, a, b, how in lin:
        if needle in f:
            if how == "plain":
                return "lin", (a, b, "1")
            if how == "x3":                       # functions.py:173-176 etc.
                return "lin", (a, b, "3" if hfs else "1")
            return "lin", (a, b, "1" if hfs else "/3")   # functions.py:187-190, 197-200
    if "hc11n.cat" in f and not hfs:              # functions.py:207
        return "lin", (123.2554, 0.1381, "1")
    if "hc11n" in f and hfs:                      # functions.py:209
        return "lin", (123.2554, 0.1381, "3")
    pw = [("propargylcyanide", 41.542, 1.5008, 0.0), ("pyrrole", 27.727, 1.4752, 0.0),
          ("cyclopropylcyanide_hfs", 38.199, 1.4975, 0.0), ("pyridine", 50.478, 1.4955, 0.0),
          ("1-cyanonaphthalene", 560.39, 1.4984, 0.0), ("2-cyanonaphthalene", 562.57, 1.4993, 0.0),
          ("furan", 33.725, 1.4982, 0.0), ("phenol", 264.20, 1.4984, 0.0),
          ("benzaldehyde", 53.798, 1.4997, 0.0), ("anisole", 54.850, 1.4992, 0.0),
          ("azulene", 96.066, 1.4988, 0.0), ("acenaphthene", 161.29, 1.4994, 0.0),
          ("acenapthylene", 151.58, 1.4988, 0.0), ("fluorene", 219.51, 1.4996, 0.0),
          ("benzonitrile", 25.896, 1.4998, 0.38109)]
    for needle, a, p, b in pw:
        if needle in f:
            return "pow", (a, p, b)
    return "sum", ()


def unique_states(cat: OracleCatalog):
    """functions.py:264-317: rows (qn7..qn(6+qns), elower) as floats,
    de-duplicated; returns (J = first lower QN, E) per unique state, in sorted
    order (the reference iterates a Python set: order is unspecified)."""
    cols = [cat.qn[:, 6 + q].astype(float) for q in range(cat.qns)] + [cat.elower]
    arr = np.stack(cols, axis=1)
    uniq = sorted(set(map(tuple, arr)))
    u = np.array(uniq, dtype=float).reshape(len(uniq), cat.qns + 1)
    return u[:, 0].copy(), u[:, cat.qns].copy()


def calc_q(cat: OracleCatalog, T: float) -> float:
    """functions.py:136-325."""
    kind, p = q_mode_for(cat.catalog_file)
    if kind == "poly":
        return float(sum(c * T ** n for n, c in enumerate(p)))
    if kind == "lin":
        a, b, s = p
        base = a * T + b
        return 3 * base if s == "3" else (base / 3 if s == "/3" else base)
    if kind == "pow":
        a, pw, b = p
        return a * T ** pw + b if b != 0.0 else a * T ** pw
    J, E = unique_states(cat)
    return float(np.sum((2 * J + 1) * np.exp(-E / (kcm * T))))       # functions.py:319-323


# --------------------------------------------------------------------------
# window trim -- functions.py:507-540 (single [ll],[ul] chunk as used by the path)
# --------------------------------------------------------------------------
def trim_bounds(frequency: np.ndarray, ll: float, ul: float):
    above = np.nonzero(frequency > ll)[0]
    if above.size:
        i0 = int(above[0])
    elif frequency[-1] < ll:
        return 0, 0                                   # functions.py:523-524: chunk skipped
    else:
        i0 = 0                                        # functions.py:526
    above = np.nonzero(frequency > ul)[0]
    i1 = int(above[0]) if above.size else len(frequency)   # functions.py:528-531
    return i0, i1


# --------------------------------------------------------------------------
# line optical depths -- classes.py:336-397 (gauss=False branch), one component
# --------------------------------------------------------------------------
def line_taus(cat: OracleCatalog, Ncol: float, Tex: float, dV: float, ll: float, ul: float):
    """Returns (freq_sim, tau_sim) over catalog lines in (ll, ul].  glow is
    dropped (cancels between classes.py:349/351 and 353)."""
    with np.errstate(all="ignore"):
        Q = calc_q(cat, Tex)                                                             # classes.py:347
        Nl_over_glow = Ncol * np.exp(-cat.elower / (0.695 * Tex)) / Q                   # classes.py:349
        num = ((ccm / (cat.frequency * 10 ** 6)) ** 2 * cat.aij * cat.gup * Nl_over_glow
               * (1 - np.exp(-(h * cat.frequency * 10 ** 6) / (k * Tex))))              # classes.py:351
        den = 8 * np.pi * (dV * cat.frequency * 10 ** 6 / ckm)                          # classes.py:353
        tau = num / den
    i0, i1 = trim_bounds(cat.frequency, ll, ul)                                          # classes.py:356-364
    return cat.frequency[i0:i1].copy(), tau[i0:i1].copy()


def stick_intensity(freq, tau, Tex, source_size, dish_size):
    """classes.py:369-377: the stick-spectrum brightness ``int_sim`` (used only
    by the data reduction, inference.py:324-327; unused by lnlike)."""
    with np.errstate(all="ignore"):
        J_T = (h * freq * 10 ** 6 / k) * (np.exp((h * freq * 10 ** 6) / (k * Tex)) - 1) ** -1
        J_bg = (h * freq * 10 ** 6 / k) * (np.exp((h * freq * 10 ** 6) / (k * TBG)) - 1) ** -1
        raw = (J_T - J_bg) * (1 - np.exp(-tau))
    return raw * beam_dilution(freq, source_size, dish_size)


def beam_dilution(freq, source_size, dish_size):
    """inference.py:33-41 / functions.py:627-650."""
    beam = cm / (freq * 1e6) * 206265 * 1.22 / dish_size
    return source_size ** 2 / (beam ** 2 + source_size ** 2)


# --------------------------------------------------------------------------
# model description shared by all layouts
# --------------------------------------------------------------------------
@dataclass
class ModelSpec:
    """One fit: K velocity components of M molecules on one channel grid.

    theta layout is given by index arrays (or a fixed value where a parameter
    is not free):
      idx_ss[c]      -> theta index of source size of component c (or -1: fixed_ss)
      idx_ncol[m][c] -> theta index of the column density of molecule m, comp c
      idx_tex, idx_dv, idx_vlsr[c]
    mask_centre: 0 for inference.py (the +aligned-aligned cancels, 51-52),
                 5.8 for TMC1_four_component.py:160
    planck_eps : 1e-10 (inference.py:56-57) or 0 (TMC1_four_component.py:168-169)
    """
    ndim: int
    K: int
    idx_ss: Sequence[int]
    idx_ncol: Sequence[Sequence[int]]
    idx_tex: int
    idx_vlsr: Sequence[int]
    idx_dv: int
    fixed_ss: float = float("nan")
    dish_size: float = 100.0
    aligned_velocity: float = 0.0
    mask_centre: float = 0.0
    planck_eps: float = 1e-10
    ll: float = 7000.0
    ul: float = 30000.0
    # prior / bounds
    lo: Optional[np.ndarray] = None       # strict lower bounds per parameter (-inf = none)
    hi: Optional[np.ndarray] = None
    vlsr_min_sep: float = float("nan")    # TMC1:229  vlsr_c < vlsr_{c+1} - sep
    vlsr_max_sep: float = float("nan")    # TMC1:230  vlsr_{c+1} < vlsr_c + sep
    guard_nonfinite: bool = True          # inference.py:162-164 (-inf); TMC1 returns NaN as is


def spec_inference(fixed_source_size, bounds, dish_size, aligned_velocity, ll, ul) -> ModelSpec:
    """theta layouts of inference.py:133-137 with the box bounds of 169-190."""
    if fixed_source_size is not None:
        s = ModelSpec(ndim=4, K=1, idx_ss=[-1], idx_ncol=[[0]], idx_tex=1, idx_vlsr=[2], idx_dv=3,
                      fixed_ss=float(fixed_source_size))
        names = ["Ncol", "Tex", "vlsr", "dV"]
    else:
        s = ModelSpec(ndim=5, K=1, idx_ss=[0], idx_ncol=[[1]], idx_tex=2, idx_vlsr=[3], idx_dv=4)
        names = ["source_size", "Ncol", "Tex", "vlsr", "dV"]
    s.dish_size = dish_size; s.aligned_velocity = aligned_velocity
    s.mask_centre = 0.0; s.planck_eps = 1e-10; s.ll = ll; s.ul = ul
    s.lo = np.array([bounds[n][0] for n in names], dtype=float)
    s.hi = np.array([bounds[n][1] for n in names], dtype=float)
    return s


def spec_tmc1(K: int = 4, n_mol: int = 1) -> ModelSpec:
    """theta layout of TMC1_four_component.py:189 generalised to K components
    and (spec-by-composition) M molecules sharing
    ss/Tex/vlsr/dV: [ss_1..K, Ncol(m=0)_1..K, ..., Ncol(m=M-1)_1..K, Tex, vlsr_1..K, dV]."""
    ndim = K + n_mol * K + 1 + K + 1
    s = ModelSpec(ndim=ndim, K=K, idx_ss=list(range(K)),
                  idx_ncol=[[K + m * K + c for c in range(K)] for m in range(n_mol)],
                  idx_tex=K + n_mol * K, idx_vlsr=[K + n_mol * K + 1 + c for c in range(K)],
                  idx_dv=K + n_mol * K + 1 + K)
    s.dish_size = 100; s.aligned_velocity = 0.0; s.mask_centre = 5.8; s.planck_eps = 0.0
    s.ll = 7000; s.ul = 30000
    lo = np.full(ndim, -np.inf); hi = np.full(ndim, np.inf)
    lo[:K] = 0.0; hi[:K] = 200.0                                  # TMC1:227
    lo[K:K + n_mol * K] = 0.0; hi[K:K + n_mol * K] = 10 ** 16.    # TMC1:228
    lo[s.idx_tex] = 2.7                                            # TMC1:231
    hi[s.idx_dv] = 0.3                                             # TMC1:231
    s.lo, s.hi = lo, hi
    s.vlsr_min_sep, s.vlsr_max_sep = 0.05, 0.3                     # TMC1:229-230
    s.guard_nonfinite = False
    return s


# --------------------------------------------------------------------------
# model spectrum -- inference.py:44-61 / TMC1_four_component.py:148-181
# --------------------------------------------------------------------------
def _planck(x, T, eps):
    """inference.py:56-57 (eps=1e-10) / TMC1_four_component.py:168-169 (eps=0)."""
    with np.errstate(all="ignore"):
        return (h * x * 1e6 / k) / (np.exp((h * x * 1e6) / (k * T)) - 1 + eps)


def component_taus(spec: ModelSpec, cats, theta):
    """Per component c: (freqs[L], taus[L]) summed over nothing -- one list per
    molecule, concatenated in molecule order (lnlike: inference.py:141-144 /
    TMC1:193-211).  Line selection by ``line_indices`` is applied by the caller."""
    out = []
    Tex = theta[spec.idx_tex]; dV = theta[spec.idx_dv]
    for c in range(spec.K):
        fr, ta = [], []
        for m, cat in enumerate(cats):
            f, t = line_taus(cat, theta[spec.idx_ncol[m][c]], Tex, dV, spec.ll, spec.ul)
            fr.append(f); ta.append(t)
        out.append((fr, ta))
    return out


def simulate(spec: ModelSpec, cats, line_indices, grid_freq, theta, windowed: bool = False):
    """Model spectrum m_j for one theta.  ``line_indices`` is one index array
    per molecule (into that molecule's trimmed line list).

    windowed=False follows the reference loop literally: for every selected
    line the velocity of *every* channel is formed and masked
    (inference.py:50-53).  windowed=True evaluates only the channels a
    searchsorted bracket says can pass the mask, then applies the *same* mask:
    identical output, O(sum W_i) instead of O(L*C); used for large grids."""
    theta = np.asarray(theta, dtype=float)
    x = np.asarray(grid_freq, dtype=float)
    Tex = theta[spec.idx_tex]; dV = theta[spec.idx_dv]
    al = spec.aligned_velocity; mc = spec.mask_centre
    comps = component_taus(spec, cats, theta)
    dJ = _planck(x, Tex, spec.planck_eps) - _planck(x, TBG, spec.planck_eps)
    total = np.zeros(x.shape)
    order_sorted = bool(np.all(np.diff(x) >= 0)) if windowed else False
    with np.errstate(all="ignore"):
        for c in range(spec.K):
            vl = theta[spec.idx_vlsr[c]]
            ss = spec.fixed_ss if spec.idx_ss[c] < 0 else theta[spec.idx_ss[c]]
            acc = np.zeros(x.shape)
            fr_m, ta_m = comps[c]
            for m in range(len(cats)):
                fsel = fr_m[m][line_indices[m]]
                tsel = ta_m[m][line_indices[m]]
                for fi, ti in zip(fsel, tsel):
                    if windowed and order_sorted and dV > 0:
                        half = (abs(mc) + 10 * dV) / ckm * fi * 1.01 + 1e-9
                        a = np.searchsorted(x, fi - half, "left"); b = np.searchsorted(x, fi + half, "right")
                        xs = x[a:b]
                        vg = (fi - xs) / fi * ckm + al
                        msk = np.abs(vg - al - mc) < dV * 10
                        acc[a:b][msk] += ti * np.exp(-0.5 * ((vg[msk] - vl) / (dV / 2.355)) ** 2)
                    else:
                        vg = (fi - x) / fi * ckm + al                                   # inference.py:51
                        msk = np.abs(vg - al - mc) < dV * 10                            # inference.py:52 / TMC1:160
                        acc[msk] += ti * np.exp(-0.5 * ((vg[msk] - vl) / (dV / 2.355)) ** 2)   # inference.py:53
            total += dJ * (1 - np.exp(-acc)) * beam_dilution(x, ss, spec.dish_size)     # inference.py:60 / TMC1:173-179
    return total


def lnlike(spec: ModelSpec, cats, datagrid, theta, windowed: bool = False):
    """inference.py:127-166 / TMC1_four_component.py:185-220.
    datagrid = (freqs, ints, yerrs, [line_indices per molecule])."""
    x, y, yerr, lidx = datagrid
    model = simulate(spec, cats, lidx, x, theta, windowed=windowed)
    with np.errstate(all="ignore"):
        inv_sigma2 = 1.0 / (np.asarray(yerr, dtype=float) ** 2)                         # inference.py:157
        tot = np.sum((y - model) ** 2 * inv_sigma2 - np.log(inv_sigma2))                # inference.py:160
    if spec.guard_nonfinite and not np.isfinite(tot):                                    # inference.py:162-164
        return -np.inf
    return -0.5 * tot                                                                    # inference.py:166


def within_bounds(spec: ModelSpec, theta) -> bool:
    """inference.py:169-190 (strict box) / TMC1_four_component.py:224-233."""
    theta = np.asarray(theta, dtype=float)
    if not (np.all(spec.lo < theta) and np.all(theta < spec.hi)):
        return False
    v = [theta[i] for i in spec.idx_vlsr]
    if not math.isnan(spec.vlsr_min_sep):
        for a, b in zip(v[:-1], v[1:]):
            if not (a < (b - spec.vlsr_min_sep)):
                return False
    if not math.isnan(spec.vlsr_max_sep):
        for a, b in zip(v[:-1], v[1:]):
            if not (b < (a + spec.vlsr_max_sep)):
                return False
    return True


def lnprior(spec: ModelSpec, theta, prior_stds, prior_means):
    """inference.py:193-236 / TMC1_four_component.py:237-268: Gaussian priors on
    every parameter except the column densities; the vlsr and dV widths are
    replaced by 0.8*mean_dV and 0.3*mean_dV."""
    theta = np.asarray(theta, dtype=float)
    mu = np.asarray(prior_means, dtype=float); sd = np.array(prior_stds, dtype=float)
    for i in spec.idx_vlsr:
        sd[i] = mu[spec.idx_dv] * 0.8
    sd[spec.idx_dv] = mu[spec.idx_dv] * 0.3
    if not within_bounds(spec, theta):
        return -np.inf
    ncol_idx = {i for row in spec.idx_ncol for i in row}
    tot = 0.0
    for i in range(spec.ndim):
        if i in ncol_idx:
            continue
        tot += np.log(1.0 / (np.sqrt(2 * np.pi) * sd[i])) - 0.5 * ((theta[i] - mu[i]) ** 2 / sd[i] ** 2)
    return tot


def lnprob(spec, cats, datagrid, theta, prior_stds, prior_means, windowed=False):
    """inference.py:239-246 / TMC1_four_component.py:272-276."""
    lp = lnprior(spec, theta, prior_stds, prior_means)
    if not np.isfinite(lp):
        return -np.inf
    ll = lnlike(spec, cats, datagrid, theta, windowed=windowed)
    if spec.guard_nonfinite and not np.isfinite(ll):
        return -np.inf
    return lp + ll


# --------------------------------------------------------------------------
# data reduction -- inference.py:108-124, 256-303 ("next" row N3)
# --------------------------------------------------------------------------
def calc_noise_std(intensity, threshold=3.5):
    """inference.py:108-124 (three identical passes against the *initial*
    mean/std; masks [chan-3, chan+3) of a copy)."""
    dummy = np.copy(intensity); noise = np.copy(intensity)
    with np.errstate(all="ignore"):
        mean0 = np.nanmean(dummy); std0 = np.nanstd(dummy)
        nm = ns = np.nan
        for _ in range(3):
            for ch in np.where(dummy - mean0 < (-std0 * threshold))[0]:
                noise[max(0, ch - 3): ch + 3] = np.nan
            for ch in np.where(dummy - mean0 > (std0 * threshold))[0]:
                noise[max(0, ch - 3): ch + 3] = np.nan
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                nm = np.nanmean(noise); ns = np.nanstd(np.real(noise))
    return nm, ns


def reduce_spectrum(freqs, intensity, restfreqs, int_sim, aligned_velocity, block_interlopers=True):
    """inference.py:256-303."""
    rel_f = np.zeros(freqs.shape); rel_i = np.zeros(intensity.shape); rel_e = np.zeros(freqs.shape)
    covered = []
    peak = np.max(int_sim)
    for i, rf in enumerate(restfreqs):
        if int_sim[i] > 0.05 * peak:                                                    # inference.py:272-273
            shift = aligned_velocity                                                    # 274: shift None -> aligned
            vel = (rf - freqs) / rf * ckm + shift
            locs = np.where((vel < (aligned_velocity + 1.5)) & (vel > (aligned_velocity - 1.5)))
            if locs[0].size != 0:
                _, nstd = calc_noise_std(intensity[locs])
                if block_interlopers and (np.max(intensity[locs]) > 3.5 * nstd):        # inference.py:279
                    continue
                covered.append(i)
                rel_f[locs] = freqs[locs]; rel_i[locs] = intensity[locs]
                rel_e[locs] = np.sqrt(nstd ** 2 + (intensity[locs] * 0.1) ** 2)         # inference.py:290
    keep = rel_f > 0                                                                    # inference.py:298
    return rel_f[keep], rel_i[keep], rel_e[keep], np.array(covered, dtype=int)


# --------------------------------------------------------------------------
# sampler -- emcee 3.1.6 StretchMove / RedBlueMove (third-party, absent here;
# restated from the published algorithm, SURVEY.md 3.5 + Appendix C; UNPINNED)
# --------------------------------------------------------------------------
def stretch_move_step(coords, log_probs, log_prob_fn, rng: np.random.RandomState, a: float = 2.0):
    """One ensemble step.  coords (nw, ndim), log_probs (nw,).  log_prob_fn is
    vectorised: (n, ndim) -> (n,).  Returns new coords, log_probs, accepted."""
    nw, ndim = coords.shape
    coords = coords.copy(); log_probs = log_probs.copy()
    accepted = np.zeros(nw, dtype=bool)
    all_inds = np.arange(nw)
    inds = all_inds % 2
    rng.shuffle(inds)                                   # randomize_split=True
    for split in range(2):
        S1 = inds == split
        sets = [coords[inds == j] for j in range(2)]
        s = sets[split]; c = sets[1 - split]
        Ns, Nc = len(s), len(c)
        zz = ((a - 1.0) * rng.rand(Ns) + 1) ** 2.0 / a
        factors = (ndim - 1.0) * np.log(zz)
        rint = rng.randint(Nc, size=(Ns,))
        q = c[rint] - (c[rint] - s) * zz[:, None]
        new_lp = log_prob_fn(q)
        lnpdiff = factors + new_lp - log_probs[all_inds[S1]]
        acc = lnpdiff > np.log(rng.rand(Ns))
        idx = all_inds[S1][acc]
        coords[idx] = q[acc]; log_probs[idx] = new_lp[acc]; accepted[idx] = True
    return coords, log_probs, accepted
