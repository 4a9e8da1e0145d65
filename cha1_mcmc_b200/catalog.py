"""Host data layer: SPCAT/CDMS ``.cat`` parser and partition-function dispatch.

Mirrors ``MolCat`` (spectral_simulator/classes.py:16-288): same attribute names
(``frequency, error, logint, dof, elower, gup, tag, qnformat, qn1..qn12, qns``)
and the same text-to-integer quirks (functions.py:330-335 ``fix_pm``,
340-501 ``fix_qn``), because the unique-state table of the partition function
depends on them.  What is NOT mirrored: the O(N^2) ``glow`` match
(classes.py:100-110, 36 s for 1-cyanonaphthalene) -- glow and gup cancel in the
optical depth (classes.py:349-354) so the device never needs them -- and the
per-call ``sijmu``/``aij`` arrays, which the device computes once
(``cha_set_molecule`` -> ``catalog_terms_kernel``).

Text parsing stays on the host; arithmetic goes to the GPU.
"""
from __future__ import annotations

import gzip
import os
from typing import Optional, Tuple

import numpy as np

# partition-function kinds (include/chalte.h)
Q_POLY, Q_LIN, Q_POW, Q_SUM = 0, 1, 2, 3


def _letter_value(text: str) -> int:
    """functions.py:340-501: an alphabetic character selects 100 + 10*(letter index)
    (upper and lower case alike) and the *second character* of the field supplies the
    units; a field with no letter (e.g. empty) becomes 0."""
    val = 0
    for ch in text:
        if ch.isalpha() and ch.isascii():
            val = 100 + 10 * (ord(ch.upper()) - ord("A")) + int(text[1])
    # the reference tests A..Z then a..z with successive ifs: with a single letter per field
    # (all SPCAT encodings) the result is the same
    return val


def _qn_column(col: np.ndarray) -> np.ndarray:
    """One quantum-number column: classes.py:180-286."""
    if np.any(col == "+") or np.any(col == "-"):           # fix_pm only for columns holding a bare sign
        col = col.copy()
        col[col == ""] = "0"
        col[col == "+"] = "1"
        col[col == "-"] = "2"
    out = np.empty(col.size, dtype=np.int64)
    for r, s in enumerate(col):
        try:
            out[r] = int(s)
        except ValueError:
            out[r] = _letter_value(s)
    return out


def resolve_q_mode(catalog_file: str, dispatch_on: str = "basename") -> Tuple[int, Tuple[float, ...]]:
    """Which branch of ``calc_q`` (functions.py:136-325) this catalog takes, resolved once.

    The reference tests substrings of the *full lower-cased path* on every call; a directory
    such as ``.../hc3n_runs/benzonitrile.cat`` would silently select the HC3N fit.  The default
    here uses the basename; ``dispatch_on="path"`` reproduces the reference literally.

    Returns (kind, params) in the C-ABI encoding of include/chalte.h:
      Q_POLY: p[n]                      Q = sum p[n] T^n
      Q_LIN : (a, b, scale, div)        Q = scale*(a T + b) or (a T + b)/div when div != 0
      Q_POW : (a, p, b, has_b)          Q = a T^p (+ b)
      Q_SUM : ()                        explicit state sum
    Spelling traps are preserved: ``1-cyanonaphthalene`` / ``2-cyanonaphthalene`` /
    ``acenapthylene`` in the reference never match the shipped ``1-cyanonapthalene.cat``,
    ``2-cyanonapthalene.cat``, ``acenaphthylene.cat`` -> those take the state sum."""
    f = (catalog_file if dispatch_on == "path" else os.path.basename(catalog_file)).lower()
    if f.endswith(".gz"):
        f = f[:-3]
    has = lambda s: s in f  # noqa: E731
    hfs = has("hfs")
    if has("n2h+_hfs.cat"):
        return Q_POLY, (3.32018827e+00, 4.01951955e+00, 3.28722820e-05, -3.13420474e-08)
    if has("acetone.cat"):
        return Q_POLY, (16431.0, -2728.3, 245.28, -5.5477, 0.05471337, -0.00021050085, 2.91296 * 10 ** (-7))
    if has("sh.cat"):
        return Q_POLY, (15.357239728157400, 0.069272946237033, 0.002288160909445, -0.000008528126823, 0.000000012549467)
    if has("h2s.cat"):
        return Q_POLY, (-1.764494755639740, 0.507648423477309, 0.005498622332982, -0.000004859941547)
    if has("hcn.cat"):
        return Q_POLY, (.386550361, 1.48629408, -1.15188755 * 10 ** -3, 4.62476813 * 10 ** -6, -1.64946939 * 10 ** -9)
    if any(has(s) for s in ("methanol.cat", "ch3oh.cat", "ch3oh_v0.cat", "ch3oh_v1.cat", "ch3oh_v2.cat", "ch3oh_vt.cat")):
        return Q_POLY, (-1.25670, 4.39632 * 10 ** -1, 2.05911 * 10 ** -1, -1.83807 * 10 ** -3, 1.27624 * 10 ** -5,
                        -4.04024 * 10 ** -8, 4.83410 * 10 ** -11)
    if has("13methanol.cat") or has("13ch3oh.cat"):
        return Q_POLY, (-31.876881967, 4.317920731, 0.076540934, 0.000050130)
    if has("c2n.cat") or has("ccn.cat"):
        return Q_POLY, (22.55770, 7.135161, 0.1837397, -1.40473 * 10 ** (-3), 5.99936 * 10 ** (-6),
                        -1.324086 * 10 ** (-8), 1.173755 * 10 ** (-11))
    if has("ch2nh.cat"):
        return Q_POW, (1.2152, 1.4863, 0.0, 0.0)
    if has("c033502.cat"):
        return Q_POW, (0.399272, 1.756329, 0.0, 0.0)
    # cyanopolyynes / isocyanides: functions.py:173-210
    x3 = lambda a, b: (Q_LIN, (a, b, 3.0 if hfs else 1.0, 0.0))       # noqa: E731  "3*(aT+b)" with hfs
    d3 = lambda a, b: (Q_LIN, (a, b, 1.0, 0.0 if hfs else 3.0))       # noqa: E731  "(aT+b)/3" without hfs
    if has("hc3n"):
        return x3(4.581898, 0.2833)
    if has("hc2nc_hfs"):
        return Q_LIN, (12.58340, 1.0604, 1.0, 0.0)
    if has("hc5n"):
        return x3(15.65419, 0.2214)
    if has("hc4nc"):
        return d3(44.62171, 0.6734)
    if has("hc7n"):
        return x3(36.94999, 0.1356)
    if has("hc6nc"):
        return d3(107.3126, 1.2714)
    if has("hc9n"):
        return x3(71.7308577, 0.02203968)
    if has("hc11n.cat") and not hfs:
        return Q_LIN, (123.2554, 0.1381, 1.0, 0.0)
    if has("hc11n") and hfs:
        return Q_LIN, (123.2554, 0.1381, 3.0, 0.0)
    for needle, a, p, b in (("propargylcyanide", 41.542, 1.5008, 0.0), ("pyrrole", 27.727, 1.4752, 0.0),
                            ("cyclopropylcyanide_hfs", 38.199, 1.4975, 0.0), ("pyridine", 50.478, 1.4955, 0.0),
                            ("1-cyanonaphthalene", 560.39, 1.4984, 0.0), ("2-cyanonaphthalene", 562.57, 1.4993, 0.0),
                            ("furan", 33.725, 1.4982, 0.0), ("phenol", 264.20, 1.4984, 0.0),
                            ("benzaldehyde", 53.798, 1.4997, 0.0), ("anisole", 54.850, 1.4992, 0.0),
                            ("azulene", 96.066, 1.4988, 0.0), ("acenaphthene", 161.29, 1.4994, 0.0),
                            ("acenapthylene", 151.58, 1.4988, 0.0), ("fluorene", 219.51, 1.4996, 0.0),
                            ("benzonitrile", 25.896, 1.4998, 0.38109)):
        if has(needle):
            return Q_POW, (a, p, b, 1.0 if b != 0.0 else 0.0)
    return Q_SUM, ()


class MolCat:
    """Parsed catalog.  Constructor signature follows classes.py:19 (``name, catalog_file``)."""

    def __init__(self, name: str, catalog_file: str, format: str = "spcat", CT: float = 300,
                 q_dispatch_on: str = "basename"):
        self.name = name
        self.catalog_file = catalog_file
        self.format = format
        self.CT = CT
        self._read(q_dispatch_on)

    def _read(self, q_dispatch_on):
        path = self.catalog_file
        if not os.path.exists(path) and os.path.exists(path + ".gz"):
            path = path + ".gz"
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as fh:
            rows = fh.readlines()
        n = len(rows)
        if n == 0:
            raise ValueError(f"empty catalog {self.catalog_file}")
        self.frequency = np.array([float(x[:13]) for x in rows])                    # classes.py:155
        self.error = np.array([float(x[13:21]) for x in rows])
        self.logint = np.array([float(x[21:29]) for x in rows])
        self.dof = np.array([int(x[29:31]) for x in rows], dtype=np.int64)
        self.elower = np.array([float(x[31:41]) for x in rows])
        gup = np.empty(n, dtype=np.int64)
        for r, x in enumerate(rows):                                                # classes.py:160-163
            g = x[41:44]
            try:
                gup[r] = int(g)
            except ValueError:
                gup[r] = _letter_value(g)
        self.gup = gup
        self.tag = np.array([int(x[44:51]) for x in rows], dtype=np.int64)
        self.qnformat = np.array([int(x[51:55]) for x in rows], dtype=np.int64)
        qn = np.empty((n, 12), dtype=np.int64)
        for q in range(12):
            if q < 11:
                col = np.array([x[55 + 2 * q:57 + 2 * q].strip() for x in rows], dtype=object)
            else:
                col = np.array([x[77:].strip() for x in rows], dtype=object)        # classes.py:178
            qn[:, q] = _qn_column(col)
        self.qn = qn
        for q in range(12):
            setattr(self, f"qn{q + 1}", qn[:, q])
        self.qns = min(int(str(int(self.qnformat[0]))[-1:]), 6)                     # classes.py:116-122
        self.eupper = self.elower + self.frequency / 29979.2458                     # classes.py:90
        self.intensity = 10 ** self.logint                                          # classes.py:126-128
        self.q_kind, self.q_params = resolve_q_mode(self.catalog_file, q_dispatch_on)
        self.state_g = np.zeros(0)
        self.state_E = np.zeros(0)
        if self.q_kind == Q_SUM:
            self._unique_states()
        if not np.all(np.diff(self.frequency) >= 0):
            raise ValueError(f"{self.catalog_file}: catalog is not frequency-sorted (all 35 shipped ones are)")

    def _unique_states(self):
        """functions.py:264-317: de-duplicate rows (qn7..qn(6+qns), elower); g = 2*qn7 + 1."""
        cols = [self.qn[:, 6 + q].astype(float) for q in range(self.qns)] + [self.elower]
        arr = np.stack(cols, axis=1)
        uniq = np.unique(arr, axis=0)
        self.state_g = 2.0 * uniq[:, 0] + 1.0
        self.state_E = uniq[:, self.qns].copy()

    # trim_array rule (functions.py:519-534) for one [ll],[ul] chunk: indices [i0, i1)
    def trim_bounds(self, ll: float, ul: float) -> Tuple[int, int]:
        f = self.frequency
        above = np.nonzero(f > ll)[0]
        if above.size:
            i0 = int(above[0])
        elif f[-1] < ll:
            return 0, 0
        else:
            i0 = 0
        above = np.nonzero(f > ul)[0]
        i1 = int(above[0]) if above.size else f.size
        return i0, max(i1, i0)


def find_catalog(cat_folder: str, mol_name: str) -> Optional[str]:
    for ext in (".cat", ".cat.gz"):
        p = os.path.join(cat_folder, mol_name + ext)
        if os.path.exists(p):
            return p
    return None
