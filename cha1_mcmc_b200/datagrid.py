"""Data reduction before the path: spectrum -> sparse "datagrid".

What the reference's ``calc_noise_std`` / ``read_file`` / the datagrid writer of ``init_setup`` produce
(inference.py:108-124, 256-303, 336-340), computed with array operations: a clipped-noise estimate by interval
dilation, and per-line velocity windows bracketed by binary search on the sorted channel axis (the reference forms
the velocity of every channel for every line: O(L*C), hours at GOTHAM size).  The selection rules -- which channels
belong to a line, which lines are vetoed, which sigma a channel ends up with -- are the reference's, rule for rule;
``tests/test_host.py`` holds the output to the reference's own ``init_setup`` on the DSN sample and to an
element-by-element restatement on random spectra.  The only arithmetic-heavy ingredient, the stick spectrum
``int_sim`` of MolSim(C=3.4e12, dV=0.89, T=7), comes from the device (``cha_stick_spectrum``)."""
from __future__ import annotations

import warnings

import numpy as np

from .constants import ckm

CLIP_RADIUS = 3          # channels blanked around an outlier: [c - 3, c + 3)   (inference.py:114)
WINDOW_KMS = 1.5         # half-width of a line's velocity window               (inference.py:274-275)
STRONG_FRACTION = 0.05   # lines weaker than 5 % of the strongest are ignored   (inference.py:272)
INTERLOPER_SIGMA = 3.5   # veto threshold on the window's maximum               (inference.py:279)


def calc_noise_std(intensity, threshold=3.5):
    """(mean, std) of the channels that survive the reference's clipping (inference.py:108-124).

    The reference runs three passes, but every pass tests the *unmodified* input against the *initial* mean and
    standard deviation, so all three blank the same channels; one pass gives the same answer.  An outlier at
    channel c blanks the half-open interval [c - 3, c + 3); the union of those intervals is built with a
    difference array instead of one slice assignment per outlier."""
    v = np.asarray(intensity, dtype=float)
    n = v.size
    if n == 0:
        return np.nan, np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        centre, spread = np.nanmean(v), np.nanstd(v)
        # strict comparisons on both sides; a NaN mean/std selects nothing, as in the reference
        hits = np.flatnonzero((v - centre < -spread * threshold) | (v - centre > spread * threshold))
        if hits.size:
            edge = np.zeros(n + 1, dtype=np.int64)
            np.add.at(edge, np.maximum(hits - CLIP_RADIUS, 0), 1)
            np.add.at(edge, np.minimum(hits + CLIP_RADIUS, n), -1)
            quiet = v[np.cumsum(edge[:-1]) == 0]
        else:
            quiet = v
        if quiet.size == 0:
            return np.nan, np.nan
        return np.nanmean(quiet), np.nanstd(quiet)


def line_windows(freqs, restfreqs, offset, aligned_velocity):
    """Channel indices (into ``freqs``) of every rest frequency's velocity window, as a list of index arrays.

    A channel x belongs to line f when  aligned - 1.5 < (f - x)/f*ckm + offset < aligned + 1.5  (strict, evaluated
    with exactly this expression).  The channel axis is sorted once; a slightly widened bracket from two binary
    searches bounds the candidates and the exact test is applied to that slice only."""
    freqs = np.asarray(freqs, dtype=float)
    rest = np.asarray(restfreqs, dtype=float)
    order = np.argsort(freqs, kind="stable")
    xs = freqs[order]
    reach = (abs(offset - aligned_velocity) + WINDOW_KMS) / ckm * np.abs(rest) * 1.001 + 1e-9
    first = np.searchsorted(xs, rest - reach, "left")
    last = np.searchsorted(xs, rest + reach, "right")
    out = []
    for f, a, b in zip(rest, first, last):
        vel = (f - xs[a:b]) / f * ckm + offset
        inside = (vel < aligned_velocity + WINDOW_KMS) & (vel > aligned_velocity - WINDOW_KMS)
        out.append(np.sort(order[a:b][inside]))
    return out


def reduce_spectrum(freqs, intensity, restfreqs, int_sim, aligned_velocity, shift=None, GHz=False,
                    block_interlopers=True, log=None):
    """The reduced spectrum ``read_file`` returns (inference.py:256-303):
    (relevant_freqs, relevant_intensity, relevant_yerrs, covered_trans).

    Lines are visited in catalog order; a channel shared by two accepted lines keeps the sigma of the LATER one
    (the reference overwrites), a vetoed line contributes nothing."""
    freqs = np.asarray(freqs, dtype=float) * (1000.0 if GHz else 1.0)
    intensity = np.asarray(intensity, dtype=float)
    int_sim = np.asarray(int_sim, dtype=float)
    restfreqs = np.asarray(restfreqs, dtype=float)
    say = log or (lambda _msg: None)
    offset = shift if shift else aligned_velocity
    strong = np.flatnonzero(int_sim > STRONG_FRACTION * np.max(int_sim)) if int_sim.size else np.empty(0, dtype=int)
    windows = line_windows(freqs, restfreqs[strong], offset, aligned_velocity)
    kept = np.zeros(freqs.size, dtype=bool)
    sigma = np.zeros(freqs.size)
    covered = []
    for i, chans in zip(strong, windows):
        tag = f"{restfreqs[i]:10.4f} MHz  |  "
        if chans.size == 0:
            say(tag + "No data.")
            continue
        seen = intensity[chans]
        _, noise = calc_noise_std(seen)
        if block_interlopers and np.max(seen) > INTERLOPER_SIGMA * noise:
            say(tag + "Interloping line detected.")
            continue
        say(tag + "Line found.")
        covered.append(int(i))
        kept[chans] = freqs[chans] > 0                       # the reference selects on relevant_freqs > 0
        sigma[chans] = np.sqrt(noise ** 2 + (seen * 0.1) ** 2)      # inference.py:290, same operation order
    return freqs[kept], intensity[kept], sigma[kept], np.array(covered, dtype=int)


def save_datagrid(path, freqs, ints, yerrs, covered_trans):
    """The reference's on-disk datagrid: a 4-element object array (inference.py:337-340)."""
    datagrid = np.empty(4, dtype=object)
    datagrid[0], datagrid[1], datagrid[2], datagrid[3] = freqs, ints, yerrs, np.asarray(covered_trans, dtype=int)
    np.save(path, datagrid, allow_pickle=True)
    return datagrid
