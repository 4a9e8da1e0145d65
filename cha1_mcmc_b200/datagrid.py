"""Data reduction before the path: spectrum -> sparse "datagrid".

Host mirror of ``SpectralFitMCMC.calc_noise_std`` / ``read_file`` / the datagrid writer of ``init_setup``
(inference.py:108-124, 256-303, 336-340).  This is orchestration on small arrays; the only arithmetic-heavy
ingredient -- the reference stick spectrum ``int_sim`` of MolSim(C=3.4e12, dV=0.89, T=7) -- comes from the
device (``cha_stick_spectrum``)."""
from __future__ import annotations

import warnings

import numpy as np

from .constants import ckm


def calc_noise_std(intensity, threshold=3.5):
    """inference.py:108-124: three passes that NaN-out [chan-3, chan+3) around every channel deviating by
    more than threshold*std from the mean of the *unmasked* input; returns (mean, std) of what is left."""
    dummy_ints = np.copy(intensity)
    noise = np.copy(intensity)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dummy_mean = np.nanmean(dummy_ints)
        dummy_std = np.nanstd(dummy_ints)
        noise_mean = noise_std = np.nan
        for _ in range(3):
            mask_radius = 3
            for chan in np.where(dummy_ints - dummy_mean < (-dummy_std * threshold))[0]:
                noise[max(0, chan - mask_radius): chan + mask_radius] = np.nan
            for chan in np.where(dummy_ints - dummy_mean > (dummy_std * threshold))[0]:
                noise[max(0, chan - mask_radius): chan + mask_radius] = np.nan
            noise_mean = np.nanmean(noise)
            noise_std = np.nanstd(np.real(noise))
    return noise_mean, noise_std


def reduce_spectrum(freqs, intensity, restfreqs, int_sim, aligned_velocity, shift=None, GHz=False,
                    block_interlopers=True, log=None):
    """inference.py:256-303.  Returns (relevant_freqs, relevant_intensity, relevant_yerrs, covered_trans)."""
    freqs = np.asarray(freqs, dtype=float)
    intensity = np.asarray(intensity, dtype=float)
    if GHz:
        freqs = freqs * 1000.0
    relevant_freqs = np.zeros(freqs.shape)
    relevant_intensity = np.zeros(intensity.shape)
    relevant_yerrs = np.zeros(freqs.shape)
    covered_trans = []
    peak = np.max(int_sim) if len(int_sim) else 0.0
    off = shift if shift else aligned_velocity
    # The reference forms the velocity of EVERY channel for every line (O(L*C): hours at GOTHAM size).  On a
    # frequency-sorted grid the channels that can pass the +-1.5 km/s test are bracketed by a binary search and the
    # same test is applied to that slice only -- identical selection, O(L*(log C + W)).
    ascending = freqs.size > 1 and bool(np.all(freqs[1:] >= freqs[:-1]))
    for i, rf in enumerate(restfreqs):
        if int_sim[i] > 0.05 * peak:                                                    # 5 % of the strongest line
            if ascending:
                half = (abs(off - aligned_velocity) + 1.5) / ckm * abs(rf) * 1.001 + 1e-9
                a = int(np.searchsorted(freqs, rf - half, "left")); b = int(np.searchsorted(freqs, rf + half, "right"))
                vel = (rf - freqs[a:b]) / rf * ckm + off
                sel = np.where((vel < (aligned_velocity + 1.5)) & (vel > (aligned_velocity - 1.5)))[0] + a
                locs = (sel,)
            else:
                vel = (rf - freqs) / rf * ckm + off
                locs = np.where((vel < (aligned_velocity + 1.5)) & (vel > (aligned_velocity - 1.5)))
            if locs[0].size != 0:
                _, noise_std = calc_noise_std(intensity[locs])
                if block_interlopers and (np.max(intensity[locs]) > 3.5 * noise_std):
                    if log:
                        log(f"{rf:10.4f} MHz  |  Interloping line detected.")
                else:
                    covered_trans.append(i)
                    if log:
                        log(f"{rf:10.4f} MHz  |  Line found.")
                    relevant_freqs[locs] = freqs[locs]
                    relevant_intensity[locs] = intensity[locs]
                    relevant_yerrs[locs] = np.sqrt(noise_std ** 2 + (intensity[locs] * 0.1) ** 2)
            elif log:
                log(f"{rf:10.4f} MHz  |  No data.")
    mask = relevant_freqs > 0
    return relevant_freqs[mask], relevant_intensity[mask], relevant_yerrs[mask], np.array(covered_trans, dtype=int)


def save_datagrid(path, freqs, ints, yerrs, covered_trans):
    """The reference's on-disk datagrid: a 4-element object array (inference.py:337-340)."""
    datagrid = np.empty(4, dtype=object)
    datagrid[0], datagrid[1], datagrid[2], datagrid[3] = freqs, ints, yerrs, np.asarray(covered_trans, dtype=int)
    np.save(path, datagrid, allow_pickle=True)
    return datagrid
