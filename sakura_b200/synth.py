"""Seeded synthetic single-dish spectra shared bit-for-bit by the oracle and the GPU path
(SURVEY.md section 8(d)): quintic continuum + optional standing-wave ripple + Gaussian
emission lines + single-channel spikes + N(0,1) noise; masks with flagged edges, 2 % random
flags and one flagged block; plus the fixed edge-case rows the reference's tests exercise
(test/lsq.cc:474-488: minimum effective data, too little data, no data, constant data,
a NaN hidden under the mask).  The 'clipping drives the row below the number of
coefficients' case needs its own tiny clip threshold (test/lsq.cc:486-488) and lives in the tests.

numpy Philox streams: seed for data, seed+1 for masks, seed+2 for line/spike placement.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 20240611


def make_spectra(num_spectra: int, num_data: int, seed: int = DEFAULT_SEED, ripple: bool = False,
                 edge: int = 32, flag_fraction: float = 0.02, max_block: int = 256,
                 lines: bool = True):
    """Returns (data float32 [R, n], mask bool [R, n])."""
    R, n = num_spectra, num_data
    g_data = np.random.Generator(np.random.Philox(seed))
    g_mask = np.random.Generator(np.random.Philox(seed + 1))
    g_feat = np.random.Generator(np.random.Philox(seed + 2))
    x = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    coef = g_data.uniform(-5.0, 5.0, size=(R, 6))
    y = np.zeros((R, n), dtype=np.float64)
    for k in range(5, -1, -1):
        y = y * x[None, :] + coef[:, k:k + 1]
    if ripple:
        amp = g_data.uniform(0.5, 2.0, size=(R, 1))
        kk = g_data.integers(1, 21, size=(R, 1)).astype(np.float64)
        ph = g_data.uniform(0.0, 2.0 * np.pi, size=(R, 1))
        y += amp * np.sin(2.0 * np.pi * kk * x[None, :] + ph)
    y += g_data.standard_normal(size=(R, n))
    if lines and n >= 64:
        chan = np.arange(n, dtype=np.float64)[None, :]
        nl = g_feat.integers(0, 4, size=R)
        for j in range(3):
            on = (nl > j)[:, None]
            peak = g_feat.uniform(5.0, 50.0, size=(R, 1))
            fwhm = g_feat.uniform(4.0, 64.0, size=(R, 1))
            cen = g_feat.uniform(0.0, n - 1.0, size=(R, 1))
            sig = fwhm / 2.354820045
            y += np.where(on, peak * np.exp(-0.5 * ((chan - cen) / sig) ** 2), 0.0)
        ns = g_feat.integers(0, 9, size=R)
        for j in range(8):
            on = ns > j
            pos = g_feat.integers(0, n, size=R)
            amp = g_feat.uniform(10.0, 100.0, size=R) * g_feat.choice([-1.0, 1.0], size=R)
            rows = np.nonzero(on)[0]
            y[rows, pos[rows]] += amp[rows]
    data = y.astype(np.float32)

    mask = g_mask.random(size=(R, n)) >= flag_fraction
    e = min(edge, n // 8)
    if e > 0:
        mask[:, :e] = False
        mask[:, n - e:] = False
    blk = g_mask.integers(0, min(max_block, max(n // 8, 1)) + 1, size=R)
    start = g_mask.integers(0, n, size=R)
    idx = np.arange(n)[None, :]
    mask &= ~((idx >= start[:, None]) & (idx < (start + blk)[:, None]))
    return data, mask


def add_edge_case_rows(data: np.ndarray, mask: np.ndarray, num_coeff: int, first_row: int = 0):
    """Overwrites rows first_row .. first_row+5 with the fixed edge cases; returns a dict
    row-name -> row index.  Needs at least 6 rows and num_data >= 4*num_coeff."""
    R, n = data.shape
    names = {}
    r = first_row
    K = num_coeff
    # all-true mask
    mask[r, :] = True
    names["all_true"] = r
    r += 1
    # exactly K unmasked channels, spread out
    mask[r, :] = False
    mask[r, np.linspace(1, n - 2, K).astype(int)] = True
    names["exactly_K"] = r
    r += 1
    # K-1 unmasked -> kNotEnoughData before the loop
    mask[r, :] = False
    mask[r, np.linspace(1, n - 2, K - 1).astype(int) if K > 1 else []] = True
    names["K_minus_1"] = r
    r += 1
    # all false
    mask[r, :] = False
    names["all_false"] = r
    r += 1
    # constant data
    data[r, :] = 10.0
    names["constant"] = r
    r += 1
    # a row with a NaN hidden under the mask (must not leak)
    data[r, n // 2] = np.nan
    mask[r, n // 2] = False
    names["nan_under_mask"] = r
    return names
