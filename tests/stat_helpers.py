"""Two-sample statistics for the distributional parity tests (GPU vs the unmodified reference).

Every assertion built on `z_two_sample` compares the MEANS of a per-trajectory statistic over n_a GPU and n_b
reference trajectories (both >= 512), with the standard error taken from the pooled per-trajectory variance.
By the CLT the statistic is N(0, 1) under the null; `Z_BOUND` = 5.3 puts the two-sided false-failure
probability of one assertion at 1.2e-7 for an exactly normal z, and below 1e-6 with the skew that means over
>= 512 draws of the heaviest-tailed statistic used here (band power, skewness 1.1) retain.  Calibration on the
CPU (tools/calibrate_stochastic.py): 4x10^5 bootstrap resamples of (4096, 1024) from 12288 reference
trajectories of the thermal-FMR test give std(z) = 1.000 and P(|z| > 4) = 7e-5 for all three of its statistics
(normal: 6.3e-5), max |z| = 5.39.
"""
import math

import numpy as np

Z_BOUND = 5.3


def z_two_sample(a, b):
    """z of mean(a) - mean(b) with the pooled standard error.  Returns (z, estimate, standard error)."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    na, nb = a.size, b.size
    pooled = (a.var() * na + b.var() * nb) / (na + nb)
    se = math.sqrt(pooled * (1.0 / na + 1.0 / nb))
    est = a.mean() - b.mean()
    return est / se if se > 0 else 0.0, est, se


def z_ratio(a, b):
    """z of mean(a) / mean(b) - 1 for positive per-trajectory statistics."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    na, nb = a.size, b.size
    ra, rb = a / a.mean(), b / b.mean()
    pooled = (ra.var() * na + rb.var() * nb) / (na + nb)
    se = math.sqrt(pooled * (1.0 / na + 1.0 / nb))
    est = a.mean() / b.mean() - 1.0
    return est / se if se > 0 else 0.0, est, se


def z_proportions(ka, na, kb, nb):
    """two-proportion z-test (pooled)"""
    pa, pb = ka / na, kb / nb
    p = (ka + kb) / (na + nb)
    se = math.sqrt(max(p * (1 - p), 1e-6) * (1.0 / na + 1.0 / nb))
    return (pa - pb) / se, pa - pb, se


def psd_slopes(x, lo_div=64, hi_div=4):
    """per-series least-squares slope of log PSD vs log f over bins [n_f/lo_div, n_f/hi_div); x: [series, n]"""
    x = np.asarray(x, dtype=np.float64)
    p = np.abs(np.fft.rfft(x - x.mean(axis=1, keepdims=True), axis=1)) ** 2
    nf = p.shape[1]
    sel = slice(max(1, nf // lo_div), max(2, nf // hi_div))
    lf = np.log(np.arange(nf)[sel])
    lp = np.log(np.maximum(p[:, sel], 1e-300))
    lf0 = lf - lf.mean()
    return (lp * lf0).sum(axis=1) / (lf0 * lf0).sum()
