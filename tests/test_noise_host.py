"""CPU restatement of the Kasdin 1/f^alpha generator (oracle/noise_oracle.py) against the reference's
BufferedAlphaNoise / VectorAlphaNoise (core/noise.hpp:138-308, kissfft in float32): RNG streams
differ, so the comparison is spectral.  CPU only."""
import numpy as np
import pytest

import os
import sys

import cmtj_b200
import helpers

sys.path.insert(0, helpers.ORACLE_DIR)
import noise_oracle as noise  # noqa: E402


def _slope(x):
    """least-squares slope of log PSD vs log f over the middle decade of the spectrum"""
    p = np.abs(np.fft.rfft(x, axis=-1)) ** 2
    p = p.reshape(-1, p.shape[-1]).mean(axis=0)
    f = np.arange(len(p))
    sel = slice(len(p) // 64, len(p) // 4)
    return np.polyfit(np.log(f[sel]), np.log(p[sel]), 1)[0]


@pytest.mark.parametrize("alpha", [0.0, 1.0, 2.0])
def test_spectral_slope_and_power_match_reference(ref, alpha):
    n, reps, std = 4096, 24, 0.7
    rng = np.random.default_rng(0)
    ours = np.stack([noise.alpha_noise_buffer(n, alpha, std, rng) for _ in range(reps)]).astype(np.float64)
    theirs = np.zeros((reps, n))
    for r in range(reps):
        g = ref.noise.BufferedAlphaNoise(n, alpha, std, 1.0)
        theirs[r] = [g.tick() for _ in range(n)]
    assert abs(_slope(ours) - (-alpha)) < 0.15
    assert abs(_slope(ours) - _slope(theirs)) < 0.15
    assert abs(np.log(ours.var() / theirs.var())) < 0.5 + 0.5 * alpha  # 1/f^2 power is dominated by the lowest bins


def test_vector_noise_is_unit_length_times_scale(ref):
    v = noise.vector_alpha_noise(512, 1.0, 0.3, 2.5, cmtj_b200.Axis.all, np.random.default_rng(1))
    assert np.allclose(np.linalg.norm(v, axis=0), 2.5)
    g = ref.noise.VectorAlphaNoise(512, 1.0, 0.3, 2.5, ref.Axis.all)
    w = np.array([g.tickVector().tolist() for _ in range(512)])
    assert np.allclose(np.linalg.norm(w, axis=1), 2.5, rtol=1e-5)
    x = noise.vector_alpha_noise(64, 1.0, 0.3, 2.0, cmtj_b200.Axis.yaxis, np.random.default_rng(1))
    assert np.all(x[0] == 0) and np.all(x[2] == 0) and np.any(x[1] != 0)
