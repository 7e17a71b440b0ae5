"""CPU restatement of the reference's 1/f^alpha generators -- TEST INFRASTRUCTURE ONLY (imported by tests/
and by the oracle adapter in tests/helpers.py; never by the product, which generates the buffers on the GPU,
cmtj_b200/csrc/onef.cu).

Follows BufferedAlphaNoise::fillBuffer (core/noise.hpp:190-227) and VectorAlphaNoise::tickVector
(:291-301): Kasdin's filter in float32 through a 2n-point complex FFT, product of the two spectra on the
lower half only with both end bins halved, unnormalised inverse, real part / n.  kissfft (vendored,
third_party/kissfft) is replaced by numpy's FFT rounded to complex64.  RNG streams differ from the
reference's mt19937(random_device): pinned spectrally against the unmodified reference in
tests/test_noise_host.py.
"""
import numpy as np


def kasdin_weights(n: int, alpha: float) -> np.ndarray:
    """noise.hpp:173-177: h0 = 1, h_i = h_{i-1} * (float)(alpha/2 + (float)(i-1)) / (float)i"""
    h = np.empty(n, dtype=np.float32)
    h[0] = 1.0
    if n > 1:
        i = np.arange(1, n, dtype=np.float32)
        w = (0.5 * float(alpha) + (i - np.float32(1.0)).astype(np.float64)).astype(np.float32) / i
        h[1:] = np.cumprod(w, dtype=np.float32)
    return h


def alpha_noise_buffer(n: int, alpha: float, std: float, rng=None) -> np.ndarray:
    """One ``fillBuffer``: returns ``result`` (float32, length n).  rng: a numpy Generator (a private one when None)."""
    if n <= 0:
        return np.zeros(0, dtype=np.float32)
    rng = np.random.default_rng() if rng is None else rng
    white = np.zeros(2 * n, dtype=np.complex64)
    white[:n] = rng.normal(0.0, std, n).astype(np.float32) if std > 0 else 0.0
    col = np.zeros(2 * n, dtype=np.complex64)
    col[:n] = kasdin_weights(n, alpha)
    W = np.fft.fft(white).astype(np.complex64)
    Cc = np.fft.fft(col).astype(np.complex64)
    P = np.zeros(2 * n, dtype=np.complex64)
    P[:n] = Cc[:n] * W[:n]
    P[0] = P[0] / 2.0
    P[n - 1] = P[n - 1] / 2.0
    y = (np.fft.ifft(P) * (2 * n)).astype(np.complex64)  # kissfft's inverse is unnormalised
    return (y[:n].real / np.float32(n)).astype(np.float32)


def vector_alpha_noise(n: int, alpha: float, std: float, scale: float, axis, rng=None) -> np.ndarray:
    """All n ticks of ``VectorAlphaNoise::tickVector`` at once: float64 [3][n], scaled; normalised to unit
    length per sample when axis == all (noise.hpp:291-301)."""
    ax = int(axis)
    out = np.zeros((3, n), dtype=np.float64)
    if ax == 3:
        for c in range(3):
            out[c] = alpha_noise_buffer(n, alpha, std, rng)
        norm = np.sqrt(out[0] ** 2 + out[1] ** 2 + out[2] ** 2)
        nz = norm != 0
        out[:, nz] /= norm[nz]
    elif ax in (0, 1, 2):
        out[ax] = alpha_noise_buffer(n, alpha, std, rng)
    else:
        raise RuntimeError("Invalid axis specified: " + str(ax))
    return out * scale


def junction_buffer(layers, n_layers: int, total_iterations: int, n_instances: int = 1, rng=None) -> np.ndarray:
    """The ``cmtjb_run_args_t.onef_buffer`` a junction run needs (getSolver, junction.hpp:1905-1916): a fresh
    VectorAlphaNoise per layer with scaleNoise != 0 and per instance; float64 [n_layers*3][iterations][B]."""
    buf = np.zeros((n_layers * 3, max(total_iterations, 1), n_instances), dtype=np.float64)
    for i in range(n_layers):
        L = layers[i]
        if L.noiseScale != 0:
            for b in range(n_instances):
                buf[3 * i : 3 * i + 3, :total_iterations, b] = vector_alpha_noise(
                    total_iterations, L.noiseAlpha, L.noiseStd, L.noiseScale, L.noiseAxis, rng
                )
    return buf
