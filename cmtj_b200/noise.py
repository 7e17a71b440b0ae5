"""1/f^alpha noise generators: ``BufferedAlphaNoise`` / ``VectorAlphaNoise``
(core/noise.hpp:138-308, python names cmtj.cpp:449-462).

The reference fills one buffer of ``bufferSize = totalIterations`` samples per axis at the
first tick (Kasdin's filter, float32, through kissfft).  It is set-up work, not the
per-timestep path, so the buffer is generated here on the host with numpy's FFT (float32
data, same filter and the same one-sided-spectrum quirks) and handed to the device as a
read-only array (``cmtjb_run_args_t.onef_buffer``).  RNG streams differ from the
reference's ``mt19937(random_device)``: parity is distributional.
"""
from __future__ import annotations

import numpy as np


def _kasdin_weights(n: int, alpha: float) -> np.ndarray:
    # noise.hpp:173-177: h0 = 1, h_i = h_{i-1} * (alpha/2 + i - 1) / i, in float32
    h = np.empty(n, dtype=np.float32)
    h[0] = 1.0
    if n > 1:
        i = np.arange(1, n, dtype=np.float32)
        w = ((np.float32(0.5 * alpha) + (i - np.float32(1.0))) / i).astype(np.float32)
        h[1:] = np.cumprod(w, dtype=np.float32)
    return h


def alpha_noise_buffer(n: int, alpha: float, std: float, rng=None) -> np.ndarray:
    """One ``fillBuffer`` (noise.hpp:190-227): returns ``result`` (float32, length n)."""
    if n <= 0:
        return np.zeros(0, dtype=np.float32)
    rng = np.random if rng is None else rng
    white = np.zeros(2 * n, dtype=np.complex64)
    white[:n] = rng.normal(0.0, std, n).astype(np.float32) if std > 0 else 0.0
    col = np.zeros(2 * n, dtype=np.complex64)
    col[:n] = _kasdin_weights(n, alpha)
    W = np.fft.fft(white).astype(np.complex64)
    Cc = np.fft.fft(col).astype(np.complex64)
    P = np.zeros(2 * n, dtype=np.complex64)
    P[:n] = Cc[:n] * W[:n]
    P[0] = P[0] / 2.0
    P[n - 1] = P[n - 1] / 2.0
    # kissfft's inverse is unnormalised; numpy's ifft divides by 2n
    y = (np.fft.ifft(P) * (2 * n)).astype(np.complex64)
    return (y[:n].real / np.float32(n)).astype(np.float32)


def vector_alpha_noise(n: int, alpha: float, std: float, scale: float, axis, rng=None) -> np.ndarray:
    """All ``n`` ticks of ``VectorAlphaNoise::tickVector`` at once: float64 [3][n], scaled;
    normalised to unit length per sample when axis == all (noise.hpp:291-301)."""
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


class BufferedAlphaNoise:
    def __init__(self, bufferSize: int, alpha: float, std: float, scale: float):
        self.bufferSize = int(bufferSize)
        self.alpha = float(alpha)
        self.std = float(std)
        self.scale = float(scale)
        self._counter = 0
        self._result = np.zeros(self.bufferSize, dtype=np.float32)

    def fillBuffer(self) -> None:
        self._result = alpha_noise_buffer(self.bufferSize, self.alpha, self.std)

    def tick(self) -> float:
        if self._counter == 0:
            self.fillBuffer()
        r = float(self._result[self._counter])
        self._counter = (self._counter + 1) % self.bufferSize
        return self.scale * r


class _NullTicker:
    def tick(self) -> float:
        return 0.0


class VectorAlphaNoise:
    def __init__(self, bufferSize: int, alpha: float, std: float, scale: float, axis=3):
        from .core import CVector

        self._CVector = CVector
        self.scale = float(scale)
        ax = int(axis)
        self._c = [_NullTicker(), _NullTicker(), _NullTicker()]
        if ax == 3:
            self._c = [BufferedAlphaNoise(bufferSize, alpha, std, 1.0) for _ in range(3)]
            self.normalized = True
        elif ax in (0, 1, 2):
            self._c[ax] = BufferedAlphaNoise(bufferSize, alpha, std, 1.0)
            self.normalized = False
        else:
            raise RuntimeError("Invalid axis specified: " + str(ax))
        self.prevSample = CVector()
        self.currentSample = CVector()

    def tickVector(self):
        self.prevSample = self.currentSample
        v = self._CVector(self._c[0].tick(), self._c[1].tick(), self._c[2].tick())
        if self.normalized:
            v.normalize()
        self.currentSample = v
        return v * self.scale

    def tick(self) -> float:
        return self._c[0].tick() * self.scale

    def getPrevSample(self):
        return self.prevSample

    def getScale(self) -> float:
        return self.scale
