"""1/f^alpha noise generators: ``BufferedAlphaNoise`` / ``VectorAlphaNoise``
(core/noise.hpp:138-308, python names cmtj.cpp:449-462).

The reference fills one buffer of ``bufferSize`` float32 samples per component at the first tick
(Kasdin's filter through a 2n-point kissfft).  Here the buffer comes from the same device generator the
junction kernels draw from (``cmtjb_alpha_noise_host`` -> cmtj_b200/csrc/onef.cu: Philox white noise, cuFFT
in float32, the reference's one-sided-spectrum product); these classes only keep the reference's tick
bookkeeping.  Junction runs do not go through them: a layer with ``setAlphaNoise`` gets its buffers generated
inside ``runSimulation`` / ``JunctionBatch.run``, one series per (instance, layer, axis).  RNG streams differ
from the reference's ``mt19937(random_device)``: parity is distributional.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def alpha_noise(n: int, alpha: float, std: float, seed: int, n_series: int = 1, series0: int = 0, device: int = 0) -> np.ndarray:
    """``n_series`` independent ``BufferedAlphaNoise::fillBuffer`` results of ``n`` samples each, generated on the GPU:
    float32 [n_series, n], unscaled."""
    out = np.zeros((int(n_series), int(n)), dtype=np.float32)
    if n <= 0 or n_series <= 0:
        return out
    lib = _abi.cuda_backend().lib
    f = lib.cmtjb_alpha_noise_host
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_double, C.c_int64, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
    _abi.cuda_backend().check(f(float(alpha), float(std), int(n), int(seed), int(series0), int(n_series), out.ctypes.data, int(device)))
    return out


class BufferedAlphaNoise:
    def __init__(self, bufferSize: int, alpha: float, std: float, scale: float):
        self.bufferSize = int(bufferSize)
        self.alpha = float(alpha)
        self.std = float(std)
        self.scale = float(scale)
        self._counter = 0
        self._result = np.zeros(self.bufferSize, dtype=np.float32)

    def fillBuffer(self) -> None:
        from .core import _next_seed

        self._result = alpha_noise(self.bufferSize, self.alpha, self.std, _next_seed())[0]

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
