"""Field-sweep geometry (host-side inputs of the procedures): the reference's
``cmtj.utils.linear.FieldScan`` (cmtj/utils/linear.py:6-128).  Angles in degrees, physics
convention (theta from +z, phi in the xy plane)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ..core import CVector


def _unit(theta, phi):
    t, p = np.deg2rad(theta), np.deg2rad(phi)
    st, ct, sp, cp = np.sin(t), np.cos(t), np.sin(p), np.cos(p)
    return st * cp, st * sp, ct


class FieldScan:
    @staticmethod
    def angle2vector(theta, phi, amplitude=1) -> CVector:
        ux, uy, uz = _unit(theta, phi)
        return CVector(ux * amplitude, uy * amplitude, uz * amplitude)

    @staticmethod
    def vector2angle(x, y, z) -> Tuple[float, float, float]:
        """(theta, phi, r) of a cartesian vector, degrees."""
        rho2 = x**2 + y**2
        return np.rad2deg(np.arctan2(np.sqrt(rho2), z)), np.rad2deg(np.arctan2(y, x)), np.sqrt(rho2 + z**2)

    @staticmethod
    def cvector2angle(vector: CVector) -> Tuple[float, float, float]:
        return FieldScan.vector2angle(vector.x, vector.y, vector.z)

    @staticmethod
    def amplitude_scan(start: float, stop: float, steps: int, theta: float, phi: float, back: bool = False):
        """Linear magnitude sweep along a fixed direction: (amplitudes, [steps, 3] field vectors);
        ``back`` appends the return sweep (the turning point is not repeated)."""
        H = np.linspace(start, stop, num=steps, endpoint=True)
        ux, uy, uz = _unit(theta, phi)
        vec = np.stack((ux * H, uy * H, uz * H), axis=1)
        if back:
            return np.concatenate((H[:-1], H[::-1])), np.concatenate((vec[:-1], vec[::-1]))
        return H, vec

    @staticmethod
    def theta_scan(start: float, stop: float, steps: int, amplitude: float, phi: float):
        th = np.linspace(start, stop, num=steps, endpoint=True)
        ux, uy, uz = _unit(th, phi)
        return th, np.stack((ux * amplitude, uy * amplitude, uz * amplitude), axis=1)

    @staticmethod
    def phi_scan(start: float, stop: float, steps: int, amplitude: float, theta: float):
        ph = np.linspace(start, stop, num=steps, endpoint=True)
        ux, uy, uz = _unit(theta, ph)
        return ph, np.stack((ux * amplitude, uy * amplitude, uz * amplitude * np.ones_like(uy)), axis=1)
