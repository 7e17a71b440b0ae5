"""Resistance models and spin-diode mixing of ``cmtj.utils.resistance`` (resistance.py:12-149),
evaluated on the GPU through the C-ABI (``cmtjb_strip_resistance_device``,
``cmtjb_compute_sd_device``, ``cmtjb_compute_gmr_device``).  Arguments may be numpy arrays (copied
to the device and back, same shapes as the reference) or CUDA ``torch`` tensors (results stay on
the device).  There is no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from .. import _abi
from .filters import butter_lowpass_coefficients


def _torch():
    import torch

    return torch


def _lib():
    return _abi.cuda_backend().lib


def _to_device(x, device: int):
    torch = _torch()
    if isinstance(x, torch.Tensor):
        return x.to(dtype=torch.float64), True
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(torch.device("cuda", device)), False


def _back(t, was_tensor: bool):
    return t if was_tensor else t.cpu().numpy()


def _res_rows(Rx0, Ry0, AMR, AHE, SMR, l, w) -> np.ndarray:
    n = len(Rx0)
    rows = np.zeros((n, _abi.RES_NFIELDS), dtype=np.float64)
    for i in range(n):
        rows[i] = (Rx0[i], Ry0[i], AMR[i], AHE[i], SMR[i], w[i] / l[i])
    return rows


def strip_resistance_device(res_rows: np.ndarray, m_dev, parallel: bool, weights=None, device: int = 0):
    """m_dev: CUDA float64 tensor [L*3, ...]; returns (Rx, Ry) tensors of the trailing shape."""
    torch = _torch()
    L = res_rows.shape[0]
    if m_dev.shape[0] != 3 * L:
        raise ValueError(f"m must have {3 * L} leading rows (layers x components), got {tuple(m_dev.shape)}")
    m_dev = m_dev.contiguous()
    n = int(m_dev[0].numel())
    Rx = torch.empty(m_dev.shape[1:], dtype=torch.float64, device=m_dev.device)
    Ry = torch.empty_like(Rx)
    rows = np.ascontiguousarray(res_rows, dtype=np.float64)
    wts = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    f = _lib().cmtjb_strip_resistance_device
    f.restype = C.c_int
    f.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p]
    rc = f(L, rows.ctypes.data, None if wts is None else wts.ctypes.data, 1 if parallel else 0, m_dev.data_ptr(), n,
           Rx.data_ptr(), Ry.data_ptr(), m_dev.device.index or 0, C.c_void_p(torch.cuda.current_stream(m_dev.device).cuda_stream))
    _abi.cuda_backend().check(rc)
    return Rx, Ry


def _layers_m(m, device: int):
    """[L, 3] or [L, 3, T] -> device tensor [L*3, T'] and a flag for the 2-D case."""
    t, was = _to_device(m, device)
    if t.dim() == 2:
        return t.reshape(t.shape[0] * 3, 1), was, True
    if t.dim() != 3:
        raise ValueError("m must have shape [number_of_layers, 3] or [number_of_layers, 3, T]")
    return t.reshape(t.shape[0] * 3, -1), was, False


def _combined(Rx0, Ry0, AMR, AHE, SMR, m, l, w, parallel: bool, device: int):
    md, was, flat = _layers_m(m, device)
    Rx, Ry = strip_resistance_device(_res_rows(Rx0, Ry0, AMR, AHE, SMR, l, w), md, parallel, device=device)
    if flat:
        Rx, Ry = Rx.reshape(()), Ry.reshape(())
    return _back(Rx, was), _back(Ry, was)


def calculate_resistance_series(Rx0, Ry0, AMR, AHE, SMR, m, l, w, device: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """resistance.py:124-149: Rx = sum_l Sx_l, Ry = sum_l Sy_l."""
    return _combined(Rx0, Ry0, AMR, AHE, SMR, m, l, w, False, device)


def calculate_resistance_parallel(Rx0, Ry0, AMR, AHE, SMR, m, l, w, device: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """resistance.py:96-121: Rx = 1 / sum_l 1/Sx_l (same for Ry)."""
    return _combined(Rx0, Ry0, AMR, AHE, SMR, m, l, w, True, device)


def compute_resistance(Rx0, Ry0, AMR, AHE, SMR, m, l, w, device: int = 0):
    """resistance.py:25-66: the per-layer Sx_l, Sy_l ([L] or [L, T])."""
    md, was, flat = _layers_m(m, device)
    L = len(Rx0)
    rows = _res_rows(Rx0, Ry0, AMR, AHE, SMR, l, w)
    torch = _torch()
    sx, sy = [], []
    for i in range(L):
        a, b = strip_resistance_device(rows[i : i + 1], md[3 * i : 3 * i + 3], False, device=device)
        sx.append(a)
        sy.append(b)
    Sx, Sy = torch.stack(sx), torch.stack(sy)
    if flat:
        Sx, Sy = Sx.reshape(L), Sy.reshape(L)
    return _back(Sx, was), _back(Sy, was)


def compute_gmr(Rp: float, Rap: float, m1, m2, device: int = 0):
    """resistance.py:68-74: Rp + 0.5 (Rap - Rp)(1 - m1.m2); m1, m2 of shape [3] or [3, T]."""
    torch = _torch()
    a, was = _to_device(m1, device)
    b, _ = _to_device(m2, device)
    a2, b2 = a.reshape(3, -1).contiguous(), b.reshape(3, -1).contiguous()
    out = torch.empty(a2.shape[1], dtype=torch.float64, device=a2.device)
    f = _lib().cmtjb_compute_gmr_device
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    rc = f(float(Rp), float(Rap), a2.data_ptr(), b2.data_ptr(), int(a2.shape[1]), out.data_ptr(), a2.device.index or 0,
           C.c_void_p(torch.cuda.current_stream(a2.device).cuda_stream))
    _abi.cuda_backend().check(rc)
    out = out.reshape(a.shape[1:]) if a.dim() > 1 else out.reshape(())
    return _back(out, was)


def calculate_magnetoresistance(Rp: float, Rap: float, m, device: int = 0):
    """resistance.py:77-89 -- NB the reference's formula here is Rp + 0.5 (Rap - Rp) * (m0.m1)
    (no "1 -"), i.e. compute_gmr evaluated with swapped roles: R = (Rp + Rap)/2 + ... is NOT what it
    returns.  Reproduced as written: Rp + 0.5 (Rap - Rp) sum(m[0] * m[1])."""
    arr = m if hasattr(m, "shape") else np.asarray(m)
    if arr.shape[0] != 2:
        raise ValueError("The magnetoresistance can only be computed for 2 layers" f". Current shape {tuple(arr.shape)}")
    g = compute_gmr(Rp, Rap, arr[0], arr[1], device=device)  # Rp + h (1 - c)
    half = 0.5 * (Rap - Rp)
    return Rp + half - (g - Rp)  # Rp + h c


def compute_sd_device(R_dev, I_dev, integration_step: float):
    """R_dev: CUDA tensor [n, B]; I_dev: [n] (shared) or [n, B]; returns CUDA tensor [B]."""
    torch = _torch()
    R_dev = R_dev.contiguous()
    I_dev = I_dev.contiguous()
    n, B = int(R_dev.shape[0]), int(R_dev.shape[1])
    per_series = I_dev.dim() == 2
    if int(I_dev.shape[0]) != n or (per_series and int(I_dev.shape[1]) != B):
        raise ValueError("dynamic_i must have the time length of dynamic_r")
    b, a = butter_lowpass_coefficients(10e6, 1.0 / integration_step, order=3)
    out = torch.empty(B, dtype=torch.float64, device=R_dev.device)
    f = _lib().cmtjb_compute_sd_device
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p]
    rc = f(R_dev.data_ptr(), I_dev.data_ptr(), 1 if per_series else 0, n, B, b.ctypes.data, a.ctypes.data, out.data_ptr(),
           R_dev.device.index or 0, C.c_void_p(torch.cuda.current_stream(R_dev.device).cuda_stream))
    _abi.cuda_backend().check(rc)
    return out


def compute_sd(dynamic_r, dynamic_i, integration_step: float, device: int = 0):
    """resistance.py:12-22: mean(butter_lowpass(-I*R, 10 MHz, order 3)).  1-D inputs give a float
    like the reference; 2-D inputs [n, B] give B values (one series per column)."""
    R, was = _to_device(dynamic_r, device)
    I, _ = _to_device(dynamic_i, device)
    one = R.dim() == 1
    out = compute_sd_device(R.reshape(R.shape[0], -1), I if I.dim() == 1 else I.reshape(I.shape[0], -1), integration_step)
    if one:
        return out[0] if was else float(out[0].item())
    return _back(out, was)
