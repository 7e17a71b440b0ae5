"""Batched, device-resident versions of the reference's sweep procedures
(cmtj/utils/procedures.py: ``PIMM_procedure`` :42-204, ``VSD_procedure`` :207-331).

The reference loops ``runSimulation`` over field (and frequency) points, logs every layer at every
step, pulls the log through pybind11 and reduces it with numpy/scipy.  Here the sweep points are the
instance axis of ONE fused launch and the reductions happen on the GPU:

* PIMM: the kernel writes only the projected signal ``sum_l w_l m_l,c`` per step
  (``cmtjb_post_t.proj``), the spectrum is a batched FFT over that buffer, and the static
  resistances come from a ``take_last_n``-sample log evaluated by ``strip_resistance_kernel``.
* VSD: the mixing ``-I(t) R(t)``, the 10 MHz Butterworth low-pass and the mean run inside the time
  loop (``cmtjb_post_t.sd_*``); nothing is logged at all.

Semantics.  The reference carries the magnetisation from one sweep point to the next (hysteresis,
cmtj/utils/procedures.py:120-141): AHE loops and up/down field scans depend on it.  That is the
default here too (``carry_state=True``): points run one after another, each from the state the previous
one left, consuming numpy's global RNG stream exactly as the reference does -- a chain of dependent runs,
so no faster than the chain allows.  ``carry_state=False`` -- or the explicitly named ``PIMM_batch`` /
``VSD_batch`` -- starts every point from the junction's current magnetisation (plus the same ``disturbance``
draws) and integrates the whole sweep as ONE batch; use it when the points are independent (spectra and
spin-diode maps away from hysteresis), it is what the GPU is for.
Everything else -- arguments, validation, return shapes -- follows the reference.
"""
from __future__ import annotations

import functools
import os as _os
import warnings

import math
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from .. import _abi
from ..batch import JunctionBatch, PostSpec
from ..core import AxialDriver, Axis, CVector, Junction, NullDriver, ScalarDriver
from . import resistance as _res
from .filters import butter_lowpass_coefficients
from .resistance import calculate_resistance_series


@dataclass
class ResistanceParameters:
    """procedures.py:15-26"""

    Rxx0: float = 0
    Rxy0: float = 0
    Rahe: float = 0
    Rsmr: float = 0
    Ramr: float = 0
    w: float = 0  # width
    l: float = 0  # length


def _torch():
    import torch

    return torch


class _Phases:
    """CMTJB_TRACE_PROC=1: wall time of each phase of a procedure call (device synchronised at every mark)."""

    def __init__(self, name):
        import os
        import time

        self.on = bool(os.environ.get("CMTJB_TRACE_PROC"))
        self.name, self.t, self.time = name, None, time
        self.mark(None)

    def mark(self, what):
        if not self.on:
            return
        _torch().cuda.synchronize()
        now = self.time.perf_counter()
        if what is not None:
            print(f"[cmtjb] {self.name}: {what} {1e3 * (now - self.t):.1f} ms", flush=True)
        self.t = now


def _spectrum_bins(n: int, int_step: float, max_frequency: float):
    """The bins compute_spectrum_strip keeps (procedures.py:29-40): first n//2 of fftfreq(n, dt),
    then those <= max_frequency.  Returns (findx [K,1], freqs[findx] [K,1])."""
    freqs = np.fft.fftfreq(n, int_step)[: n // 2]
    findx = np.argwhere(freqs <= max_frequency)
    return findx, freqs[findx]


def spectrum_device(signal, int_step: float, max_frequency: float):
    """|FFT| of B real series stored time-major [n, B] on the device (rows may be a view with a pitch), restricted
    like compute_spectrum_strip.  Returns (CUDA tensor [B, K], freqs [K, 1]).  ``cmtjb_rfft_abs_device``: cuFFT in
    double precision over column slices of the buffer as it lies (no transposed copy), only the K bins up to
    ``max_frequency`` are kept, so a whole-sweep projection log (tens of GB) never needs a second buffer of its
    own size."""
    import ctypes as C

    torch = _torch()
    n, B = int(signal.shape[0]), int(signal.shape[1])
    findx, freqs = _spectrum_bins(n, int_step, max_frequency)
    K = int(findx.size)
    assert signal.dtype == torch.float64 and signal.stride(1) == 1 and (n == 1 or signal.stride(0) >= B)
    if K and not np.array_equal(findx.ravel(), np.arange(K)):  # fftfreq is increasing on [0, n/2): always a prefix
        raise RuntimeError("unexpected frequency selection")
    out = torch.empty((B, K), dtype=torch.float64, device=signal.device)
    if K == 0 or B == 0:
        return out, freqs
    be = _abi.cuda_backend()
    dev = signal.device.index or 0
    q = be.lib.cmtjb_rfft_abs_workspace
    q.restype = C.c_int
    q.argtypes = [C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    per_series = C.c_int64(0)
    be.check(q(n, K, dev, C.byref(per_series)))
    ws, ws_ptr, ws_bytes = None, None, 0
    import os

    if per_series.value > 0 and not os.environ.get("CMTJB_SPECTRUM_NO_WS"):
        # moments of a slab of series: from torch's caching allocator, so that a sweep's repeated calls do not allocate;
        # a quarter of the free memory, at most 2 GiB, at least 1024 series where that fits
        free_b = torch.cuda.mem_get_info(signal.device)[0]
        cap = max(min(free_b // 4, 2 << 30), min(free_b // 2, 1024 * per_series.value))
        if os.environ.get("CMTJB_SPECTRUM_SLAB_MB"):
            cap = int(os.environ["CMTJB_SPECTRUM_SLAB_MB"]) << 20
        slab = max(1, min(B, cap // per_series.value))
        ws = torch.empty(slab * per_series.value, dtype=torch.uint8, device=signal.device)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    f = be.lib.cmtjb_rfft_abs_device_ws
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    ld = int(signal.stride(0)) if n > 1 else B
    stream = torch.cuda.current_stream(signal.device).cuda_stream
    be.check(f(signal.data_ptr(), n, B, ld, K, out.data_ptr(), dev, C.c_void_p(stream), ws_ptr, ws_bytes))
    return out, freqs


class _SpectrumStream:
    """|FFT| of B series that arrive in time chunks (``cmtjb_spectrum_stream_*``): a sweep's full-rate projection log
    (1e5 points x 2e5 steps: 160 GB) is never resident -- each chunk is reduced to its Chebyshev block moments right after
    the time loop has written it, and the moments (a third of the log here) are transformed at the end."""

    def __init__(self, n: int, B: int, int_step: float, max_frequency: float, device: int):
        import ctypes as C

        torch = _torch()
        self.n, self.B, self.device = int(n), int(B), int(device)
        findx, self.freqs = _spectrum_bins(self.n, int_step, max_frequency)
        self.K = int(findx.size)
        self.block_len = 0
        if self.K == 0 or self.B == 0 or self.n <= 0:
            return
        self.be = _abi.cuda_backend()
        f = self.be.lib.cmtjb_spectrum_stream_plan
        f.restype = C.c_int
        f.argtypes = [C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        bl, per = C.c_int64(0), C.c_int64(0)
        self.be.check(f(self.n, self.K, self.device, C.byref(bl), C.byref(per)))
        self.block_len, self.bytes_per_series = int(bl.value), int(per.value)
        self.n_blocks = self.n // self.block_len if self.block_len else 0
        self.moments = None
        self._C = C

    def allocate(self):
        torch = _torch()
        self.moments = torch.empty(self.B * self.bytes_per_series, dtype=torch.uint8, device=torch.device("cuda", self.device))

    def add(self, chunk, q0: int, nq: int, background: bool = False) -> None:
        """chunk: device tensor [nq * block_len, B] (row 0 = first sample of block q0), reduced on the current stream;
        ``background``: the low-footprint kernel that fits beside a resident time-loop kernel (see the C header)"""
        C, torch = self._C, _torch()
        assert chunk.dtype == torch.float64 and chunk.stride(1) == 1 and chunk.shape[1] == self.B and chunk.shape[0] >= nq * self.block_len
        f = self.be.lib.cmtjb_spectrum_stream_moments
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int,
                      C.c_void_p]
        stream = torch.cuda.current_stream(chunk.device).cuda_stream
        self.be.check(f(chunk.data_ptr(), int(chunk.stride(0)), self.B, self.n, self.K, int(q0), int(nq), self.moments.data_ptr(),
                        1 if background else 0, self.device, C.c_void_p(stream)))

    def _combine(self, b_lo: int, b_hi: int, out) -> None:
        """series [b_lo, b_hi): moments -> |X[k]|, k < K, into out[b_lo:b_hi] (device), on the current stream"""
        C, torch = self._C, _torch()
        f = self.be.lib.cmtjb_spectrum_stream_finish
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        stream = torch.cuda.current_stream(out.device).cuda_stream
        self.be.check(f(self.moments.data_ptr() + b_lo * self.bytes_per_series, self.n, b_hi - b_lo, self.K,
                        out.data_ptr() + b_lo * self.K * 8, self.device, C.c_void_p(stream)))

    def finish(self):
        """[B, K] spectra on the device"""
        torch = _torch()
        out = torch.empty((self.B, self.K), dtype=torch.float64, device=torch.device("cuda", self.device))
        self._combine(0, self.B, out)
        return out

    def finish_to_host(self, n_slabs: int = 4) -> np.ndarray:
        """[B, K] spectra in page-locked host memory: the series are transformed in slabs and the copy of one slab to the
        host runs (second stream) while the next is transformed"""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        out = torch.empty((self.B, self.K), dtype=torch.float64, device=dev)
        host = torch.empty((self.B, self.K), dtype=torch.float64, pin_memory=True)
        cur, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
        step = max(1, -(-self.B // max(1, n_slabs)))
        for b_lo in range(0, self.B, step):
            b_hi = min(self.B, b_lo + step)
            self._combine(b_lo, b_hi, out)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(done)
                host[b_lo:b_hi].copy_(out[b_lo:b_hi], non_blocking=True)
        side.synchronize()
        out.record_stream(side)
        return host.numpy()


class _Rows(list):
    """The per-point lists the reference's procedures return (one ``output["Rx"].append(...)`` per field point,
    procedures.py:183-193) for a whole batch: a ``list`` of row views of ONE [n_points, ...] array, materialised when it
    is first used as a list -- a 1e5-point sweep does not pay for 4e5 Python objects it may never look at
    (``np.asarray(rows)`` hands back the array itself)."""

    def __init__(self, array: np.ndarray):
        super().__init__()
        self._array = array  # None once the rows exist as list items

    def _fill(self) -> None:
        if self._array is not None:
            a, self._array = self._array, None
            list.extend(self, a)

    def __array__(self, dtype=None, copy=None):
        a = self._array if self._array is not None else np.array(list(self))
        return a if dtype is None else a.astype(dtype, copy=False)

    def __reduce__(self):
        self._fill()
        return (list, (list(self),))


def _filled(name):
    inner = getattr(list, name)

    def method(self, *a, **k):
        self._fill()
        return inner(self, *a, **k)

    method.__name__ = name
    return method


for _n in ("__len__", "__getitem__", "__iter__", "__contains__", "__reversed__", "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__",
           "__add__", "__mul__", "__rmul__", "__repr__", "__setitem__", "__delitem__", "__iadd__", "__imul__", "index", "count", "copy",
           "append", "extend", "insert", "pop", "remove", "clear", "sort", "reverse"):
    setattr(_Rows, _n, _filled(_n))


def _to_host(t):
    """device tensor -> numpy through page-locked memory (torch caches the pinned blocks between calls)"""
    torch = _torch()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def compute_spectrum_strip(input_m, int_step: float, max_frequency: float):
    """procedures.py:29-40 for one host series (runs on the device)."""
    torch = _torch()
    x = torch.from_numpy(np.ascontiguousarray(np.asarray(input_m, dtype=np.float64))).cuda().reshape(-1, 1)
    yf, freqs = spectrum_device(x, int_step, max_frequency)
    return yf.cpu().numpy().reshape(-1, 1), freqs


def _steps(total: float, dt: float) -> int:
    """(unsigned)(totalTime / timeStep) of Junction::runSimulation (junction.hpp:1956)."""
    return int(total / dt)


def _time_for_steps(n: int, dt: float) -> float:
    """A totalTime whose (unsigned)(totalTime/dt) is exactly n (used to stop a run after n steps)."""
    t = n * dt
    for _ in range(8):
        k = int(t / dt)
        if k == n:
            return t
        t = math.nextafter(t, math.inf if k < n else -math.inf)
    raise RuntimeError("could not represent the requested step count")


def _res_rows(resistance_params: List[ResistanceParameters]) -> np.ndarray:
    return _res._res_rows(
        [r.Rxx0 for r in resistance_params],
        [r.Rxy0 for r in resistance_params],
        [r.Ramr for r in resistance_params],
        [r.Rahe for r in resistance_params],
        [r.Rsmr for r in resistance_params],
        [r.l for r in resistance_params],
        [r.w for r in resistance_params],
    )


def _resistance_kind(fn: Callable) -> Optional[bool]:
    """True = parallel, False = series, None = an unknown callable (evaluated on host arrays)."""
    name = getattr(fn, "__name__", "")
    if name == "calculate_resistance_parallel":
        return True
    if name == "calculate_resistance_series":
        return False
    return None


def _oersted_driver(direction: Axis, driver: ScalarDriver) -> Tuple[str, int, AxialDriver]:
    if direction == Axis.zaxis:
        return "z", 2, AxialDriver(NullDriver(), NullDriver(), driver)
    if direction == Axis.yaxis:
        return "y", 1, AxialDriver(NullDriver(), driver, NullDriver())
    return "x", 0, AxialDriver(driver, NullDriver(), NullDriver())


def _disturbed(junction: Junction, layer_ids: List[str], n: int, disturbance: float) -> Dict[str, np.ndarray]:
    """Initial magnetisation of n instances: the junction's current m plus the reference's
    np.random.normal(0, disturbance) per component (same draw order: point, layer, x/y/z),
    normalised (procedures.py:131-140).  With disturbance == 0 the reference never touches the
    magnetisation, so a freshly constructed layer starts its first step from the un-normalised
    constructor input; callers store the result with ``set_magnetisation(..., normalise=bool(disturbance))``."""
    base = np.array([junction.getLayerMagnetisation(lid).tolist() for lid in layer_ids])  # [L,3]
    m = np.repeat(base[None], n, axis=0)
    if disturbance:
        m = m + np.random.normal(0, disturbance, size=m.shape)
        m = m / np.sqrt((m * m).sum(axis=2, keepdims=True))
    return {lid: m[:, i, :] for i, lid in enumerate(layer_ids)}


def _instance_chunks(n: int, bytes_per_instance: float, budget: float) -> List[Tuple[int, int]]:
    per = max(1, int(budget // max(bytes_per_instance, 1.0)))
    return [(lo, min(n, lo + per)) for lo in range(0, n, per)]


def _warn_sequential(name: str, batch_name: str, n_points: int) -> None:
    if n_points >= 64:
        warnings.warn(
            f"{name}: integrating {n_points} sweep points one after another to carry the magnetisation from point to "
            f"point like the reference does; if the points are independent, carry_state=False (or {batch_name}) runs "
            "them as one batch",
            stacklevel=3,
        )


# ---------------------------------------------------------------------------------------------------
# PIMM
# ---------------------------------------------------------------------------------------------------
def _pimm_batch(junction, Hc, m0, int_step, res_rows, kind, resistance_fn, resistance_params, comp, weights, simulation_duration,
                wait_time, max_frequency, take_last_n, full_output, static_only, device, normalise):
    """One batch of field points (all start from m0).  Returns a dict of host arrays."""
    torch = _torch()
    ph = _Phases("PIMM batch")
    B = Hc.shape[0]
    L = junction.layerNo
    layer_ids = junction.getLayerIds()
    dev = torch.device("cuda", device)
    n_total = _steps(simulation_duration, int_step)
    n_last = min(int(take_last_n), n_total)
    batch = JunctionBatch(junction, B, device=device)
    batch.set_external_field("all", Hc)
    for lid in layer_ids:
        batch.set_magnetisation(lid, m0[lid], normalise=normalise)
    out: Dict[str, Any] = {}
    n_proj = 0 if static_only else (1 + L if full_output else 1)
    if n_proj > _abi.MAX_PROJ:
        raise NotImplementedError("full_output supports junctions of up to 3 layers")
    proj = np.zeros((max(n_proj, 1), 3 * L))
    for l in range(L):
        proj[0, 3 * l + comp] = weights[l]
        if full_output and not static_only:
            proj[1 + l, 3 * l + comp] = weights[l]
    # where the transformed series starts (procedures.py:148: the samples with time >= wait_time)
    n0 = min(n_total, max(0, int(math.ceil(wait_time / int_step)) - 2))
    while n0 < n_total and n0 * float(int_step) < wait_time:
        n0 += 1
    # streamed spectrum: the projected signal goes through a chunk buffer and is reduced chunk by chunk
    stream = None
    if n_proj == 1 and n_total - n0 > 0 and not _os.environ.get("CMTJB_PIMM_RESIDENT_LOG"):
        stream = _SpectrumStream(n_total - n0, B, int_step, max_frequency, device)
        chunk_bytes = float(_os.environ.get("CMTJB_PIMM_CHUNK_MB", 10240)) * 2.0**20
        chunks: List[Tuple[int, int]] = []
        if stream.block_len:
            cq = max(2, int(chunk_bytes // (8.0 * stream.block_len * B)) // 2 * 2)  # blocks per chunk, even
            if cq >= 64:
                cq = cq // 64 * 64  # the moments kernel hands 8 x 8 blocks to a thread block: whole ones only
            chunks = [(q0, min(cq, stream.n_blocks - q0)) for q0 in range(0, stream.n_blocks, cq)]
            if len(chunks) > 1 and chunks[-1][1] * stream.block_len < n_last:  # the last-n log must not straddle chunks
                (qa, na), (_, nb) = chunks[-2], chunks[-1]
                chunks[-2:] = [(qa, na + nb)]
            if chunks[-1][1] * stream.block_len < n_last:
                chunks = []
        if not chunks:
            stream = None  # resident log instead
    if stream is not None:
        stream.allocate()
        # Default: one stream, one chunk buffer -- the time loop of chunk c, then its reduction (4.3 ms per 10 GB chunk).
        # CMTJB_PIMM_OVERLAP=1: two buffers, the reduction of chunk c on a second stream (low-footprint background
        # kernel) while chunk c + 1 integrates.  Measured on B200 at 1e5 x 2e5: 465 ms of device time against 470 ms for
        # the default -- the reduction's FP64 instructions compete with the time loop for the same pipe (each chunk's
        # time loop slows from 24 to 27 ms), so hiding it buys nothing; kept as a measured experiment.
        overlap = len(chunks) > 1 and bool(_os.environ.get("CMTJB_PIMM_OVERLAP"))
        ring = torch.empty((2 if overlap else 1, 1, max(nq for _, nq in chunks) * stream.block_len, B), dtype=torch.float64, device=dev)
        proj_log = ring[0]
    else:
        proj_log = torch.empty((n_proj, n_total, B), dtype=torch.float64, device=dev) if n_proj else None

    def spec(sample0, buf=None):
        return PostSpec(proj=proj[:n_proj], proj_log=proj_log if buf is None else buf, proj_sample0=sample0) if n_proj else None

    ph.mark("set-up (batch, inputs, projection log buffer)")
    if stream is not None:
        n_a = n_total - n_last
        cur = torch.cuda.current_stream(dev)
        if overlap:
            # the time loop on a high-priority stream (its blocks are placed first and keep their share of every SM), the
            # reductions on a second one: they fill the registers the time-loop kernel leaves free
            s_int, s_mom = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev)
            s_int.wait_stream(cur)
            s_mom.wait_stream(cur)
        else:
            s_int = s_mom = cur
        slot_free = [None] * ring.shape[0]
        marks = []
        if ph.on:
            ev_start = torch.cuda.Event(enable_timing=True)
            ev_start.record(s_int)
            h_start = ph.time.perf_counter()
        with torch.cuda.stream(s_int):
            if n0 > 0:
                batch.run_device(_time_for_steps(n0, int_step), int_step, int_step, log=False)
        L2 = stream.block_len
        for c, (q0, nq) in enumerate(chunks):
            t0, t1 = n0 + q0 * L2, n0 + (q0 + nq) * L2
            k = c % ring.shape[0]
            buf = ring[k]
            with torch.cuda.stream(s_int):
                if slot_free[k] is not None:
                    s_int.wait_event(slot_free[k])
                if t1 <= n_a:
                    batch.run_device(_time_for_steps(t1, int_step), int_step, int_step, log=False, post=spec(0, buf), start_iteration=t0)
                else:  # the last chunk: its final n_last steps also log every layer's m (static resistances)
                    if n_a > t0:
                        batch.run_device(_time_for_steps(n_a, int_step), int_step, int_step, log=False, post=spec(0, buf), start_iteration=t0)
                    res = batch.run_device(simulation_duration, int_step, int_step, post=spec(max(n_a, t0) - t0, buf),
                                           start_iteration=max(n_a, t0))
                    log_last = res["_log"][:, res["_log"].shape[1] - n_last :, :]
                    log_last.record_stream(cur)
                done = torch.cuda.Event(enable_timing=ph.on)
                done.record(s_int)
            with torch.cuda.stream(s_mom):
                s_mom.wait_event(done)
                stream.add(buf[0], q0, nq, background=overlap and c + 1 < len(chunks))
                slot_free[k] = torch.cuda.Event(enable_timing=ph.on)
                slot_free[k].record(s_mom)
            if ph.on:
                marks.append((ph.time.perf_counter(), done, slot_free[k]))
        if overlap:
            cur.wait_stream(s_int)
            cur.wait_stream(s_mom)
        if ph.on and marks:
            torch.cuda.synchronize()
            print("[cmtjb] PIMM batch chunks: host enqueue done / time loop done / moments done, ms after the phase started "
                  f"(host wall until everything had finished: {1e3 * (ph.time.perf_counter() - h_start):.1f}):", flush=True)
            for i, (h, e_int, e_mom) in enumerate(marks):
                print(f"[cmtjb]   chunk {i:3d}: {1e3 * (h - h_start):8.1f} {ev_start.elapsed_time(e_int):8.1f} {ev_start.elapsed_time(e_mom):8.1f}",
                      flush=True)
    elif full_output and not static_only:
        res = batch.run_device(simulation_duration, int_step, int_step, post=spec(0))
        log_all = res["_log"]
        log_last = log_all[:, n_total - n_last :, :]
        out["m_traj_dev"] = log_all[: 3 * L]
    else:
        n_a = n_total - n_last
        if n_a > 0:
            batch.run_device(_time_for_steps(n_a, int_step), int_step, int_step, log=False, post=spec(0))
        res = batch.run_device(simulation_duration, int_step, int_step, post=spec(n_a), start_iteration=n_a)
        log_last = res["_log"]
    ph.mark(f"integration of {B} instances x {n_total} steps")
    m_last = log_last[: 3 * L].contiguous()  # [3L, n_last, B]
    w_dev = torch.tensor(weights, dtype=torch.float64, device=dev)
    # static resistance over the last samples, from the thickness*Ms-weighted m (procedures.py:152-162)
    if kind is not None:
        Rx, Ry = _res.strip_resistance_device(res_rows, m_last, kind, weights=weights, device=device)
        out["Rx"], out["Ry"] = _to_host(Rx.t().contiguous()), _to_host(Ry.t().contiguous())  # [B, n_last]
    else:
        mh = (m_last.reshape(L, 3, n_last, B) * w_dev[:, None, None, None]).cpu().numpy()
        rx, ry = [], []
        for i in range(B):
            a, b = resistance_fn(
                [r.Rxx0 for r in resistance_params], [r.Rxy0 for r in resistance_params], [r.Ramr for r in resistance_params],
                [r.Rahe for r in resistance_params], [r.Rsmr for r in resistance_params], mh[:, :, :, i],
                l=[r.l for r in resistance_params], w=[r.w for r in resistance_params],
            )  # fmt: skip
            rx.append(a)
            ry.append(b)
        out["Rx"], out["Ry"] = np.asarray(rx), np.asarray(ry)
    final = m_last[:, -1, :].reshape(L, 3, B)
    out["m_avg"] = _to_host((final * w_dev[:, None, None]).sum(0).t().contiguous())  # [B,3]
    out["m_final"] = _to_host(final.contiguous())  # [L,3,B]
    ph.mark("static resistances + final state to the host")
    if not static_only:
        if stream is not None:
            out["spectrum"], freqs = stream.finish_to_host(), stream.freqs  # [B,K]
        else:
            yf, freqs = spectrum_device(proj_log[0, n0:, :], int_step, max_frequency)
            ph.mark("spectrum (batched FFT)")
            out["spectrum"] = _to_host(yf)  # [B,K]
        out["freqs"] = freqs
        ph.mark("spectrum to the host")
        if full_output:
            out["layer_spectra"] = [_to_host(spectrum_device(proj_log[1 + l, n0:, :], int_step, max_frequency)[0]) for l in range(L)]
            mt = out.pop("m_traj_dev").reshape(L, 3, n_total, B) * w_dev[:, None, None, None]
            out["m_traj"] = mt.permute(3, 0, 1, 2).cpu().numpy()  # [B,L,3,T]
    return out


def PIMM_procedure(
    junction: Junction,
    Hvecs: np.ndarray,
    int_step: float,
    resistance_params: List[ResistanceParameters],
    Hoe_direction: Axis = Axis.zaxis,
    Hoe_excitation: float = 50,
    Hoe_duration: int = 3,
    simulation_duration: float = 5e-9,
    wait_time: float = 0e-9,
    max_frequency: float = 80e9,
    resistance_fn: Callable = calculate_resistance_series,
    disturbance: float = 1e-3,
    take_last_n: int = 100,
    full_output: bool = False,
    disable_tqdm: bool = False,
    static_only: bool = False,
    carry_state: bool = True,
    device: int = 0,
    memory_budget: Optional[float] = None,
) -> Tuple[np.ndarray, np.ndarray, Dict[str, Any]]:
    """Pulse-induced microwave magnetometry over a field sweep (procedures.py:42-204): returns
    ``(spectrum [n_H, n_f], frequencies [n_f, 1], other_data)`` with other_data keys ``H``, ``Rx``,
    ``Ry``, ``m_avg`` (+ ``m_traj`` and per-layer spectra when ``full_output``).

    ``carry_state`` / ``device`` / ``memory_budget`` are additions: see the module docstring."""
    if wait_time > simulation_duration:
        raise ValueError("wait_time must be smaller than simulation_duration!")
    comp_name, comp, oedriver = _oersted_driver(
        Hoe_direction, ScalarDriver.getStepDriver(0, Hoe_excitation, 0, int_step * Hoe_duration)
    )
    layer_ids = junction.getLayerIds()
    if len(layer_ids) != len(resistance_params):
        raise ValueError("The number of layers in the junction must match the number of resistance parameters!")
    Hvecs = np.asarray(Hvecs, dtype=np.float64).reshape(-1, 3)
    layers = junction.layers
    normalising_factor = np.sum([layer.thickness * layer.Ms for layer in layers])
    weights = np.array([layer.thickness * layer.Ms / normalising_factor for layer in layers])
    kind = _resistance_kind(resistance_fn)
    rows = _res_rows(resistance_params)
    junction.clearLog()
    junction.setLayerOerstedFieldDriver("all", oedriver)
    n_total = _steps(simulation_duration, int_step)
    L = len(layer_ids)

    def run(Hc, m0):
        return _pimm_batch(junction, Hc, m0, int_step, rows, kind, resistance_fn, resistance_params, comp, weights,
                           simulation_duration, wait_time, max_frequency, take_last_n, full_output, static_only, device, bool(disturbance))

    parts: List[Dict[str, Any]] = []
    if carry_state:
        _warn_sequential("PIMM_procedure", "PIMM_batch", len(Hvecs))
        for H in Hvecs:
            r = run(H[None], _disturbed(junction, layer_ids, 1, disturbance))
            for i, lid in enumerate(layer_ids):
                junction.setLayerMagnetisation(lid, CVector(*r["m_final"][i, :, 0]))
            parts.append(r)
    else:
        m0 = _disturbed(junction, layer_ids, Hvecs.shape[0], disturbance)
        n_series = 0 if static_only else (1 + L if full_output else 1)
        if memory_budget is None:
            # everything that is free now, less what the spectrum slices, their cuFFT work area and the small logs take:
            # splitting a sweep into instance chunks costs throughput (fewer warps per SM sub-partition), so the whole
            # sweep stays one batch whenever its projection log fits (1e5 points x 20 ns: 160 GB)
            torch = _torch()
            free_b, _ = torch.cuda.mem_get_info(device)
            free_b += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)  # torch's own cache is reusable
            memory_budget = max(free_b - 8e9, 0.25 * free_b)
        # device bytes per instance: the projection log(s), the full-rate log when full_output, the last-n log
        # (twice: the weighted m_traj copy handed back to the host is staged on the device)
        per_inst = 8.0 * n_total * (n_series + (2.0 * (3 * L + 3) if full_output else 0.0)) + 8.0 * 200 * (3 * L + 3)
        if n_series == 1 and not _os.environ.get("CMTJB_PIMM_RESIDENT_LOG"):
            # streamed spectrum (_SpectrumStream): block moments instead of the log, plus one chunk buffer for the batch
            n0 = int(np.searchsorted(np.arange(n_total, dtype=np.float64) * float(int_step), wait_time, side="left"))
            probe = _SpectrumStream(n_total - n0, 1, int_step, max_frequency, device) if n_total > n0 else None
            if probe is not None and probe.block_len and probe.block_len >= min(int(take_last_n), n_total):
                per_inst = float(probe.bytes_per_series) + 8.0 * 200 * (3 * L + 3) + 8.0 * max(take_last_n, 100) * (3 * L + 3)
                memory_budget = max(memory_budget - float(_os.environ.get("CMTJB_PIMM_CHUNK_MB", 10240)) * 2.0**20, 0.5 * memory_budget)
        for lo, hi in _instance_chunks(Hvecs.shape[0], per_inst, memory_budget):
            parts.append(run(Hvecs[lo:hi], {k: v[lo:hi] for k, v in m0.items()}))
        if parts:
            last = parts[-1]["m_final"]
            for i, lid in enumerate(layer_ids):
                junction.setLayerMagnetisation(lid, CVector(*last[i, :, -1]))
    if len(Hvecs):
        junction.setLayerExternalFieldDriver("all", AxialDriver(*[ScalarDriver.getConstantDriver(float(h)) for h in Hvecs[-1]]))

    output: Dict[str, Any] = defaultdict(list)
    spectrum: List[np.ndarray] = []
    freqs = None
    if not parts:
        return np.squeeze(np.asarray(spectrum)), freqs, output

    def cat(key):
        return parts[0][key] if len(parts) == 1 else np.concatenate([r[key] for r in parts], axis=0)

    # the reference appends one entry per field point (procedures.py:183-193); same lists, assembled
    # from whole-batch arrays (row views) instead of 1e5 Python-level appends
    n = sum(r["Rx"].shape[0] for r in parts)
    output["H"] = _Rows(Hvecs[:n])
    output["Rx"] = _Rows(cat("Rx"))
    output["Ry"] = _Rows(cat("Ry"))
    output["m_avg"] = _Rows(cat("m_avg"))
    spectrum_arr = np.squeeze(np.asarray(spectrum))
    if not static_only:
        freqs = parts[-1]["freqs"]
        spectrum_arr = np.squeeze(cat("spectrum")[:, :, None])
        if full_output:
            output["m_traj"] = _Rows(cat("m_traj"))
            for li, lid in enumerate(layer_ids):
                output[lid] = np.concatenate([r["layer_spectra"][li] for r in parts], axis=0)[:, :, None].squeeze()
    return spectrum_arr, freqs, output


# ---------------------------------------------------------------------------------------------------
# VSD
# ---------------------------------------------------------------------------------------------------
def _vsd_batch(junction, Hc, fc, m0, int_step, sd_mode, rows, comp, simulation_duration, device, normalise):
    torch = _torch()
    B = Hc.shape[0]
    dev = torch.device("cuda", device)
    layer_ids = junction.getLayerIds()
    L = len(layer_ids)
    n_total = _steps(simulation_duration, int_step)
    batch = JunctionBatch(junction, B, device=device)
    batch.set_external_field("all", Hc)
    batch.set_override("all", _abi.SLOT_HOE + comp, "frequency", fc)
    for lid in layer_ids:
        batch.set_magnetisation(lid, m0[lid], normalise=normalise)
    b, a = butter_lowpass_coefficients(10e6, 1.0 / int_step, order=3)  # compute_sd, resistance.py:19-21
    omega = torch.from_numpy(np.ascontiguousarray(2 * math.pi * fc)).to(dev)  # procedures.py:327
    state = torch.zeros((4, B), dtype=torch.float64, device=dev)
    post = PostSpec(sd_mode=sd_mode, res=rows if rows is not None else np.zeros((L, _abi.RES_NFIELDS)), sd_b=b, sd_a=a / a[0],
                    sd_omega=omega, sd_state=state)
    batch.run_device(simulation_duration, int_step, int_step, log=False, post=post)
    vmix = (state[3] / float(n_total)).cpu().numpy()
    return vmix, batch.get_magnetisation


def VSD_procedure(
    junction: Junction,
    Hvecs: np.ndarray,
    frequencies: np.ndarray,
    int_step: float,
    resistance_params: Optional[List[ResistanceParameters]] = None,
    Hoe_direction: Axis = Axis.yaxis,
    Hoe_excitation: float = 50,
    simulation_duration: float = 30e-9,
    disturbance: float = 1e-3,
    Rtype: str = "Rz",
    resistance_fn: Callable = calculate_resistance_series,
    disable_tqdm: bool = False,
    carry_state: bool = True,
    device: int = 0,
    max_batch: int = 1 << 22,
) -> np.ndarray:
    """Voltage spin-diode map over (field, frequency) (procedures.py:207-331): every (H, f) pair is one
    instance; the Oersted sine excitation, the mixing with sin(2 pi f t), the low-pass and the mean
    are evaluated inside the time loop.  Returns ``spectrum [len(Hvecs), len(frequencies)]``."""
    if resistance_params is None:
        resistance_params = []
    layer_ids = junction.getLayerIds()
    if Rtype == "Rz" and len(layer_ids) > 2:
        raise ValueError("Rz can only be used for 2 layer junctions. Use Rx or Ry instead.")
    elif len(resistance_params) != len(layer_ids):  # also for Rz, as in the reference (procedures.py:247-250)
        raise ValueError("The number of layers in the junction must match the number of resistance parameters!")
    Hvecs = np.asarray(Hvecs, dtype=np.float64).reshape(-1, 3)
    frequencies = np.asarray(frequencies, dtype=np.float64).reshape(-1)
    rows = None
    if Rtype == "Rz":
        if len(layer_ids) == 1:
            raise KeyError("Resistance")  # the reference asks a 1-layer log for a key it never writes (procedures.py:301)
        sd_mode = _abi.SD_RZ
    elif Rtype in ("Rx", "Ry"):
        kind = _resistance_kind(resistance_fn)
        if kind is None:
            raise NotImplementedError("VSD_procedure fuses calculate_resistance_series / calculate_resistance_parallel only")
        rows = _res_rows(resistance_params)
        sd_mode = {("Rx", False): _abi.SD_RX_SERIES, ("Ry", False): _abi.SD_RY_SERIES, ("Rx", True): _abi.SD_RX_PARALLEL,
                   ("Ry", True): _abi.SD_RY_PARALLEL}[(Rtype, kind)]
    else:
        raise ValueError("Rtype must be either Rx or Ry or Rz")
    _, comp, oedriver = _oersted_driver(
        Hoe_direction, ScalarDriver.getSineDriver(0, Hoe_excitation, float(frequencies[0]) if len(frequencies) else 0.0, 0)
    )
    junction.clearLog()
    junction.setLayerOerstedFieldDriver("all", oedriver)
    nH, nF = Hvecs.shape[0], frequencies.shape[0]
    Hall = np.repeat(Hvecs, nF, axis=0)  # instance = h * nF + f (the reference's loop order)
    fall = np.tile(frequencies, nH)
    spectrum = np.zeros(nH * nF)
    if carry_state:
        _warn_sequential("VSD_procedure", "VSD_batch", nH * nF)
        for i in range(nH * nF):
            v, getm = _vsd_batch(junction, Hall[i : i + 1], fall[i : i + 1], _disturbed(junction, layer_ids, 1, disturbance),
                                 int_step, sd_mode, rows, comp, simulation_duration, device, bool(disturbance))
            spectrum[i] = v[0]
            for lid in layer_ids:
                junction.setLayerMagnetisation(lid, CVector(*getm(lid)[0]))
    else:
        m0 = _disturbed(junction, layer_ids, nH * nF, disturbance)
        getm = None
        for lo in range(0, nH * nF, max_batch):
            hi = min(nH * nF, lo + max_batch)
            spectrum[lo:hi], getm = _vsd_batch(junction, Hall[lo:hi], fall[lo:hi], {k: v[lo:hi] for k, v in m0.items()},
                                               int_step, sd_mode, rows, comp, simulation_duration, device, bool(disturbance))
        if getm is not None:
            for lid in layer_ids:
                junction.setLayerMagnetisation(lid, CVector(*getm(lid)[-1]))
    return spectrum.reshape(nH, nF)


#: ``PIMM_procedure`` / ``VSD_procedure`` with every sweep point started fresh and the sweep integrated as one batch
PIMM_batch = functools.partial(PIMM_procedure, carry_state=False)
VSD_batch = functools.partial(VSD_procedure, carry_state=False)
