"""Batched parameter sweeps: ``n_instances`` copies of one :class:`~cmtj_b200.core.Junction`
integrated by a single fused kernel launch, one GPU thread per instance.

This is the entry the reference does not have: its sweeps are Python ``for`` loops over
``runSimulation`` (cmtj/utils/procedures.py:120-141, curated-examples/*.py).  Instances
share the junction's structure and differ by

* their magnetisation state (``set_magnetisation`` / carried over between runs), and
* per-instance *overrides* of any driver field of any layer, e.g. the external field
  vector, a current amplitude, a sine frequency (``set_override`` and helpers).

Two ways to run:

``run(...)``         host (numpy) buffers through ``cmtjb_run_host``; H2D/D2H included.
``run_device(...)``  state, overrides and logs stay on the GPU as ``torch`` tensors
                     (``cmtjb_run_device``, asynchronous on the current torch stream).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _abi
from . import core as _core
from .core import CVector, Junction, SolverMode, _next_seed

_FIELD_INDEX = {n: i for i, n in enumerate(_abi.FIELD_NAMES)}
_SLOT_BY_NAME = {
    "Hext_x": 0, "Hext_y": 1, "Hext_z": 2,
    "Hoe_x": 3, "Hoe_y": 4, "Hoe_z": 5,
    "Hdmi_x": 6, "Hdmi_y": 7, "Hdmi_z": 8,
    "Hres_x": 9, "Hres_y": 10, "Hres_z": 11,
    "K1": 12, "K2": 13,
    "IEC_top": 14, "IEC_bottom": 15, "IECquad_top": 16, "IECquad_bottom": 17,
    "IDMI_top_x": 18, "IDMI_top_y": 19, "IDMI_top_z": 20,
    "IDMI_bottom_x": 21, "IDMI_bottom_y": 22, "IDMI_bottom_z": 23,
    "current": 24, "fieldLike": 25, "dampingLike": 26, "fieldLike2": 27, "dampingLike2": 28,
    "temperature": 29,
}  # fmt: skip


class BatchLog(dict):
    """name -> array [n_samples, n_instances] (``time``: [n_samples], or [n_samples, B] for
    Dormand-Prince).  ``instance(i)`` gives the reference-shaped ``getLog()`` dict of one instance."""

    n_samples_per_instance: Optional[np.ndarray] = None

    def instance(self, i: int) -> Dict[str, np.ndarray]:
        n = None if self.n_samples_per_instance is None else int(self.n_samples_per_instance[i])
        out = {}
        for k, v in self.items():
            a = v if v.ndim == 1 else v[:, i]
            out[k] = np.asarray(a[:n])
        return out

    # ---- on-disk formats ---------------------------------------------------------------
    def save_csv(self, filename: str, instance: int) -> None:
        """One instance in the reference's ``saveLog`` format (Junction::saveLogs, junction.hpp:1638-1656):
        ';'-terminated fields, a header row of keys, one row per sample, values in the C++ stream
        default (``%g``).  Columns follow this log's key order (the reference's is its hash order)."""
        if filename == "":
            raise RuntimeError("The filename may not be empty!")
        cols = self.instance(instance)
        keys = list(cols.keys())
        n = min((len(v) for v in cols.values()), default=0)
        with open(filename, "w") as f:
            f.write("".join(k + ";" for k in keys) + "\n")
            for i in range(n):
                f.write("".join(f"{cols[k][i]:g};" for k in keys) + "\n")

    def save_npz(self, filename: str, compressed: bool = False) -> None:
        """The whole sweep in one ``.npz``: every channel as float64 [n_samples, B] (lossless, unlike the
        6-digit CSV), plus ``__n_samples_per_instance`` for adaptive-step runs."""
        arrays = {k: np.asarray(v) for k, v in self.items()}
        if self.n_samples_per_instance is not None:
            arrays["__n_samples_per_instance"] = np.asarray(self.n_samples_per_instance)
        (np.savez_compressed if compressed else np.savez)(filename, **arrays)

    @classmethod
    def load_npz(cls, filename: str) -> "BatchLog":
        out = cls()
        with np.load(filename) as z:
            for k in z.files:
                if k == "__n_samples_per_instance":
                    out.n_samples_per_instance = z[k]
                else:
                    out[k] = z[k]
        return out


class JunctionBatch:
    def __init__(self, junction: Junction, n_instances: int, device: int = 0, instance_offset: int = 0):
        if n_instances < 0:
            raise ValueError("n_instances must be non-negative")
        self.junction = junction
        self.B = int(n_instances)
        self.device = int(device)
        self.instance_offset = int(instance_offset)
        L = junction.layerNo
        m0 = np.array([c for l in junction._layers for c in l.mag.tolist()], dtype=np.float64)
        # [L*3][B], instance fastest
        self.m_state = np.repeat(m0[:, None], self.B, axis=1).copy()
        self._m_dev = None
        self._overrides: Dict[tuple, np.ndarray] = {}
        self._ovr_dev: Dict[tuple, object] = {}
        self.last_plan: Optional[_abi.Plan] = None

    # ---- per-instance inputs ----------------------------------------------------------------------
    def _layer_index(self, layer_id: str) -> int:
        if layer_id == "all":
            return -1
        ids = self.junction.getLayerIds()
        if layer_id not in ids:
            raise RuntimeError("Failed to find a layer with a given id: " + layer_id + "!")
        return ids.index(layer_id)

    def set_override(self, layer_id: str, slot: Union[int, str], field: Union[int, str], values) -> None:
        """values[i] replaces ``<layer>.<slot driver>.<field>`` for instance i."""
        s = _SLOT_BY_NAME[slot] if isinstance(slot, str) else int(slot)
        f = _FIELD_INDEX[field] if isinstance(field, str) else int(field)
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        if v.shape[0] != self.B:
            raise ValueError(f"override needs {self.B} values, got {v.shape[0]}")
        key = (self._layer_index(layer_id), s, f)
        self._overrides[key] = v
        self._ovr_dev.pop(key, None)

    def clear_overrides(self) -> None:
        self._overrides.clear()
        self._ovr_dev.clear()

    def set_external_field(self, layer_id: str, H) -> None:
        """Per-instance constant external field, H: [B,3] in A/m (the batched form of
        ``setLayerExternalFieldDriver(id, AxialDriver(Hx,Hy,Hz))``)."""
        H = np.asarray(H, dtype=np.float64).reshape(self.B, 3)
        for c in range(3):
            self.set_override(layer_id, _abi.SLOT_HEXT + c, 0, H[:, c])

    def set_magnetisation(self, layer_id: str, m, normalise: bool = True) -> None:
        """Per-instance magnetisation, m: [B,3]; normalised like ``Layer::setMagnetisation``
        (junction.hpp:491-498).  ``normalise=False`` stores m as given: the state of a layer whose
        magnetisation was only ever set by the ``Layer`` constructor, which keeps the un-normalised
        input until the first step (junction.hpp:178-195)."""
        m = np.asarray(m, dtype=np.float64).reshape(self.B, 3)
        n = np.sqrt((m * m).sum(axis=1))
        if np.any(n == 0):
            raise RuntimeError("Initial magnetisation was set to a zero vector!")
        if normalise:
            m = m / n[:, None]
        self.sync_state_to_host()
        for l in range(self.junction.layerNo) if layer_id == "all" else [self._layer_index(layer_id)]:
            self.m_state[3 * l : 3 * l + 3, :] = m.T
        self._m_dev = None

    def get_magnetisation(self, layer_id: str) -> np.ndarray:
        self.sync_state_to_host()
        l = self._layer_index(layer_id)
        return self.m_state[3 * l : 3 * l + 3, :].T.copy()

    def sync_state_to_host(self) -> None:
        if self._m_dev is not None:
            self.m_state = self._m_dev.cpu().numpy()

    # ---- shared plumbing -------------------------------------------------------------------------------
    def _args(self, totalTime, timeStep, writeFrequency, solverMode, log, seed, onef, start_iteration=0) -> _abi.RunArgs:
        a = _abi.RunArgs()
        a.thermal_normals = _core._thermal_normals
        a.totalTime = float(totalTime)
        a.timeStep = float(timeStep)
        a.writeFrequency = float(writeFrequency)
        a.solverMode = int(solverMode)
        a.log_mode = _abi.LOG_FULL if log else _abi.LOG_NONE
        a.seed = _next_seed() if seed is None else int(seed)
        a.n_instances = self.B
        a.instance_offset = self.instance_offset
        a.start_iteration = int(start_iteration)
        a.onef_buffer = onef
        return a

    def _override_table(self, pointers: Sequence[int]):
        n = len(self._overrides)
        tab = (_abi.Override * max(n, 1))()
        for k, ((l, s, f), _) in enumerate(self._overrides.items()):
            tab[k].layer, tab[k].slot, tab[k].field, tab[k].values = l, s, f, pointers[k]
        return tab, n

    def selected_variant(self, totalTime: float, timeStep: float = 1e-13, writeFrequency: float = 1e-11,
                         solverMode: SolverMode = SolverMode.RK4) -> str:
        """Name of the compiled kernel variant a run with these arguments and the current per-instance overrides
        would launch for its driver-evaluating segment (``cmtjb_select_variant``; host logic, no GPU needed)."""
        lib = _abi.load_library()
        jn, layers = self.junction._pack()
        a = self._args(totalTime, timeStep, writeFrequency, solverMode, True, 0, None)
        tab, n = self._override_table([v.ctypes.data for v in self._overrides.values()])
        name = C.create_string_buffer(96)
        f = lib.cmtjb_select_variant
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_char_p, C.c_int32]
        rc = f(C.byref(jn), layers, tab, n, C.byref(a), name, 96)
        if rc != _abi.OK:
            lib.cmtjb_last_error.restype = C.c_char_p
            raise RuntimeError((lib.cmtjb_last_error() or b"").decode())
        return name.value.decode()

    def _package(self, plan: _abi.Plan, log: Optional[np.ndarray], timeStep: float, tlog, nsamp) -> BatchLog:
        out = BatchLog()
        if log is None:
            return out
        for ch, name in enumerate(self.junction._channel_names()):
            out[name] = log[ch]
        if plan.variable_samples:
            out["time"] = tlog
            out.n_samples_per_instance = nsamp
        else:
            out["time"] = np.arange(0, int(plan.total_iterations), int(plan.write_every), dtype=np.float64) * float(
                timeStep
            )
        return out

    # ---- host path -----------------------------------------------------------------------------------------
    def run(
        self,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        solverMode: SolverMode = SolverMode.RK4,
        log: bool = True,
        seed: Optional[int] = None,
        out: Optional[np.ndarray] = None,
    ) -> BatchLog:
        """Integrate every instance; host buffers in, host buffers out (H2D/D2H inside).
        ``out`` may be a preallocated (e.g. pinned) float64 array [n_channels, n_samples, B]."""
        self.sync_state_to_host()
        self._m_dev = None
        be = _abi.backend()
        jn, layers = self.junction._pack()
        a = self._args(totalTime, timeStep, writeFrequency, solverMode, log, seed, None)
        plan = be.plan(jn, layers, a)
        self.last_plan = plan
        custom = self.junction._custom_table(layers, plan, timeStep)
        if custom is not None:
            a.custom_table = custom.ctypes.data
            a.n_custom = custom.shape[1]
        n_ch, n_s = int(plan.n_channels), int(plan.n_samples)
        buf = None
        if log:
            if out is not None:
                if out.shape != (n_ch, n_s, self.B) or out.dtype != np.float64 or not out.flags.c_contiguous:
                    raise ValueError(f"out must be C-contiguous float64 of shape {(n_ch, n_s, self.B)}")
                buf = out
            else:
                buf = np.empty((n_ch, n_s, self.B), dtype=np.float64)
        status = np.zeros(self.B, dtype=np.int32)
        nsamp = np.zeros(self.B, dtype=np.int32) if plan.variable_samples else None
        tlog = np.zeros((n_s, self.B), dtype=np.float64) if plan.variable_samples and log else None
        tab, n = self._override_table([v.ctypes.data for v in self._overrides.values()])
        rc = be.run_host(jn, layers, tab, n, a, self.m_state, buf, status, nsamp, tlog, self.device)
        self.status = status
        be.check(rc)
        return self._package(plan, buf, timeStep, tlog, nsamp)

    # ---- device path -----------------------------------------------------------------------------------------
    def run_device(
        self,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        solverMode: SolverMode = SolverMode.RK4,
        log: bool = True,
        seed: Optional[int] = None,
        out=None,
        post: Optional["PostSpec"] = None,
        start_iteration: int = 0,
    ):
        """Same run with everything resident on the GPU.  Returns a dict of ``torch`` tensors
        (views of one [n_channels, n_samples, B] tensor); asynchronous on the current stream.

        ``post`` fuses the reductions of the sweep procedures into the time loop
        (``cmtjb_run_device_post``: projection log, spin-diode IIR); ``start_iteration`` continues a
        run whose first steps were integrated by an earlier call (state carried on the device)."""
        import torch

        be = _abi.cuda_backend()
        lib = be.lib
        dev = torch.device("cuda", self.device)
        jn, layers = self.junction._pack()
        a = self._args(totalTime, timeStep, writeFrequency, solverMode, log, seed, None, start_iteration)
        plan = be.plan(jn, layers, a)
        self.last_plan = plan
        n_ch, n_s = int(plan.n_channels), int(plan.n_samples)
        custom = self.junction._custom_table(layers, plan, timeStep)
        if custom is not None:
            custom_dev = torch.from_numpy(custom).to(dev)  # kept alive until the call below has been enqueued
            a.custom_table = custom_dev.data_ptr()
            a.n_custom = custom.shape[1]
        if self._m_dev is None:
            self._m_dev = torch.from_numpy(self.m_state).to(dev)
        for key, v in self._overrides.items():
            if key not in self._ovr_dev:
                self._ovr_dev[key] = torch.from_numpy(v).to(dev)
        tab, n = self._override_table([self._ovr_dev[k].data_ptr() for k in self._overrides])
        buf = None
        if log:
            buf = out if out is not None else torch.empty((n_ch, n_s, self.B), dtype=torch.float64, device=dev)
            assert tuple(buf.shape) == (n_ch, n_s, self.B) and buf.is_contiguous() and buf.dtype == torch.float64
        status = torch.zeros(self.B, dtype=torch.int32, device=dev)
        nsamp = torch.zeros(self.B, dtype=torch.int32, device=dev) if plan.variable_samples else None
        tlog = torch.zeros((n_s, self.B), dtype=torch.float64, device=dev) if plan.variable_samples and log else None

        def p(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        stream = torch.cuda.current_stream(dev).cuda_stream
        if post is None:
            f = lib.cmtjb_run_device
            f.restype = C.c_int
            f.argtypes = [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
            rc = f(
                C.byref(jn), layers, tab, n, C.byref(a), p(self._m_dev), p(buf), p(status), p(nsamp), p(tlog),
                self.device, C.c_void_p(stream),
            )  # fmt: skip
        else:
            f = lib.cmtjb_run_device_post
            f.restype = C.c_int
            f.argtypes = [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
            pc = post._pack(self.B, L=self.junction.layerNo)
            rc = f(
                C.byref(jn), layers, tab, n, C.byref(a), C.byref(pc), p(self._m_dev), p(buf), p(status),
                self.device, C.c_void_p(stream),
            )  # fmt: skip
        be.check(rc)
        self.status_dev = status
        res = {}
        if buf is not None:
            for ch, name in enumerate(self.junction._channel_names()):
                res[name] = buf[ch]
            if plan.variable_samples:
                res["time"] = tlog
                res["n_samples"] = nsamp
            else:
                we = int(plan.write_every)
                first = -(-int(start_iteration) // we) * we  # first logged step index of this call
                res["time"] = torch.arange(first, int(plan.total_iterations), we, dtype=torch.float64, device=dev) * float(
                    timeStep
                )
        res["_log"] = buf
        return res


class PostSpec:
    """Python view of ``cmtjb_post_t`` (include/cmtj_b200.h): what to reduce inside the time loop.

    proj        [n_proj, 3*L] projection rows (row k: weights on mx,my,mz of every layer)
    proj_log    torch float64 [n_proj, proj_n_samples, B] on the device
    sd_mode     _abi.SD_*; res [L, 6] ResistanceParameters rows (Rxx0, Rxy0, Ramr, Rahe, Rsmr, w/l)
    sd_b, sd_a  order-3 IIR; sd_omega torch [B]; sd_state torch [4, B]
    """

    def __init__(self, proj=None, proj_log=None, proj_sample0=0, sd_mode=_abi.SD_NONE, res=None, sd_b=None, sd_a=None,
                 sd_omega=None, sd_state=None):
        self.proj = None if proj is None else np.atleast_2d(np.asarray(proj, dtype=np.float64))
        self.proj_log = proj_log
        self.proj_sample0 = int(proj_sample0)
        self.sd_mode = int(sd_mode)
        self.res = None if res is None else np.atleast_2d(np.asarray(res, dtype=np.float64))
        self.sd_b, self.sd_a = sd_b, sd_a
        self.sd_omega, self.sd_state = sd_omega, sd_state

    def _pack(self, B: int, L: int) -> _abi.Post:
        c = _abi.Post()
        if self.proj is not None:
            k, w = self.proj.shape
            if k > _abi.MAX_PROJ or w != 3 * L:
                raise ValueError(f"proj must be [<= {_abi.MAX_PROJ}, {3 * L}]")
            t = self.proj_log
            if t is None or t.dim() != 3 or t.shape[0] != k or t.shape[2] != B or not t.is_contiguous():
                raise ValueError("proj_log must be a contiguous float64 device tensor [n_proj, n_samples, B]")
            c.n_proj = k
            for r in range(k):
                for q in range(w):
                    c.proj[r][q] = float(self.proj[r, q])
            c.proj_log = t.data_ptr()
            c.proj_n_samples = int(t.shape[1])
            c.proj_sample0 = self.proj_sample0
        c.sd_mode = self.sd_mode
        if self.res is not None:
            if self.res.shape != (L, _abi.RES_NFIELDS):
                raise ValueError(f"res must be [{L}, {_abi.RES_NFIELDS}]")
            for l in range(L):
                for q in range(_abi.RES_NFIELDS):
                    c.res[l][q] = float(self.res[l, q])
        if self.sd_mode != _abi.SD_NONE:
            for q in range(4):
                c.sd_b[q] = float(self.sd_b[q])
                c.sd_a[q] = float(self.sd_a[q])
            if tuple(self.sd_omega.shape) != (B,) or tuple(self.sd_state.shape) != (4, B):
                raise ValueError("sd_omega must be [B] and sd_state [4, B]")
            c.sd_omega = self.sd_omega.data_ptr()
            c.sd_state = self.sd_state.data_ptr()
        return c
