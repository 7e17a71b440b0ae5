"""ctypes mirror of ``include/cmtj_b200.h`` and the loader of the CUDA shared library.

This is the thin FFI layer that replaces the pybind11 binding of the reference
(``python/cmtj.cpp``) for the junction-integration path.  There is deliberately NO CPU
fallback: if ``libcmtj_b200.so`` is missing or no CUDA device is visible, every compute
call raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ABI_VERSION = 6
MAX_LAYERS = 8
NSLOTS = 30

# --- status codes ---------------------------------------------------------------------
OK, EINVAL, ERUNTIME, ENOTSUP, ECUDA, ENAN = 0, -1, -2, -3, -4, -5

# --- slots (cmtjb_slot) ---------------------------------------------------------------
SLOT_HEXT = 0
SLOT_HOE = 3
SLOT_HDMI = 6
SLOT_HRES = 9
SLOT_K1 = 12
SLOT_K2 = 13
SLOT_IEC_TOP = 14
SLOT_IEC_BOTTOM = 15
SLOT_IECQ_TOP = 16
SLOT_IECQ_BOTTOM = 17
SLOT_IDMI_TOP = 18
SLOT_IDMI_BOTTOM = 21
SLOT_CURRENT = 24
SLOT_FL = 25
SLOT_DL = 26
SLOT_FL2 = 27
SLOT_DL2 = 28
SLOT_TEMPERATURE = 29

FIELD_NAMES = (
    "constantValue",
    "amplitude",
    "frequency",
    "phase",
    "period",
    "cycle",
    "timeStart",
    "timeStop",
    "edgeTime",
    "steadyTime",
)

LOG_FULL, LOG_NONE, LOG_ENERGY = 0, 1, 2


class Driver(C.Structure):
    _fields_ = [("update", C.c_int32), ("custom_column", C.c_int32)] + [(n, C.c_double) for n in FIELD_NAMES]


class Layer(C.Structure):
    _fields_ = [
        ("mag", C.c_double * 3),
        ("anis", C.c_double * 3),
        ("Ms", C.c_double),
        ("thickness", C.c_double),
        ("cellSurface", C.c_double),
        ("damping", C.c_double),
        ("demagTensor", C.c_double * 9),
        ("dipoleBottom", C.c_double * 9),
        ("dipoleTop", C.c_double * 9),
        ("fieldLikeTorque", C.c_double),
        ("dampingLikeTorque", C.c_double),
        ("SlonczewskiSpacerLayerParameter", C.c_double),
        ("beta", C.c_double),
        ("kappa", C.c_double),
        ("spinPolarisation", C.c_double),
        ("referenceLayer", C.c_double * 3),
        ("secondaryReferenceLayer", C.c_double * 3),
        ("Rx0", C.c_double),
        ("Ry0", C.c_double),
        ("AMR_X", C.c_double),
        ("AMR_Y", C.c_double),
        ("SMR_X", C.c_double),
        ("SMR_Y", C.c_double),
        ("AHE", C.c_double),
        ("noiseAlpha", C.c_double),
        ("noiseStd", C.c_double),
        ("noiseScale", C.c_double),
        ("abs_tol", C.c_double),
        ("rel_tol", C.c_double),
        ("max_factor", C.c_double),
        ("min_factor", C.c_double),
        ("safety_factor", C.c_double),
        ("ki", C.c_double),
        ("kp", C.c_double),
        ("kd", C.c_double),
        ("prev_error_ratio", C.c_double),
        ("integral_error", C.c_double),
        ("referenceType", C.c_int32),
        ("includeSTT", C.c_int32),
        ("includeSOT", C.c_int32),
        ("dynamicSOT", C.c_int32),
        ("alternativeSTT", C.c_int32),
        ("temperatureSet", C.c_int32),
        ("noiseAxis", C.c_int32),
        ("use_pid_control", C.c_int32),
        ("drivers", Driver * NSLOTS),
    ]


class Junction(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("mr_mode", C.c_int32),
        ("Rp", C.c_double),
        ("Rap", C.c_double),
        ("MAGNETIC_PERMEABILITY", C.c_double),
        ("GYRO", C.c_double),
        ("TtoAm", C.c_double),
        ("HBAR", C.c_double),
        ("ELECTRON_CHARGE", C.c_double),
        ("BOLTZMANN_CONST", C.c_double),
    ]


class Override(C.Structure):
    _fields_ = [
        ("layer", C.c_int32),
        ("slot", C.c_int32),
        ("field", C.c_int32),
        ("reserved", C.c_int32),
        ("values", C.c_void_p),
    ]


class RunArgs(C.Structure):
    _fields_ = [
        ("totalTime", C.c_double),
        ("timeStep", C.c_double),
        ("writeFrequency", C.c_double),
        ("solverMode", C.c_int32),
        ("log_mode", C.c_int32),
        ("seed", C.c_uint64),
        ("n_instances", C.c_int64),
        ("instance_offset", C.c_int64),
        ("start_iteration", C.c_int64),
        ("onef_buffer", C.c_void_p),
        ("custom_table", C.c_void_p),
        ("n_custom", C.c_int64),
        ("thermal_normals", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("effective_solver", C.c_int32),
        ("solver_overridden", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_channels", C.c_int32),
        ("total_iterations", C.c_int64),
        ("write_every", C.c_int64),
        ("n_samples", C.c_int64),
        ("variable_samples", C.c_int32),
        ("feature_mask", C.c_int32),
        ("variant", C.c_char * 64),
    ]


LAYER_SYSTEM = -2
STACK_SERIES, STACK_PARALLEL = 0, 1
INT_NULL, INT_DIPOLE, INT_DIPOLE_NOMURA, INT_TABLE = 0, 1, 2, 3


class Stack(C.Structure):
    _fields_ = [
        ("n_junctions", C.c_int32),
        ("kind", C.c_int32),
        ("use_kcl", C.c_int32),
        ("delayed", C.c_int32),
        ("top_layer", C.c_int32),
        ("bottom_layer", C.c_int32),
        ("reserved0", C.c_int32),
        ("reserved1", C.c_int32),
        ("phaseOffset", C.c_double),
        ("polarisation", C.c_double * 3),
        ("coupling", C.c_void_p),
        ("currentDriver", Driver),
    ]


class Group(C.Structure):
    _fields_ = [
        ("n_junctions", C.c_int32),
        ("interaction", C.c_int32),
        ("top_layer", C.c_int32),
        ("reserved", C.c_int32),
        ("coordinates", C.c_void_p),
        ("pair_table", C.c_void_p),
    ]


class CoupledPlan(C.Structure):
    _fields_ = [
        ("effective_solver", C.c_int32),
        ("solver_overridden", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_junction_channels", C.c_int32),
        ("n_system_channels", C.c_int32),
        ("cooperative", C.c_int32),
        ("total_iterations", C.c_int64),
        ("write_every", C.c_int64),
        ("n_samples", C.c_int64),
    ]


MAX_PROJ, POST_MAX_LAYERS, RES_NFIELDS = 4, 4, 6
SD_NONE, SD_RZ, SD_RX_SERIES, SD_RY_SERIES, SD_RX_PARALLEL, SD_RY_PARALLEL = range(6)


class Post(C.Structure):
    """cmtjb_post_t: fused reductions of the sweep procedures (projection log, spin-diode IIR)."""

    _fields_ = [
        ("n_proj", C.c_int32),
        ("sd_mode", C.c_int32),
        ("proj", (C.c_double * (3 * POST_MAX_LAYERS)) * MAX_PROJ),
        ("res", (C.c_double * RES_NFIELDS) * POST_MAX_LAYERS),
        ("sd_b", C.c_double * 4),
        ("sd_a", C.c_double * 4),
        ("sd_omega", C.c_void_p),
        ("proj_log", C.c_void_p),
        ("proj_n_samples", C.c_int64),
        ("proj_sample0", C.c_int64),
        ("sd_state", C.c_void_p),
    ]


_P = C.POINTER


def raise_for_status(rc: int, message: str) -> None:
    """Map a C-ABI status to the exception pybind11 would raise for the reference
    (SURVEY 8a note 11): invalid_argument -> ValueError, runtime_error -> RuntimeError."""
    if rc == OK:
        return
    if rc == EINVAL:
        raise ValueError(message)
    if rc == ENOTSUP:
        raise NotImplementedError(message)
    raise RuntimeError(message)


# cmtjb_exchange_fn (include/cmtj_b200.h): (ctx, buf, count, step, stream) -> 0 on success
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p)


class Backend:
    """Uniform view over a shared library exporting the run/plan entry points.

    The product uses exactly one instance of this class, bound to ``libcmtj_b200.so``
    (`cuda_backend()`).  (tests/helpers.py binds a second instance to the CPU oracle and
    monkeypatches it in as the checker; nothing in this package knows about the oracle.)
    """

    def __init__(self, lib: C.CDLL, prefix: str, has_device_arg: bool):
        self.lib = lib
        self.prefix = prefix
        self.has_device_arg = has_device_arg
        f = getattr(lib, prefix + "plan")
        f.restype = C.c_int
        f.argtypes = [_P(Junction), _P(Layer), _P(RunArgs), _P(Plan)]
        self._plan = f
        f = getattr(lib, prefix + "run_host")
        f.restype = C.c_int
        f.argtypes = [
            _P(Junction),
            _P(Layer),
            _P(Override),
            C.c_int32,
            _P(RunArgs),
            C.c_void_p,
            C.c_void_p,
            C.c_void_p,
            C.c_void_p,
            C.c_void_p,
        ] + ([C.c_int] if has_device_arg else [])
        self._run_host = f
        f = getattr(lib, prefix + "last_error")
        f.restype = C.c_char_p
        f.argtypes = []
        self._last_error = f

    # ---- coupled systems (Stack / GroupInteraction) ------------------------------------------
    def _fn(self, name, argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        f.argtypes = argtypes
        return f

    def stack_plan(self, st: "Stack", jns, layers, args: RunArgs) -> "CoupledPlan":
        p = CoupledPlan()
        f = self._fn("stack_plan", [_P(Stack), C.c_void_p, C.c_void_p, _P(RunArgs), _P(CoupledPlan)])
        self.check(f(C.byref(st), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.byref(args), C.byref(p)))
        return p

    def group_plan(self, g: "Group", jns, layers, args: RunArgs) -> "CoupledPlan":
        p = CoupledPlan()
        f = self._fn("group_plan", [_P(Group), C.c_void_p, C.c_void_p, _P(RunArgs), _P(CoupledPlan)])
        self.check(f(C.byref(g), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.byref(args), C.byref(p)))
        return p

    @staticmethod
    def _ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def stack_run_host(self, st, jns, layers, overrides, n_overrides, args, m_state, jlog, slog, status, device=0) -> int:
        f = self._fn(
            "stack_run_host",
            [_P(Stack), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _P(RunArgs)] + [C.c_void_p] * 4
            + ([C.c_int] if self.has_device_arg else []),
        )  # fmt: skip
        extra = (device,) if self.has_device_arg else ()
        return f(
            C.byref(st), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.cast(overrides, C.c_void_p), n_overrides,
            C.byref(args), self._ptr(m_state), self._ptr(jlog), self._ptr(slog), self._ptr(status), *extra,
        )  # fmt: skip

    def group_run_host(self, g, jns, layers, overrides, n_overrides, args, m_state, jlog, status, device=0) -> int:
        f = self._fn(
            "group_run_host",
            [_P(Group), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _P(RunArgs)] + [C.c_void_p] * 3
            + ([C.c_int] if self.has_device_arg else []),
        )  # fmt: skip
        extra = (device,) if self.has_device_arg else ()
        return f(
            C.byref(g), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.cast(overrides, C.c_void_p), n_overrides,
            C.byref(args), self._ptr(m_state), self._ptr(jlog), self._ptr(status), *extra,
        )  # fmt: skip

    def group_run_sharded(self, g, jns, layers, args, rank, world, exchange, m_state_ptr, jlog_ptr, status_ptr, device=0, stream=0) -> int:
        """``cmtjb_group_run_sharded``: one rank of a dipole array split over ``world`` processes.  The state / log /
        status arguments are raw addresses (device memory for the CUDA library); ``exchange`` is an ``EXCHANGE_FN``."""
        f = self._fn(
            "group_run_sharded",
            [_P(Group), C.c_void_p, C.c_void_p, _P(RunArgs), C.c_int32, C.c_int32, EXCHANGE_FN, C.c_void_p]
            + [C.c_void_p] * 3 + [C.c_int, C.c_void_p],
        )  # fmt: skip
        return f(
            C.byref(g), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.byref(args), int(rank), int(world), exchange, None,
            C.c_void_p(m_state_ptr), C.c_void_p(jlog_ptr), C.c_void_p(status_ptr), int(device), C.c_void_p(stream),
        )  # fmt: skip

    def group_run_sharded_peer(self, g, jns, layers, args, rank, world, areas, m_state_ptr, jlog_ptr, status_ptr, device=0, stream=0) -> int:
        """``cmtjb_group_run_sharded_peer``: the exchange over peer memory; ``areas`` = the ``world`` exchange-area addresses
        as seen from this process."""
        f = self._fn(
            "group_run_sharded_peer",
            [_P(Group), C.c_void_p, C.c_void_p, _P(RunArgs), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)] + [C.c_void_p] * 3
            + [C.c_int, C.c_void_p],
        )  # fmt: skip
        arr = (C.c_void_p * int(world))(*[int(a) for a in areas])
        return f(
            C.byref(g), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.byref(args), int(rank), int(world), arr,
            C.c_void_p(m_state_ptr), C.c_void_p(jlog_ptr), C.c_void_p(status_ptr), int(device), C.c_void_p(stream),
        )  # fmt: skip

    def last_error(self) -> str:
        s = self._last_error()
        return s.decode("utf-8", "replace") if s else ""

    def check(self, rc: int) -> None:
        if rc != OK:
            raise_for_status(rc, self.last_error())

    def plan(self, jn: Junction, layers, args: RunArgs) -> Plan:
        p = Plan()
        self.check(self._plan(C.byref(jn), layers, C.byref(args), C.byref(p)))
        return p

    def run_host(
        self,
        jn: Junction,
        layers,
        overrides,
        n_overrides: int,
        args: RunArgs,
        m_state: np.ndarray,
        log: Optional[np.ndarray],
        status: np.ndarray,
        n_samples_out: Optional[np.ndarray],
        time_log: Optional[np.ndarray],
        device: int = 0,
    ) -> int:
        def ptr(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        extra = (device,) if self.has_device_arg else ()
        return self._run_host(
            C.byref(jn),
            layers,
            overrides,
            n_overrides,
            C.byref(args),
            ptr(m_state),
            ptr(log),
            ptr(status),
            ptr(n_samples_out),
            ptr(time_log),
            *extra,
        )


_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcmtj_b200.so")
_cuda: Optional[Backend] = None
_cuda_lib: Optional[C.CDLL] = None


def last_variant() -> str:
    """Name of the kernel variant the calling thread's last sweep launch selected (``cmtjb_last_variant``)."""
    lib = load_library()
    lib.cmtjb_last_variant.restype = C.c_char_p
    return (lib.cmtjb_last_variant() or b"").decode()


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(cmtj_b200 has no CPU fallback)"
            )
        _cuda_lib = C.CDLL(_LIB_PATH)
        _cuda_lib.cmtjb_abi_version.restype = C.c_int
        if _cuda_lib.cmtjb_abi_version() != ABI_VERSION:
            raise RuntimeError("libcmtj_b200.so ABI version mismatch: rebuild the library")
        _cuda_lib.cmtjb_device_count.restype = C.c_int
    return _cuda_lib


def cuda_backend() -> Backend:
    global _cuda
    if _cuda is None:
        lib = load_library()
        if lib.cmtjb_device_count() <= 0:
            raise RuntimeError(
                "cmtj_b200: no CUDA device visible; the junction integrator runs only on the GPU "
                "(there is no CPU fallback)"
            )
        _cuda = Backend(lib, "cmtjb_", has_device_arg=True)
    return _cuda


def backend() -> Backend:
    """The backend every compute call of the facade goes through: always the CUDA library."""
    return cuda_backend()
