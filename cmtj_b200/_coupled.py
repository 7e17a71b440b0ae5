"""Shared host plumbing of the coupled systems (Stack, GroupInteraction): packs N junctions into
the C-ABI records and runs ``n_instances`` copies of the whole system in one call."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi
from . import core as _core
from .core import Junction, _next_seed


def _uint(value, what: str) -> int:
    """pybind11 ``unsigned int`` argument: negative / non-integers are a TypeError."""
    if isinstance(value, bool) or not isinstance(value, (int, np.integer)) or value < 0:
        raise TypeError(f"{what}: incompatible function arguments (expected unsigned int)")
    return int(value)


def pack_junctions(junctions: Sequence[Junction]):
    n = len(junctions)
    nl = junctions[0].layerNo
    jns = (_abi.Junction * n)()
    layers = (_abi.Layer * (n * nl))()
    for k, j in enumerate(junctions):
        if j.layerNo != nl:
            raise NotImplementedError("junctions of a coupled system must all have the same number of layers in this build")
        jn, ls = j._pack()
        C.memmove(C.byref(jns, k * C.sizeof(_abi.Junction)), C.byref(jn), C.sizeof(_abi.Junction))
        C.memmove(C.byref(layers, k * nl * C.sizeof(_abi.Layer)), ls, nl * C.sizeof(_abi.Layer))
    return jns, layers, nl


def run_args(totalTime, timeStep, writeFrequency, n_instances, log=True, seed=None, instance_offset=0) -> _abi.RunArgs:
    a = _abi.RunArgs()
    a.thermal_normals = _core._thermal_normals
    a.totalTime = float(totalTime)
    a.timeStep = float(timeStep)
    a.writeFrequency = float(writeFrequency)
    a.solverMode = 1  # the reference always asks for RK4 here (stack.hpp:50, reservoir.hpp:200)
    a.log_mode = _abi.LOG_FULL if log else _abi.LOG_NONE
    a.seed = _next_seed() if seed is None else int(seed)
    a.n_instances = int(n_instances)
    a.instance_offset = int(instance_offset)
    a.start_iteration = 0
    a.onef_buffer = None
    return a


def initial_state(junctions: Sequence[Junction], n_instances: int) -> np.ndarray:
    """[n_layers*3][n_instances * N], element = instance * N + junction."""
    nl = junctions[0].layerNo
    one = np.array([[c for c in l.mag.tolist()] for j in junctions for l in j._layers], dtype=np.float64)
    one = one.reshape(len(junctions), nl * 3).T  # [nl*3, N]
    return np.ascontiguousarray(np.tile(one, (1, n_instances)))


def override_table(overrides: Dict[tuple, np.ndarray]):
    n = len(overrides)
    tab = (_abi.Override * max(n, 1))()
    for k, ((l, s, f), v) in enumerate(overrides.items()):
        tab[k].layer, tab[k].slot, tab[k].field, tab[k].values = l, s, f, v.ctypes.data
    return tab, n


def sample_times(plan, timeStep: float) -> np.ndarray:
    return np.arange(0, int(plan.total_iterations), int(plan.write_every), dtype=np.float64) * float(timeStep)


def write_back(junctions: Sequence[Junction], m_state: np.ndarray, jlog: Optional[np.ndarray], times, inst: int = 0) -> None:
    """Persist instance `inst` into the junction objects: m carried to the next run, logs appended."""
    from .core import CVector

    n = len(junctions)
    for k, j in enumerate(junctions):
        e = inst * n + k
        for li, l in enumerate(j._layers):
            l.mag = CVector(m_state[3 * li, e], m_state[3 * li + 1, e], m_state[3 * li + 2, e])
        if jlog is not None and len(times):
            j._append_samples(jlog[:, :, e], times)
