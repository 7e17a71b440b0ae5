"""``cmtj.reservoir``: dipole-coupled junction arrays (``GroupInteraction``,
core/reservoir.hpp:122-245, bound at python/cmtj.cpp:429-445), integrated on the GPU by
``group_block_kernel`` / ``group_grid_kernel`` through ``cmtjb_group_run_host``."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _abi, _coupled
from .constants import PhysicalConstants
from .core import CVector, Junction, Layer, c_dot


def nullDipoleInteraction(r1, r2, layer1: Layer, layer2: Layer) -> CVector:
    """core/reservoir.hpp:63-68"""
    return CVector(0, 0, 0)


def computeDipoleInteraction(r1, r2, layer1: Layer, layer2: Layer) -> CVector:
    """Kanao et al. (core/reservoir.hpp:79-97); host formula for the API function only."""
    r1, r2 = CVector._coerce(r1), CVector._coerce(r2)
    rij = r1 - r2
    if rij.length() < 1e-10:
        raise RuntimeError("Points are too close for stable dipole calculation")
    r3, r5 = rij.length() ** 3, rij.length() ** 5
    m1 = layer2.mag
    pref = (layer2.Ms / PhysicalConstants.magnetic_permeability()) * (layer2.thickness * layer2.cellSurface)
    return ((rij * (3 * c_dot(m1, rij))) * (1.0 / r5) - m1 * (1.0 / r3)) * pref


def computeDipoleInteractionNoumra(r1, r2, layer1: Layer, layer2: Layer) -> CVector:
    """Nomura et al. (core/reservoir.hpp:36-56, :108-120)."""
    r1, r2 = CVector._coerce(r1), CVector._coerce(r2)
    rij = r1 - r2
    if rij.length() < 1e-14:
        raise RuntimeError("Points are too close for stable dipole calculation")
    r_mag = rij.length() ** 2
    mult = 3 / (4 * math.pi * rij.length() ** 5)
    rows = [
        CVector(rij.x**2 - r_mag / 3, rij.x * rij.y, rij.x * rij.z) * mult,
        CVector(rij.x * rij.y, rij.y**2 - r_mag / 3, rij.y * rij.z) * mult,
        CVector(rij.x * rij.z, rij.y * rij.z, rij.z**2 - r_mag / 3) * mult,
    ]
    m = layer2.mag
    Ms = layer2.Ms * layer2.thickness * layer2.cellSurface
    return CVector(*(c_dot(row, m) for row in rows)) * (Ms / PhysicalConstants.magnetic_permeability())


_INTERACTIONS = {
    nullDipoleInteraction: _abi.INT_NULL,
    computeDipoleInteraction: _abi.INT_DIPOLE,
    computeDipoleInteractionNoumra: _abi.INT_DIPOLE_NOMURA,
}


class GroupInteraction:
    def __init__(self, coordinateMatrix: Sequence, junctionList: Sequence[Junction], topId: str = "free"):
        coords = [CVector._coerce(c).copy() for c in coordinateMatrix]
        junctionList = list(junctionList)
        if len(coords) != len(junctionList):
            raise RuntimeError("Coordinate matrix and junction list must have the same size!")
        if not coords or not junctionList:
            raise RuntimeError("Coordinate matrix and junction list cannot be empty!")
        self.topId = str(topId)
        self.noElements = len(junctionList)
        self.coordinateMatrix = coords
        self.junctionList: List[Junction] = [j._clone() for j in junctionList]
        xyz = np.array([c.tolist() for c in coords], dtype=np.float64)
        if len(np.unique(xyz, axis=0)) != len(xyz):
            raise RuntimeError("Coordinate vectors must be unique!")
        self._interaction = _abi.INT_DIPOLE

    def setInteractionFunction(self, interactionFunc: Callable) -> None:
        if interactionFunc not in _INTERACTIONS:
            raise NotImplementedError(
                "only nullDipoleInteraction, computeDipoleInteraction and computeDipoleInteractionNoumra run on the device"
            )
        self._interaction = _INTERACTIONS[interactionFunc]

    def clearLogs(self) -> None:
        for j in self.junctionList:
            j.clearLog()

    def getLog(self, id: Optional[int] = None) -> Dict[str, List[float]]:
        if id is None:
            raise RuntimeError("Not implemented!")
        k = _coupled._uint(id, "getLog")
        if k < self.noElements:
            return self.junctionList[k].getLog()
        raise RuntimeError("Asking for id of a non-existing junction!")

    def _pack(self):
        jns, layers, nl = _coupled.pack_junctions(self.junctionList)
        g = _abi.Group()
        g.n_junctions = self.noElements
        g.interaction = self._interaction
        ids = self.junctionList[0].getLayerIds()
        for j in self.junctionList:
            j.getLayer(self.topId)
            if j.getLayerIds() != ids:
                raise NotImplementedError("junctions of a group must use the same layer ids in the same order in this build")
        g.top_layer = ids.index(self.topId)
        xyz = np.ascontiguousarray([c.tolist() for c in self.coordinateMatrix], dtype=np.float64)
        g.coordinates = xyz.ctypes.data
        return g, jns, layers, nl, xyz

    def runSimulation(self, totalTime: float, timeStep: float = 1e-13, writeFrequency: float = 1e-11) -> None:
        """``GroupInteraction::runSimulation`` (core/reservoir.hpp:185-226) on the GPU."""
        be = _abi.backend()
        g, jns, layers, nl, xyz = self._pack()
        a = _coupled.run_args(totalTime, timeStep, writeFrequency, 1)
        plan = be.group_plan(g, jns, layers, a)
        if plan.solver_overridden:
            print("[WARNING] Solver automatically changed to Euler Heun for stochastic calculation.")
        n, ns = self.noElements, int(plan.n_samples)
        m = _coupled.initial_state(self.junctionList, 1)
        jlog = np.zeros((int(plan.n_junction_channels), max(ns, 1), n), dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        rc = be.group_run_host(g, jns, layers, None, 0, a, m, jlog, status)
        times = _coupled.sample_times(plan, timeStep)[:ns]
        if rc in (_abi.OK, _abi.ENAN):
            _coupled.write_back(self.junctionList, m, jlog[:, :ns, :], times)
        be.check(rc)

    def run_batch(
        self,
        n_instances: int,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        overrides: Optional[Dict[tuple, np.ndarray]] = None,
        m0: Optional[np.ndarray] = None,
        seed: Optional[int] = None,
        log: bool = True,
        device: int = 0,
    ) -> Dict[str, np.ndarray]:
        """``n_instances`` copies of this array in one call; ``overrides`` maps (layer_id | "all",
        slot name, field name) -> [n_instances, N].  Returns ``"<name>"`` -> [n_samples,
        n_instances, N] per junction channel (junction 0's channel names) plus ``m_state``."""
        from .batch import _SLOT_BY_NAME

        be = _abi.backend()
        g, jns, layers, nl, xyz = self._pack()
        B, n = int(n_instances), self.noElements
        a = _coupled.run_args(totalTime, timeStep, writeFrequency, B, log=log, seed=seed)
        plan = be.group_plan(g, jns, layers, a)
        ns = int(plan.n_samples)
        fidx = {nm: i for i, nm in enumerate(_abi.FIELD_NAMES)}
        ids = self.junctionList[0].getLayerIds()
        ovr: Dict[tuple, np.ndarray] = {}
        for (lid, slot, field), vals in (overrides or {}).items():
            v = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(B * n))
            li = -1 if lid == "all" else ids.index(lid)
            ovr[(li, _SLOT_BY_NAME[slot] if isinstance(slot, str) else int(slot), fidx[field])] = v
        tab, n_ovr = _coupled.override_table(ovr)
        m = _coupled.initial_state(self.junctionList, B) if m0 is None else np.ascontiguousarray(m0, dtype=np.float64)
        jlog = np.zeros((int(plan.n_junction_channels), max(ns, 1), B * n), dtype=np.float64) if log else None
        status = np.zeros(B * n, dtype=np.int32)
        be.check(be.group_run_host(g, jns, layers, tab, n_ovr, a, m, jlog, status, device))
        out: Dict[str, np.ndarray] = {"time": _coupled.sample_times(plan, timeStep)[:ns], "m_state": m}
        if log:
            jl = jlog[:, :ns, :].reshape(jlog.shape[0], ns, B, n)
            for ch, name in enumerate(self.junctionList[0]._channel_names()):
                out[name] = jl[ch]
        return out
