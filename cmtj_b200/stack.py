"""``cmtj.stack``: electrically coupled junctions (``SeriesStack`` / ``ParallelStack``,
core/stack.hpp:16-480, bound at python/cmtj.cpp:358-426), integrated on the GPU by
``stack_kernel`` through ``cmtjb_stack_run_host``."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _abi, _coupled
from .core import AxialDriver, CVector, Junction, MR_CLASSIC, NullDriver, ScalarDriver, _as_driver

_FIELD_INDEX = {n: i for i, n in enumerate(_abi.FIELD_NAMES)}


class _Stack:
    _KIND = _abi.STACK_SERIES

    def __init__(
        self,
        junctionList: Sequence[Junction],
        topId_a: str = "free",
        bottomId: str = "bottom",
        phaseOffset: float = 0.0,
        useKCL: bool = True,
    ):
        junctionList = list(junctionList)
        if len(junctionList) < 2:
            raise RuntimeError("Stack must have at least 2 junctions!")
        self.junctionList: List[Junction] = [j._clone() for j in junctionList]
        if any(j.MR_mode != MR_CLASSIC for j in self.junctionList):
            raise RuntimeError(
                "Junction has a non-classic magnetoresitance mode! Define the junction with Rp and Rap resistance values."
            )
        self.topId = str(topId_a)
        self.bottomId = str(bottomId)
        self.phaseOffset = float(phaseOffset)
        self.useKCL = bool(useKCL)
        self.stackSize = len(self.junctionList)
        self.couplingStrength: List[float] = [0.0]
        self.delayed = True
        self.currentDriver: ScalarDriver = ScalarDriver()
        self.currentDriverSet = False
        self._log: Dict[str, list] = {}

    # -- configuration (stack.hpp:156-253) --------------------------------------------------------
    def isTwoLayerMemberStack(self) -> bool:
        for j in self.junctionList:
            if j.layerNo >= 3:
                raise RuntimeError("At least one junction has more than 2 layers! It's not supported now.")
            if j.layerNo != 2:
                return False
        return True

    def setDelayed(self, delayed: bool) -> None:
        if not delayed and not self.isTwoLayerMemberStack():
            raise RuntimeError("Non delayed coupling is only supported for 2 layer stacks!")
        self.delayed = bool(delayed)

    def setMagnetisation(self, junction: int, layerId: str, mag) -> None:
        self.junctionList[_coupled._uint(junction, "setMagnetisation")].setLayerMagnetisation(layerId, mag)

    def getMagnetisation(self, junction: int, layerId: str) -> CVector:
        return self.junctionList[_coupled._uint(junction, "getMagnetisation")].getLayerMagnetisation(layerId)

    def getJunction(self, junctionId: int) -> Junction:
        k = _coupled._uint(junctionId, "getJunction")
        if k >= self.stackSize:
            raise IndexError("vector::_M_range_check")
        return self.junctionList[k]

    def setJunctionAnisotropyDriver(self, junctionId: int, layerId: str, k: ScalarDriver) -> None:
        self.junctionList[_coupled._uint(junctionId, "setJunctionAnisotropyDriver")].setLayerAnisotropyDriver(layerId, k)

    def setOerstedFieldDriver(self, driver: AxialDriver) -> None:
        for j in self.junctionList:
            j.setLayerOerstedFieldDriver("all", driver)

    def setExternalFieldDriver(self, driver: AxialDriver) -> None:
        for j in self.junctionList:
            j.setLayerExternalFieldDriver("all", driver)

    def resetCoupledCurrentDriver(self) -> None:
        self.currentDriver = ScalarDriver()
        for j in self.junctionList:
            j.setLayerCurrentDriver("all", self.currentDriver)
        self.currentDriverSet = False

    def setCoupledCurrentDriver(self, driver: ScalarDriver) -> None:
        self.currentDriver = _as_driver(driver)
        for j in self.junctionList:
            j.setLayerCurrentDriver("all", self.currentDriver)
        self.currentDriverSet = True

    def setCouplingStrength(self, coupling: Union[float, Sequence[float]]) -> None:
        if isinstance(coupling, (list, tuple, np.ndarray)):
            if len(coupling) != self.stackSize - 1:
                raise RuntimeError("Coupling strength vector must have size of stack size - 1!")
            self.couplingStrength = [float(c) for c in coupling]
        else:
            self.couplingStrength = [float(coupling)]

    def getCoupling(self, order: int) -> float:
        if not self.couplingStrength:
            raise RuntimeError("Coupling strength is not set!")
        if len(self.couplingStrength) == 1:
            return self.couplingStrength[0]
        return self.couplingStrength[order]

    def getPolarisationVector(self) -> CVector:
        probe = self.junctionList[0].getLayer(self.topId).referenceLayer
        for j in self.junctionList[1:]:
            if probe != j.getLayer(self.topId).referenceLayer:
                raise RuntimeError("Polarisation vectors are not equal in stack")
        if not probe.length():
            raise RuntimeError("Polarisation is not set!")
        return probe.copy()

    # -- logs (stack.hpp:222-283) ------------------------------------------------------------------
    def clearLogs(self) -> None:
        for j in self.junctionList:
            j.clearLog()
        self._log = {}

    def getLog(self, id: Optional[int] = None) -> Dict[str, List[float]]:
        if id is None:
            return {k: list(v) for k, v in self._log.items()}
        k = _coupled._uint(id, "getLog")
        if k < self.stackSize:
            return self.junctionList[k].getLog()
        raise RuntimeError("Asking for id of a non-existing junction!")

    def saveLogs(self, fileSave: str) -> None:
        if fileSave == "":
            print("Ignoring file save to an empty filename")
            return
        keys = list(self._log.keys())
        with open(fileSave, "w") as f:
            f.write("".join(k + ";" for k in keys) + "\n")
            for i in range(len(self._log.get("time", []))):
                f.write("".join(f"{self._log[k][i]:g};" for k in keys) + "\n")

    # -- packing ----------------------------------------------------------------------------------------
    def _pack(self):
        two = self.isTwoLayerMemberStack()
        jns, layers, nl = _coupled.pack_junctions(self.junctionList)
        st = _abi.Stack()
        st.n_junctions = self.stackSize
        st.kind = self._KIND
        st.use_kcl = int(self.useKCL)
        st.delayed = int(self.delayed)
        ids = self.junctionList[0].getLayerIds()
        for j in self.junctionList:
            j.getLayer(self.topId)  # raises like getLayerMagnetisation(topId) would
            if j.getLayerIds() != ids:
                raise NotImplementedError("junctions of a stack must use the same layer ids in the same order in this build")
        st.top_layer = ids.index(self.topId)
        if two:
            self.junctionList[0].getLayer(self.bottomId)
            st.bottom_layer = ids.index(self.bottomId)
        else:
            st.bottom_layer = -1
            pol = self.getPolarisationVector()
            st.polarisation[:] = pol.tolist()
        st.phaseOffset = self.phaseOffset
        coup = np.array([self.getCoupling(k) for k in range(self.stackSize - 1)], dtype=np.float64)
        st.coupling = coup.ctypes.data
        self.currentDriver._pack(st.currentDriver)
        return st, jns, layers, nl, coup

    def _channel_names(self) -> List[str]:
        n = self.stackSize
        return ["Resistance"] + [f"I_{j}" for j in range(n)] + [f"C_{j}_{j + 1}" for j in range(n - 1)]

    # -- the hot path entry -------------------------------------------------------------------------------
    def runSimulation(self, totalTime: float, timeStep: float = 1e-13, writeFrequency: float = 1e-11) -> None:
        """``Stack::runSimulation`` (core/stack.hpp:311-423) on the GPU."""
        be = _abi.backend()
        st, jns, layers, nl, coup = self._pack()
        a = _coupled.run_args(totalTime, timeStep, writeFrequency, 1)
        plan = be.stack_plan(st, jns, layers, a)
        if plan.solver_overridden:
            print("[WARNING] Solver automatically changed to Euler Heun for stochastic calculation.")
        n, ns = self.stackSize, int(plan.n_samples)
        m = _coupled.initial_state(self.junctionList, 1)
        jlog = np.zeros((int(plan.n_junction_channels), max(ns, 1), n), dtype=np.float64)
        slog = np.zeros((int(plan.n_system_channels), max(ns, 1), 1), dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        rc = be.stack_run_host(st, jns, layers, None, 0, a, m, jlog, slog, status)
        times = _coupled.sample_times(plan, timeStep)[:ns]
        if rc in (_abi.OK, _abi.ENAN):
            _coupled.write_back(self.junctionList, m, jlog[:, :ns, :], times)
            for name, row in zip(self._channel_names(), slog[:, :ns, 0]):
                self._log.setdefault(name, []).extend(row.tolist())
            self._log.setdefault("time", []).extend(times.tolist())
        be.check(rc)

    def run_batch(
        self,
        n_instances: int,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        current: Optional[Dict[str, Sequence[float]]] = None,
        overrides: Optional[Dict[tuple, np.ndarray]] = None,
        m0: Optional[np.ndarray] = None,
        seed: Optional[int] = None,
        device: int = 0,
    ) -> Dict[str, np.ndarray]:
        """``n_instances`` copies of this stack in one launch (a sweep over stack-level inputs).
        ``current`` maps a field name of the coupled current driver (``"amplitude"``,
        ``"frequency"``, ``"constantValue"`` ...) to one value per instance; ``overrides`` maps
        (layer_id | "all", slot name, field name) to an array [n_instances, N] of per-junction
        driver fields.  Returns stack channels as [n_samples, n_instances] and junction channels
        as ``"<j>/<name>"`` -> [n_samples, n_instances]; the stack object itself is not modified."""
        from .batch import _SLOT_BY_NAME

        be = _abi.backend()
        st, jns, layers, nl, coup = self._pack()
        B, n = int(n_instances), self.stackSize
        a = _coupled.run_args(totalTime, timeStep, writeFrequency, B, seed=seed)
        plan = be.stack_plan(st, jns, layers, a)
        ns = int(plan.n_samples)
        ovr: Dict[tuple, np.ndarray] = {}
        for name, vals in (current or {}).items():
            v = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(B))
            ovr[(_abi.LAYER_SYSTEM, 0, _FIELD_INDEX[name])] = v
        ids = self.junctionList[0].getLayerIds()
        for (lid, slot, field), vals in (overrides or {}).items():
            v = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(B * n))
            li = -1 if lid == "all" else ids.index(lid)
            ovr[(li, _SLOT_BY_NAME[slot] if isinstance(slot, str) else int(slot), _FIELD_INDEX[field])] = v
        tab, n_ovr = _coupled.override_table(ovr)
        m = _coupled.initial_state(self.junctionList, B) if m0 is None else np.ascontiguousarray(m0, dtype=np.float64)
        jlog = np.zeros((int(plan.n_junction_channels), max(ns, 1), B * n), dtype=np.float64)
        slog = np.zeros((int(plan.n_system_channels), max(ns, 1), B), dtype=np.float64)
        status = np.zeros(B * n, dtype=np.int32)
        be.check(be.stack_run_host(st, jns, layers, tab, n_ovr, a, m, jlog, slog, status, device))
        out: Dict[str, np.ndarray] = {"time": _coupled.sample_times(plan, timeStep)[:ns]}
        for name, row in zip(self._channel_names(), slog[:, :ns, :]):
            out[name] = row
        jl = jlog[:, :ns, :].reshape(jlog.shape[0], ns, B, n)
        for k, j in enumerate(self.junctionList):
            for ch, name in enumerate(j._channel_names()):
                out[f"{k}/{name}"] = jl[ch, :, :, k]
        out["m_state"] = m
        return out


class SeriesStack(_Stack):
    """core/stack.hpp:425-451: R = sum R_j, C_j = c_j (m_j.p_j + m_(j+1).p_(j+1)), phase j*offset."""

    _KIND = _abi.STACK_SERIES


class ParallelStack(_Stack):
    """core/stack.hpp:453-480: R = 1/sum(1/R_j), C_j = c_j (m_j.p_j - m_(j+1).p_(j+1)), phase offset."""

    _KIND = _abi.STACK_PARALLEL
