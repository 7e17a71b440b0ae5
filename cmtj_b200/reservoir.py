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


class _Shard:
    """This rank's slice of a sharded array and the sizes of its buffers."""

    def __init__(self, n_elements: int, rank: int, world: int, n_layers: int, n_channels: int, n_samples: int):
        self.N, self.rank, self.world = int(n_elements), int(rank), int(world)
        self.nl, self.ch, self.ns = int(n_layers), int(n_channels), int(n_samples)
        self.per = -(-self.N // self.world)  # ceil(N / world): contiguous ranges, the last ones may be short or empty
        self.lo = min(self.N, self.rank * self.per)
        self.hi = min(self.N, self.lo + self.per)
        self.own = self.hi - self.lo

    def range_of(self, r: int):
        lo = min(self.N, r * self.per)
        return lo, min(self.N, lo + self.per)


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
        self._custom_interaction: Optional[Callable] = None
        self._pair_table: Optional[np.ndarray] = None

    def setInteractionFunction(self, interactionFunc: Callable) -> None:
        """``GroupInteraction::setInteractionFunction`` (core/reservoir.hpp:181, python/cmtj.cpp:438).  The three built-in
        functions run as device code.  Any other callable ``f(r1, r2, layer1, layer2) -> CVector`` is accepted if it is
        AFFINE in the partner's magnetisation ``layer2.mag`` and does not depend on ``layer1.mag`` (every linear-response
        coupling is): it is probed once per ordered pair at run time -- ``layer2.mag`` = 0, e_x, e_y, e_z -- and the device
        evaluates the resulting 3x3 tensor + offset, instead of the reference's N^2 Python calls per step.  A callable that
        fails the affinity check at a random magnetisation raises ``NotImplementedError`` when the simulation starts."""
        if interactionFunc in _INTERACTIONS:
            self._interaction, self._custom_interaction = _INTERACTIONS[interactionFunc], None
        elif callable(interactionFunc):
            self._interaction, self._custom_interaction = _abi.INT_TABLE, interactionFunc
        else:
            raise TypeError("interactionFunc must be callable")

    def _probe_interaction(self) -> np.ndarray:
        """[N b][12][N a] table of the custom interaction: T[a][b] (rows 0..8, row-major) and c[a][b] (rows 9..11)."""
        f, n = self._custom_interaction, self.noElements
        tops = [j.getLayer(self.topId) for j in self.junctionList]
        table = np.zeros((n, 12, n), dtype=np.float64)
        rng = np.random.default_rng(12345)
        basis = [CVector(0, 0, 0), CVector(1, 0, 0), CVector(0, 1, 0), CVector(0, 0, 1)]

        def call(a, b, mag_b, mag_a=None):
            l1, l2 = tops[a]._clone(), tops[b]._clone()
            l2.mag = mag_b
            if mag_a is not None:
                l1.mag = mag_a
            return np.array(CVector._coerce(f(self.coordinateMatrix[a].copy(), self.coordinateMatrix[b].copy(), l1, l2)).tolist())

        worst, scale = 0.0, 0.0
        for a in range(n):
            for b in range(n):
                if a == b:
                    continue
                c, ex, ey, ez = (call(a, b, m.copy()) for m in basis)
                T = np.stack([ex - c, ey - c, ez - c], axis=1)  # columns = responses to the unit vectors
                table[b, 0:9, a] = T.reshape(9)
                table[b, 9:12, a] = c
                if (a + b) % max(1, n // 4) == 0 or n <= 8:  # affinity / independence of layer1.mag, on a sample of the pairs
                    v = rng.normal(size=3)
                    v /= np.linalg.norm(v)
                    w = rng.normal(size=3)
                    w /= np.linalg.norm(w)
                    got = call(a, b, CVector(*v), CVector(*w))
                    worst = max(worst, float(np.abs(got - (c + T @ v)).max()))
                    scale = max(scale, float(np.abs(got).max()), float(np.abs(c).max()))
        if worst > 1e-9 * max(scale, 1e-300):
            raise NotImplementedError(
                "the interaction function is not affine in layer2.mag (or depends on layer1.mag): only such functions, and "
                "nullDipoleInteraction / computeDipoleInteraction / computeDipoleInteractionNoumra, run on the device"
            )
        return table

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
        self._pair_table = None
        if self._interaction == _abi.INT_TABLE:
            self._pair_table = np.ascontiguousarray(self._probe_interaction())  # kept alive on the object for the call
            g.pair_table = self._pair_table.ctypes.data
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

    def run_sharded(
        self,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        rank: Optional[int] = None,
        world: Optional[int] = None,
        device: Optional[int] = None,
        seed: Optional[int] = None,
        exchange: Optional[Callable] = None,
        gather: bool = True,
        transport: str = "auto",
        _host_memory: bool = False,
        _peer_areas: Optional[Callable] = None,
    ) -> Dict[str, np.ndarray]:
        """``runSimulation`` for an array split over the ranks of a ``torch.distributed`` job, one process per GPU
        (core/reservoir.hpp:129-152 with the element loop cut into contiguous ranges of ceil(N / world)).  Every rank
        constructs the SAME GroupInteraction and calls this; rank r integrates its own elements and the top-layer
        magnetisations are all-gathered once per step (NCCL, in place on the device, stream-ordered: no host
        synchronisation inside the time loop).  The pair sums are accumulated in an order that depends on N alone, so
        the trajectories are bit-identical for every world size.

        ``transport``: ``"peer"`` -- the step kernels store their slice straight into every rank's exchange area over
        NVLink (CUDA IPC mappings, set up here through ``torch.distributed``) and hand-shake through arrival counters: no
        collective, no host in the time loop; ``"nccl"`` -- the all-gather above; ``"auto"`` -- peer memory when the job
        runs NCCL on one node with at most 16 ranks, else the collective.
        ``exchange(buf_address, count, step, stream_address) -> 0`` replaces the collective (tests, other transports).
        With ``gather`` every rank ends up with all junction logs, as after the reference's runSimulation; without it
        only this rank's junctions are written back.  Returns this rank's ``{"lo", "hi", "m_state", "log"}``."""
        import os

        from .sweep import _dist

        be = _abi.backend()
        dist = _dist()
        rank = int(rank if rank is not None else (dist.get_rank() if dist else int(os.environ.get("RANK", 0))))
        world = int(world if world is not None else (dist.get_world_size() if dist else int(os.environ.get("WORLD_SIZE", 1))))
        dev = int(device if device is not None else os.environ.get("LOCAL_RANK", 0))
        if seed is None:
            box = [_coupled._next_seed() if rank == 0 else None]
            if dist is not None and world > 1:
                dist.broadcast_object_list(box, src=0)  # one Philox key for the whole array
            seed = box[0] if box[0] is not None else 0
        g, jns, layers, nl, xyz = self._pack()
        a = _coupled.run_args(totalTime, timeStep, writeFrequency, 1, seed=seed)
        plan = be.group_plan(g, jns, layers, a)
        if plan.solver_overridden and rank == 0:
            print("[WARNING] Solver automatically changed to Euler Heun for stochastic calculation.")
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
        use_peer = (transport == "peer") or (
            transport == "auto" and exchange is None and not _host_memory and world > 1 and world <= 16
            and dist is not None and dist.get_backend() == "nccl" and _single_node()
        )
        if use_peer and (_host_memory or exchange is not None):
            raise ValueError("the peer-memory transport takes no exchange callback")
        sh = _Shard(self.noElements, rank, world, nl, int(plan.n_junction_channels), int(plan.n_samples))
        N, ns, ch, own, lo, hi = sh.N, sh.ns, sh.ch, sh.own, sh.lo, sh.hi
        m_host = np.ascontiguousarray(_coupled.initial_state(self.junctionList, 1)[:, lo:hi])
        if _host_memory:  # a host-memory stand-in for the library (CPU tests of this plumbing)
            torch = None
            m_dev, jlog_dev = m_host, np.zeros((ch, max(ns, 1), max(own, 1)), dtype=np.float64)
            st_dev = np.zeros(max(own, 1), dtype=np.int32)
            addr = lambda t: t.ctypes.data  # noqa: E731
            stream_ptr = 0
        else:
            import torch

            tdev = torch.device("cuda", dev)
            stream = torch.cuda.Stream(tdev)
            stream_ptr = stream.cuda_stream
            with torch.cuda.stream(stream):
                m_dev = torch.from_numpy(m_host if own > 0 else np.zeros((m_host.shape[0], 1))).to(tdev)  # never a null pointer
                jlog_dev = torch.zeros((ch, max(ns, 1), max(own, 1)), dtype=torch.float64, device=tdev)
                st_dev = torch.zeros(max(own, 1), dtype=torch.int32, device=tdev)
            addr = lambda t: t.data_ptr()  # noqa: E731
        if use_peer:
            peers = (_peer_areas or _ipc_peer_areas)(be, dist, rank, world, dev, N)
            try:
                rc = be.group_run_sharded_peer(g, jns, layers, a, rank, world, peers.areas, addr(m_dev), addr(jlog_dev), addr(st_dev), dev, stream_ptr)
                stream.synchronize()
            finally:
                peers.close()  # barrier, then unmap and free
            return self._finish_sharded(be, rc, dist, sh, dev, m_dev, jlog_dev, plan, timeStep, gather, torch)
        failure: List[BaseException] = []
        if exchange is None:
            exchange = _collective_exchange(dist, rank, world, None if _host_memory else dev)

        def _cb(ctx, buf, count, step, stream_arg):
            try:
                return int(exchange(buf, int(count), int(step), stream_arg or 0) or 0)
            except BaseException as e:  # an exception must not unwind through the C frames
                failure.append(e)
                return 1

        cb = _abi.EXCHANGE_FN(_cb)
        rc = be.group_run_sharded(g, jns, layers, a, rank, world, cb, addr(m_dev), addr(jlog_dev), addr(st_dev), dev, stream_ptr)
        if failure:
            raise failure[0]
        return self._finish_sharded(be, rc, dist, sh, dev, m_dev, jlog_dev, plan, timeStep, gather, torch, stream if torch is not None else None)

    def _finish_sharded(self, be, rc, dist, sh: _Shard, dev, m_dev, jlog_dev, plan, timeStep, gather, torch, stream=None):
        """results of this rank to the host, the optional gather of every rank's logs, write-back into the junction objects"""
        world, nl, ch, ns, N, per, lo, hi, own = sh.world, sh.nl, sh.ch, sh.ns, sh.N, sh.per, sh.lo, sh.hi, sh.own
        if torch is not None and stream is not None:
            stream.synchronize()
        if torch is not None:
            m_own, jlog_own = m_dev.cpu().numpy()[:, :own], jlog_dev.cpu().numpy()[:, :ns, :own]
        else:
            m_own, jlog_own = m_dev, jlog_dev[:, :ns, :own]
        times = _coupled.sample_times(plan, timeStep)[:ns]
        ok = rc in (_abi.OK, _abi.ENAN)
        if ok and gather and world > 1:
            if dist is None:
                raise RuntimeError("gather=True needs an initialised torch.distributed job")
            import torch as _t

            rows = np.zeros((3 * nl + ch * ns, per), dtype=np.float64)
            rows[: 3 * nl, :own] = m_own
            rows[3 * nl :, :own] = jlog_own.reshape(ch * ns, own)
            t = _t.from_numpy(rows)
            if dist.get_backend() == "nccl":
                t = t.to(_t.device("cuda", dev))
            parts = [_t.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            for r, part in enumerate(parts):
                r_lo, r_hi = sh.range_of(r)
                if r_hi > r_lo:
                    pr = part.cpu().numpy()[:, : r_hi - r_lo]
                    _coupled.write_back(self.junctionList[r_lo:r_hi], pr[: 3 * nl], pr[3 * nl :].reshape(ch, ns, r_hi - r_lo), times)
        elif ok and own > 0:
            _coupled.write_back(self.junctionList[lo:hi], m_own, jlog_own, times)
        be.check(rc)
        return {"lo": lo, "hi": hi, "m_state": m_own, "log": jlog_own, "time": times}

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


class _DeviceSpan:
    """A raw device address as a ``__cuda_array_interface__`` object (float64, 1-D), so that torch can view it."""

    def __init__(self, address: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(address), False), "version": 2}


def _collective_exchange(dist, rank: int, world: int, device: Optional[int]):
    """The default ``cmtjb_exchange_fn``: an in-place all-gather of the [world][count] exchange buffer over
    ``torch.distributed`` -- NCCL on the library's stream for device buffers (``device`` given), the job's CPU backend
    for host buffers.  A single rank has nothing to exchange."""
    if world == 1:
        return lambda buf, count, step, stream: 0
    if dist is None:
        raise RuntimeError("a sharded array over more than one rank needs an initialised torch.distributed job")
    import ctypes as C

    import torch

    if device is None:

        def host_exchange(buf, count, step, stream):
            full = torch.from_numpy(np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(world * count,)))
            parts = [torch.empty(count, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, full[rank * count : (rank + 1) * count].clone())
            for r, p in enumerate(parts):
                full[r * count : (r + 1) * count] = p
            return 0

        return host_exchange

    tdev = torch.device("cuda", device)
    cache: Dict[tuple, tuple] = {}

    def device_exchange(buf, count, step, stream):
        key = (buf, count, stream)
        if key not in cache:
            full = torch.as_tensor(_DeviceSpan(buf, world * count), device=tdev)
            cache[key] = (full, full[rank * count : (rank + 1) * count], torch.cuda.ExternalStream(stream, device=tdev))
        full, mine, ext = cache[key]
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(full, mine)  # in place: NCCL's sendbuff == recvbuff + rank * count case
        return 0

    return device_exchange


def _single_node() -> bool:
    import os

    return int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0) == int(os.environ.get("WORLD_SIZE", "-1") or -1)


class _PeerAreas:
    """The exchange areas of a peer-memory run as seen from this process: ``areas[r]`` for every rank, own one included."""

    def __init__(self, be, device, own_ptr, areas, opened, barrier):
        self.be, self.device, self.own_ptr, self.areas, self.opened, self.barrier = be, device, own_ptr, areas, opened, barrier

    def close(self):
        import ctypes as C

        self.barrier()  # nobody unmaps or frees while a peer may still store into an area
        lib = self.be.lib
        for p in self.opened:
            lib.cmtjb_peer_close(C.c_int(self.device), C.c_void_p(p))
        self.barrier()
        lib.cmtjb_peer_free(C.c_int(self.device), C.c_void_p(self.own_ptr))


def _ipc_peer_areas(be, dist, rank: int, world: int, device: int, n_elements: int) -> _PeerAreas:
    """Allocate this rank's exchange area, trade CUDA IPC handles through ``torch.distributed`` and map the peers' areas."""
    import ctypes as C

    lib = be.lib
    lib.cmtjb_group_peer_area_bytes.restype = C.c_int64
    nbytes = int(lib.cmtjb_group_peer_area_bytes(C.c_int32(n_elements), C.c_int32(world)))
    own = C.c_void_p()
    be.check(lib.cmtjb_peer_alloc(C.c_int(device), C.c_int64(nbytes), C.byref(own)))
    handle = (C.c_ubyte * 64)()
    be.check(lib.cmtjb_peer_export(C.c_int(device), own, handle))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))  # also the barrier after which every area exists and is zero
    areas, opened = [], []
    for r in range(world):
        if r == rank:
            areas.append(own.value)
            continue
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(handles[r])
        be.check(lib.cmtjb_peer_open(C.c_int(device), h, C.byref(p)))
        areas.append(p.value)
        opened.append(p.value)
    return _PeerAreas(be, device, own.value, areas, opened, dist.barrier)
