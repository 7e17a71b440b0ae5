"""Multi-GPU sweeps: one process per GPU, instances sharded contiguously, NO data-path
collective (sweep instances are independent units).  ``torch.distributed`` is used
only to learn rank / world size and for the optional host-side gather of results on rank 0.

The reference's analogue is ``cmtj.utils.parallel.distribute`` (cmtj/utils/parallel.py:13-50), a
``multiprocess.Pool`` over sweep points."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from .batch import BatchLog, JunctionBatch
from .core import Junction, SolverMode


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n instances for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


class ShardedSweep:
    """A :class:`JunctionBatch` over this rank's shard of a global sweep of ``n_instances``.

    Per-instance inputs are given for the GLOBAL sweep (every rank passes the same arrays) and
    sliced locally; thermal-noise streams are keyed by the global instance index, so the union of
    the shards is identical to one big batch on one GPU."""

    def __init__(self, junction: Junction, n_instances: int, rank: Optional[int] = None, world: Optional[int] = None,
                 device: Optional[int] = None):
        d = _dist()
        self.rank = int(rank if rank is not None else (d.get_rank() if d else int(os.environ.get("RANK", 0))))
        self.world = int(world if world is not None else (d.get_world_size() if d else int(os.environ.get("WORLD_SIZE", 1))))
        self.n_global = int(n_instances)
        self.lo, self.hi = shard_range(self.n_global, self.rank, self.world)
        dev = int(device if device is not None else os.environ.get("LOCAL_RANK", 0))
        self.batch = JunctionBatch(junction, self.hi - self.lo, device=dev, instance_offset=self.lo)

    def _local(self, values, width=None):
        v = np.asarray(values, dtype=np.float64)
        if v.shape[0] != self.n_global:
            raise ValueError(f"expected {self.n_global} global values, got {v.shape[0]}")
        return v[self.lo : self.hi]

    def set_external_field(self, layer_id: str, H) -> None:
        self.batch.set_external_field(layer_id, self._local(H))

    def set_magnetisation(self, layer_id: str, m) -> None:
        self.batch.set_magnetisation(layer_id, self._local(m))

    def set_override(self, layer_id: str, slot, field, values) -> None:
        self.batch.set_override(layer_id, slot, field, self._local(values))

    def run(self, totalTime: float, timeStep: float = 1e-13, writeFrequency: float = 1e-11,
            solverMode: SolverMode = SolverMode.RK4, log: bool = True, seed: int = 0, gather: bool = True):
        """Integrate the local shard.  With ``gather`` the per-shard logs are collected on rank 0
        (returns the global :class:`BatchLog` there and ``None`` elsewhere); without it every rank
        keeps its own shard's log (the scalable mode: results stay where they were produced)."""
        local = self.batch.run(totalTime, timeStep, writeFrequency, solverMode, log=log, seed=seed)
        if not gather or self.world == 1:
            return local
        return gather_logs(local, self.n_global, self.rank, self.world, self.batch.device)

    def gather_magnetisation(self, layer_id: str) -> Optional[np.ndarray]:
        m = self.batch.get_magnetisation(layer_id)
        if self.world == 1:
            return m
        parts = _gather_arrays(np.ascontiguousarray(m.T), self.n_global, self.rank, self.world, self.batch.device)
        return None if parts is None else parts.T


def _gather_arrays(a: np.ndarray, n_global: int, rank: int, world: int, device: Optional[int] = None) -> Optional[np.ndarray]:
    """Gather arrays [..., n_local] along the last axis onto rank 0 (shards padded to equal size).  With the NCCL
    backend the staging tensor lives on ``device`` (the sweep's own GPU), not on torch's current device."""
    import torch

    dist = _dist()
    if dist is None:
        raise RuntimeError("torch.distributed is not initialised")
    pad = shard_range(n_global, 0, world)[1]  # largest shard
    lead = a.shape[:-1]
    buf = np.zeros(lead + (pad,), dtype=np.float64)
    buf[..., : a.shape[-1]] = a
    cuda = dist.get_backend() == "nccl"
    t = torch.from_numpy(buf)
    if cuda:
        t = t.to(torch.device("cuda", int(device if device is not None else os.environ.get("LOCAL_RANK", 0))))
    outs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, outs, dst=0)
    if rank != 0:
        return None
    res = np.empty(lead + (n_global,), dtype=np.float64)
    for r in range(world):
        lo, hi = shard_range(n_global, r, world)
        res[..., lo:hi] = outs[r].cpu().numpy()[..., : hi - lo]
    return res


def gather_logs(local: BatchLog, n_global: int, rank: int, world: int, device: Optional[int] = None) -> Optional[BatchLog]:
    keys = [k for k in local if k != "time"]
    out = BatchLog() if rank == 0 else None
    if keys:
        stacked = _gather_arrays(np.stack([local[k] for k in keys]), n_global, rank, world, device)
        if rank == 0:
            for i, k in enumerate(keys):
                out[k] = stacked[i]
    if rank == 0 and "time" in local:
        out["time"] = local["time"]
    return out
