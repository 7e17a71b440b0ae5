"""Device replacement of ``cmtj.utils.parallel.distribute`` (cmtj/utils/parallel.py:13-50).

The reference maps a Python ``simulation_fn`` over the Cartesian product of parameter lists with a
``multiprocess.Pool`` and yields ``(index, output)``.  Here the product becomes the instance axis of
ONE batch: every parameter list is bound to a driver field of the junction and the whole grid integrates
in a single fused launch.  Inside an initialised ``torch.distributed`` job the grid is sharded contiguously
over the ranks (:class:`~cmtj_b200.sweep.ShardedSweep`, one GPU per rank, no collective on the data path);
every rank then yields the ``(index, output)`` pairs of ITS shard."""
from __future__ import annotations

from itertools import product
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from ..core import Junction, SolverMode
from ..sweep import ShardedSweep


def sweep_product(spaces: Sequence[Sequence[float]]) -> Tuple[np.ndarray, np.ndarray]:
    """Cartesian product in the reference's order (itertools.product: last space fastest).
    Returns (indexes [N, n_spaces] int, values [N, n_spaces])."""
    spaces = [np.asarray(s, dtype=np.float64) for s in spaces]
    idx = np.array(list(product(*[range(len(s)) for s in spaces])), dtype=np.int64).reshape(-1, len(spaces))
    vals = np.stack([s[idx[:, k]] for k, s in enumerate(spaces)], axis=1) if len(spaces) else np.zeros((1, 0))
    return idx, vals


def distribute(
    junction: Junction,
    bindings: Sequence[Tuple[str, str, str]],
    spaces: Sequence[Sequence[float]],
    totalTime: float,
    timeStep: float = 1e-13,
    writeFrequency: float = 1e-11,
    solverMode: SolverMode = SolverMode.RK4,
    reduce_fn: Optional[Callable[[Dict[str, np.ndarray]], object]] = None,
    shuffle: bool = False,
    device: Optional[int] = None,
    seed: Optional[int] = None,
) -> Iterator[Tuple[List[int], object]]:
    """Integrate ``junction`` at every point of ``product(*spaces)`` in one batched launch.

    bindings[k] = (layer_id, slot, field) names the driver field that spaces[k] sweeps, e.g.
    ("all", "Hext_x", "constantValue") or ("free", "current", "amplitude").  Yields
    ``(index, output)`` like the reference, where ``index`` is the position of the point in each
    space and ``output = reduce_fn(log_of_that_instance)`` (default: the instance's log dict).
    ``device`` defaults to LOCAL_RANK (0 outside a distributed job); ``seed`` must be the same on every rank
    for the shards to be the pieces of one sweep (Philox streams are keyed by the global instance index)."""
    if len(bindings) != len(spaces):
        raise ValueError("one binding (layer, slot, field) per parameter space is required")
    idx, vals = sweep_product(spaces)
    sweep = ShardedSweep(junction, idx.shape[0], device=device)
    for k, (layer, slot, field) in enumerate(bindings):
        sweep.set_override(layer, slot, field, vals[:, k])
    log = sweep.batch.run(totalTime, timeStep, writeFrequency, solverMode, seed=seed)
    order = np.arange(sweep.lo, sweep.hi)
    if shuffle:
        np.random.shuffle(order)
    for i in order:
        inst = log.instance(int(i) - sweep.lo)
        yield [int(v) for v in idx[i]], (reduce_fn(inst) if reduce_fn else inst)
