"""Device replacement of ``cmtj.utils.parallel.distribute`` (cmtj/utils/parallel.py:13-50).

The reference maps a Python ``simulation_fn`` over the Cartesian product of parameter lists with a
``multiprocess.Pool`` and yields ``(index, output)``.  Here the product becomes the instance axis of
ONE :class:`~cmtj_b200.JunctionBatch`: every parameter list is bound to a driver field (or the
magnetisation) of the junction, the whole grid integrates in a single fused launch (sharded over the
ranks of a ``torch.distributed`` job, if any), and the same ``(index, output)`` pairs come back."""
from __future__ import annotations

from itertools import product
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from ..batch import JunctionBatch
from ..core import Junction, SolverMode


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
    device: int = 0,
    seed: Optional[int] = None,
) -> Iterator[Tuple[List[int], object]]:
    """Integrate ``junction`` at every point of ``product(*spaces)`` in one batched launch.

    bindings[k] = (layer_id, slot, field) names the driver field that spaces[k] sweeps, e.g.
    ("all", "Hext_x", "constantValue") or ("free", "current", "amplitude").  Yields
    ``(index, output)`` like the reference, where ``index`` is the position of the point in each
    space and ``output = reduce_fn(log_of_that_instance)`` (default: the instance's log dict)."""
    if len(bindings) != len(spaces):
        raise ValueError("one binding (layer, slot, field) per parameter space is required")
    idx, vals = sweep_product(spaces)
    order = np.arange(idx.shape[0])
    if shuffle:
        np.random.shuffle(order)
    batch = JunctionBatch(junction, idx.shape[0], device=device)
    for k, (layer, slot, field) in enumerate(bindings):
        batch.set_override(layer, slot, field, vals[:, k])
    log = batch.run(totalTime, timeStep, writeFrequency, solverMode, seed=seed)
    for i in order:
        inst = log.instance(int(i))
        yield [int(v) for v in idx[i]], (reduce_fn(inst) if reduce_fn else inst)
