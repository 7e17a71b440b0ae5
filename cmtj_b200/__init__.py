"""cmtj_b200 -- B200-native (sm_100a) implementation of cmtj's junction-integration path.

The public names mirror the reference's ``cmtj`` package (``from _cmtj import *``,
cmtj/__init__.py:11-14) so scripts can switch with ``import cmtj_b200 as cmtj`` or
``cmtj_b200.install_as_cmtj()``.  Compute runs only on the GPU through the C-ABI library
``libcmtj_b200.so`` (see include/cmtj_b200.h); there is no CPU fallback.
"""
from . import constants, noise, reservoir, stack, utils  # noqa: F401
from .constants import PhysicalConstants  # noqa: F401
from .batch import BatchLog, JunctionBatch  # noqa: F401
from .sweep import ShardedSweep, shard_range  # noqa: F401
from .core import (  # noqa: F401
    AdaptiveIntegrationParams,
    AxialDriver,
    Axis,
    CVector,
    Junction,
    Layer,
    NullDriver,
    Reference,
    ScalarDriver,
    SolverMode,
    UpdateType,
    c_dot,
    constantDriver,
    gaussianImpulseDriver,
    gaussianStepDriver,
    posSineDriver,
    pulseDriver,
    seed,
    thermal_normals,
    sineDriver,
    stepDriver,
    trapezoidDriver,
)

# export_values() of the reference's enums (cmtj.cpp:129-176).  Later exports shadow earlier
# ones exactly as in pybind11: `none` ends up being Reference.none.
xaxis, yaxis, zaxis, all, none = Axis.xaxis, Axis.yaxis, Axis.zaxis, Axis.all, Axis.none
none, fixed, top, bottom = Reference.none, Reference.fixed, Reference.top, Reference.bottom
RK4, Heun, EulerHeun, DormandPrince = SolverMode.RK4, SolverMode.Heun, SolverMode.EulerHeun, SolverMode.DormandPrince
constant, pulse, sine, step, posine, halfsine, trapezoid, gaussimpulse, gaussstep, custom = (
    UpdateType.constant,
    UpdateType.pulse,
    UpdateType.sine,
    UpdateType.step,
    UpdateType.posine,
    UpdateType.halfsine,
    UpdateType.trapezoid,
    UpdateType.gaussimpulse,
    UpdateType.gaussstep,
    UpdateType.custom,
)

__version__ = "0.1.0"


def install_as_cmtj() -> None:
    """Make ``import cmtj`` (and ``cmtj.stack``, ``cmtj.reservoir``, ``cmtj.noise``,
    ``cmtj.constants``, ``_cmtj``) resolve to this package, so scripts written against the
    reference run unchanged (cmtj/__init__.py:11-14 re-exports ``_cmtj`` the same way)."""
    import sys

    me = sys.modules[__name__]
    sys.modules["cmtj"] = me
    sys.modules["_cmtj"] = me
    for sub in ("stack", "reservoir", "noise", "constants", "utils"):
        sys.modules["cmtj." + sub] = getattr(me, sub)
        sys.modules["_cmtj." + sub] = getattr(me, sub)
    for sub in ("procedures", "resistance", "linear", "filters", "parallel", "constants"):
        sys.modules["cmtj.utils." + sub] = getattr(me.utils, sub)
