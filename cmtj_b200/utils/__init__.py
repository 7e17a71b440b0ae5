"""Callers of the junction-integration path: the sweep procedures of
``cmtj.utils.procedures`` and the resistance / spin-diode post-processing of
``cmtj.utils.resistance``, rebuilt so that a whole sweep is ONE batched device run and the
reductions happen on the GPU (fused into the time loop where the reference reduces a full-rate log).
Same function names and signatures as the reference's ``cmtj.utils``."""
from .constants import (  # noqa: F401
    AmtoOe,
    AmtoT,
    OetoAm,
    TtoAm,
    bohr_magneton,
    boltzmann,
    echarge,
    gamma,
    gamma_rad,
    gyromagnetic_ratio,
    hbar,
    hplanck,
    me,
    mu0,
)
from .filters import Filters  # noqa: F401
from .linear import FieldScan  # noqa: F401
from .parallel import distribute, sweep_product  # noqa: F401
from .procedures import PIMM_batch, PIMM_procedure, ResistanceParameters, VSD_batch, VSD_procedure, compute_spectrum_strip  # noqa: F401
from .resistance import (  # noqa: F401
    calculate_magnetoresistance,
    calculate_resistance_parallel,
    calculate_resistance_series,
    compute_gmr,
    compute_resistance,
    compute_sd,
)
