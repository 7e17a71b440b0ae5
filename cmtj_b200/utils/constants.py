"""Unit conversions and physical constants of ``cmtj.utils`` (cmtj/utils/constants.py): module-level
values read ONCE at import from :class:`cmtj_b200.constants.PhysicalConstants` (the same mutable
table the kernels snapshot at every run), plus the fixed conversion factors that table lacks."""
import math

from ..constants import PhysicalConstants as _PC

mu0 = _PC.magnetic_permeability()
gyromagnetic_ratio = _PC.gyromagnetic_ratio()
TtoAm = _PC.TtoAm()
AmtoT = 1.0 / TtoAm
hbar = _PC.hbar()
hplanck = hbar * 2 * math.pi
echarge = _PC.elementary_charge()
boltzmann = _PC.boltzmann_constant()
gamma = 28024e6  # Hz/T
gamma_rad = 1.76e11  # rad/(s T)
OetoAm = 79.57747
AmtoOe = 1.0 / OetoAm
me = 9.109e-31
bohr_magneton = abs(echarge) * hbar / (2 * me)
