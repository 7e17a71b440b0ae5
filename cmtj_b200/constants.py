"""Mutable global physical constants: ``PhysicalConstants`` (core/constants.hpp:6-53,
python names cmtj.cpp:497-511).  A snapshot is taken at every run and travels to the
device in the junction header, so changing a constant between runs takes effect exactly
as it does in the reference ("Config / flags")."""
import math


class PhysicalConstants:
    _defaults = dict(
        MAGNETIC_PERMEABILITY=12.566e-7,
        GYRO=220880.0,
        TtoAm=795774.715459,
        HBAR=6.62607015e-34 / (2.0 * math.pi),
        ELECTRON_CHARGE=1.60217662e-19,
        BOLTZMANN_CONST=1.380649e-23,
    )
    _v = dict(_defaults)

    @classmethod
    def set_magnetic_permeability(cls, value: float) -> None:
        cls._v["MAGNETIC_PERMEABILITY"] = float(value)

    @classmethod
    def set_gyromagnetic_ratio(cls, value: float) -> None:
        cls._v["GYRO"] = float(value)

    @classmethod
    def set_TtoAm(cls, value: float) -> None:
        cls._v["TtoAm"] = float(value)

    @classmethod
    def set_hbar(cls, value: float) -> None:
        cls._v["HBAR"] = float(value)

    @classmethod
    def set_elementary_charge(cls, value: float) -> None:
        cls._v["ELECTRON_CHARGE"] = float(value)

    @classmethod
    def set_boltzmann_constant(cls, value: float) -> None:
        cls._v["BOLTZMANN_CONST"] = float(value)

    @classmethod
    def magnetic_permeability(cls) -> float:
        return cls._v["MAGNETIC_PERMEABILITY"]

    @classmethod
    def gyromagnetic_ratio(cls) -> float:
        return cls._v["GYRO"]

    @classmethod
    def TtoAm(cls) -> float:
        return cls._v["TtoAm"]

    @classmethod
    def hbar(cls) -> float:
        return cls._v["HBAR"]

    @classmethod
    def elementary_charge(cls) -> float:
        return cls._v["ELECTRON_CHARGE"]

    @classmethod
    def boltzmann_constant(cls) -> float:
        return cls._v["BOLTZMANN_CONST"]

    @classmethod
    def resetToDefaults(cls) -> None:
        cls._v = dict(cls._defaults)
