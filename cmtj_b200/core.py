"""Host-side mirror of the reference's ``_cmtj`` module for the junction path.

Same names, argument meaning and error behaviour as ``python/cmtj.cpp`` so that scripts
written against ``cmtj`` run unchanged (``CVector`` :101-127, enums :129-176,
``ScalarDriver`` :178-206, ``NullDriver`` :208-211, ``AxialDriver`` :213-224, ``Layer``
:226-289, ``Junction`` :291-355).  These classes only *describe* the system; every
``runSimulation`` packs the description into the C-ABI records (``include/cmtj_b200.h``)
and runs it on the GPU through ``libcmtj_b200.so``.
"""
from __future__ import annotations

import copy
import enum
import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np

from . import _abi
from .constants import PhysicalConstants


# --------------------------------------------------------------------------------------
# enums (export_values() in the reference: members are also module-level names)
# --------------------------------------------------------------------------------------
class _CEnum(enum.Enum):
    def __int__(self) -> int:
        return int(self.value)

    def __index__(self) -> int:
        return int(self.value)


class Axis(_CEnum):
    """core/cvector.hpp:11"""

    xaxis = 0
    yaxis = 1
    zaxis = 2
    all = 3
    none = 4


class Reference(_CEnum):
    """core/junction.hpp:141 (python names: cmtj.cpp:137-142)"""

    none = 0
    fixed = 1
    top = 2
    bottom = 3


class SolverMode(_CEnum):
    """core/junction.hpp:143 (python names: cmtj.cpp:144-149)"""

    EulerHeun = 0
    RK4 = 1
    DormandPrince = 2
    Heun = 3


class UpdateType(_CEnum):
    """core/drivers.hpp:18-29"""

    constant = 0
    pulse = 1
    sine = 2
    step = 3
    posine = 4
    halfsine = 5
    trapezoid = 6
    gaussimpulse = 7
    gaussstep = 8
    custom = 100


# --------------------------------------------------------------------------------------
# CVector
# --------------------------------------------------------------------------------------
class CVector:
    """3-vector with the reference's semantics (core/cvector.hpp:13-220)."""

    __slots__ = ("x", "y", "z")

    def __init__(self, x: float = 0.0, y: float = 0.0, z: float = 0.0):
        self.x = float(x)
        self.y = float(y)
        self.z = float(z)

    # pybind: py::implicitly_convertible<std::list<double>, CVector>
    @staticmethod
    def _coerce(v) -> "CVector":
        if isinstance(v, CVector):
            return v
        if isinstance(v, (list, tuple, np.ndarray)):
            if len(v) != 3:
                raise RuntimeError("Failed to create vector -- passed list was not of len 3!")
            return CVector(v[0], v[1], v[2])
        raise TypeError(f"cannot convert {type(v).__name__} to CVector")

    def length(self) -> float:
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def normalize(self) -> None:
        mag = self.length()
        if mag != 0:
            self.x = self.x / mag
            self.y = self.y / mag
            self.z = self.z / mag

    def tolist(self) -> List[float]:
        return [self.x, self.y, self.z]

    def copy(self) -> "CVector":
        return CVector(self.x, self.y, self.z)

    def __add__(self, o: "CVector") -> "CVector":
        o = CVector._coerce(o)
        return CVector(self.x + o.x, self.y + o.y, self.z + o.z)

    def __iadd__(self, o: "CVector") -> "CVector":
        o = CVector._coerce(o)
        self.x += o.x
        self.y += o.y
        self.z += o.z
        return self

    def __sub__(self, o: "CVector") -> "CVector":
        o = CVector._coerce(o)
        return CVector(self.x - o.x, self.y - o.y, self.z - o.z)

    def __isub__(self, o: "CVector") -> "CVector":
        o = CVector._coerce(o)
        self.x -= o.x
        self.y -= o.y
        self.z -= o.z
        return self

    def __mul__(self, s: float) -> "CVector":
        if isinstance(s, CVector):
            return NotImplemented
        return CVector(self.x * s, self.y * s, self.z * s)

    def __rmul__(self, s: float) -> "CVector":
        return CVector(s * self.x, s * self.y, s * self.z)

    def __imul__(self, s: float) -> "CVector":
        self.x *= s
        self.y *= s
        self.z *= s
        return self

    def __eq__(self, o) -> bool:
        if not isinstance(o, CVector):
            return NotImplemented
        return self.x == o.x and self.y == o.y and self.z == o.z

    def __ne__(self, o) -> bool:
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    __hash__ = None  # mutable value type

    def __getitem__(self, i: int) -> float:
        if i == 0:
            return self.x
        if i == 1:
            return self.y
        return self.z  # the reference returns z for any other index (cvector.hpp:161-168)

    def __len__(self) -> int:
        return 3

    def __iter__(self):
        yield self.x
        yield self.y
        yield self.z

    def __repr__(self) -> str:
        return f"[x:{self.x:g}, y:{self.y:g}, z:{self.z:g}]"

    __str__ = __repr__

    @staticmethod
    def fromSpherical(theta: float, phi: float, r: float = 1.0) -> "CVector":
        return CVector(
            r * math.sin(theta) * math.cos(phi), r * math.sin(theta) * math.sin(phi), r * math.cos(theta)
        )


def c_dot(a: CVector, b: CVector) -> float:
    """core/junction.hpp:82-84"""
    a = CVector._coerce(a)
    b = CVector._coerce(b)
    return a.x * b.x + a.y * b.y + a.z * b.z


def _tensor(t) -> List[CVector]:
    rows = [CVector._coerce(r).copy() for r in t]
    if len(rows) != 3:
        raise RuntimeError("tensor must have 3 rows")
    return rows


# --------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------
class ScalarDriver:
    """core/drivers.hpp:65-354.  A plain description; evaluated inline on the device."""

    __slots__ = ("update",) + _abi.FIELD_NAMES + ("callback",)

    def __init__(
        self,
        update: UpdateType = UpdateType.constant,
        constantValue: float = 0.0,
        amplitude: float = 0.0,
        frequency: float = -1.0,
        phase: float = 0.0,
        period: float = -1.0,
        cycle: float = -1.0,
        timeStart: float = -1.0,
        timeStop: float = -1.0,
        edgeTime: float = -1.0,
        steadyTime: float = -1.0,
        callback: Optional[Callable[[float], float]] = None,
    ):
        self.update = update
        self.constantValue = float(constantValue)
        self.amplitude = float(amplitude)
        self.frequency = float(frequency)
        self.phase = float(phase)
        self.period = float(period)
        self.cycle = float(cycle)
        self.timeStart = float(timeStart)
        self.timeStop = float(timeStop)
        self.edgeTime = float(edgeTime)
        self.steadyTime = float(steadyTime)
        self.callback = callback
        if update == UpdateType.pulse and (period == -1 or cycle == -1):
            raise RuntimeError("Selected pulse train driver type but either period or cycle were not set")
        if update == UpdateType.sine and frequency == -1:
            raise RuntimeError("Selected sine driver type but frequency was not set")

    # -- factories (drivers.hpp:135-260) ------------------------------------------------
    @staticmethod
    def getConstantDriver(constantValue: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.constant, constantValue)

    @staticmethod
    def getPulseDriver(constantValue: float, amplitude: float, period: float, cycle: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.pulse, constantValue, amplitude, -1, -1, period, cycle)

    @staticmethod
    def getSineDriver(constantValue: float, amplitude: float, frequency: float, phase: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.sine, constantValue, amplitude, frequency, phase)

    @staticmethod
    def getPosSineDriver(constantValue: float, amplitude: float, frequency: float, phase: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.posine, constantValue, amplitude, frequency, phase)

    @staticmethod
    def getHalfSineDriver(constantValue: float, amplitude: float, frequency: float, phase: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.halfsine, constantValue, amplitude, frequency, phase)

    @staticmethod
    def getStepDriver(constantValue: float, amplitude: float, timeStart: float, timeStop: float) -> "ScalarDriver":
        if timeStop <= timeStart:
            raise RuntimeError("Start time cannot be later than stop time!")
        return ScalarDriver(UpdateType.step, constantValue, amplitude, -1, -1, -1, -1, timeStart, timeStop)

    @staticmethod
    def getTrapezoidDriver(
        constantValue: float, amplitude: float, timeStart: float, edgeTime: float, steadyTime: float
    ) -> "ScalarDriver":
        return ScalarDriver(
            UpdateType.trapezoid, constantValue, amplitude, -1, -1, -1, -1, timeStart, -1, edgeTime, steadyTime
        )

    @staticmethod
    def getGaussianImpulseDriver(constantValue: float, amplitude: float, t0: float, sigma: float) -> "ScalarDriver":
        return ScalarDriver(UpdateType.gaussimpulse, constantValue, amplitude, -1, -1, -1, -1, t0, -1, sigma)

    @staticmethod
    def getGaussianStepDriver(constantValue: float, amplitude: float, t0: float, sigma: float) -> "ScalarDriver":
        # The reference builds a *gaussimpulse* here (drivers.hpp:241-245); kept for parity.
        return ScalarDriver(UpdateType.gaussimpulse, constantValue, amplitude, -1, -1, -1, -1, t0, -1, sigma)

    @staticmethod
    def getCustomDriver(callback: Callable[[float], float]) -> "ScalarDriver":
        if not callback:
            raise RuntimeError("Callback function is not set")
        code = getattr(callback, "__code__", None)
        if code is None or code.co_argcount != 1:
            raise RuntimeError("Callback function must have one argument")
        return ScalarDriver(UpdateType.custom, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, callback)

    # -- host evaluation (API surface only; the integrator evaluates drivers on device) ---
    def getCurrentScalarValue(self, time: float) -> float:
        """core/drivers.hpp:262-305"""
        u = self.update
        ret = self.constantValue
        if u == UpdateType.pulse:
            n = int(time / self.period)
            dT = self.cycle * self.period
            nT = n * self.period
            if nT <= time <= (nT + dT):
                ret += self.amplitude
        elif u == UpdateType.sine:
            ret += self.amplitude * math.sin(2 * math.pi * time * self.frequency + self.phase)
        elif u == UpdateType.posine:
            ret += abs(self.amplitude * math.sin(2 * math.pi * time * self.frequency + self.phase))
        elif u == UpdateType.halfsine:
            tamp = self.amplitude * math.sin(2 * math.pi * time * self.frequency + self.phase)
            if tamp <= 0:
                ret += tamp
        elif u == UpdateType.step:
            if self.timeStart <= time <= self.timeStop:
                ret += self.amplitude
        elif u == UpdateType.trapezoid:
            t0, e, s, a = self.timeStart, self.edgeTime, self.steadyTime, self.amplitude
            if time < t0:
                v = 0.0
            elif time <= t0 + e:
                v = (a / e) * (time - t0)
            elif time <= t0 + e + s:
                v = a
            elif time <= t0 + 2 * e + s:
                v = a - (a / e) * (time - (t0 + e + s))
            else:
                v = 0.0
            ret += v
        elif u == UpdateType.gaussimpulse:
            ret += self.amplitude * math.exp(-((time - self.timeStart) ** 2) / (2 * self.edgeTime**2))
        elif u == UpdateType.gaussstep:
            ret += 0.5 * self.amplitude * (1 + math.erf((time - self.timeStart) / (math.sqrt(2) * self.edgeTime)))
        elif u == UpdateType.custom:
            try:
                return float(self.callback(time))
            except Exception as e:  # drivers.hpp:299-302
                raise RuntimeError("Error in Python callback") from e
        return ret

    def _clone(self) -> "ScalarDriver":
        d = ScalarDriver.__new__(ScalarDriver)
        for s in ScalarDriver.__slots__:
            setattr(d, s, getattr(self, s))
        return d

    # -- arithmetic (drivers.hpp:313-353): scales/shifts BOTH constantValue and amplitude --
    def __mul__(self, v: float) -> "ScalarDriver":
        d = self._clone()
        d.callback = None
        d.constantValue = self.constantValue * v
        d.amplitude = self.amplitude * v
        return d

    def __imul__(self, v: float) -> "ScalarDriver":
        self.constantValue *= v
        self.amplitude *= v
        return self

    def __add__(self, v: float) -> "ScalarDriver":
        d = self._clone()
        d.callback = None
        d.constantValue = self.constantValue + v
        d.amplitude = self.amplitude + v
        return d

    def __iadd__(self, v: float) -> "ScalarDriver":
        self.constantValue += v
        self.amplitude += v
        return self

    def phaseShift(self, phase: float) -> None:
        self.phase += phase

    def setConstantValue(self, v: float) -> None:
        self.constantValue = float(v)

    def _is_constant(self) -> bool:
        return self.update == UpdateType.constant

    def _pack(self, out: _abi.Driver) -> None:
        out.update = int(self.update)
        for n in _abi.FIELD_NAMES:
            setattr(out, n, getattr(self, n))

    def __repr__(self) -> str:
        return f"ScalarDriver({self.update.name}, const={self.constantValue:g}, amp={self.amplitude:g})"


class NullDriver(ScalarDriver):
    """core/drivers.hpp:356-360"""

    def __init__(self):
        super().__init__()

    def getCurrentScalarValue(self, time: float) -> float:
        return 0.0


def _as_driver(d) -> ScalarDriver:
    if not isinstance(d, ScalarDriver):
        raise TypeError(f"expected ScalarDriver, got {type(d).__name__}")
    return d._clone() if type(d) is ScalarDriver else _sliced(d)


def _sliced(d: ScalarDriver) -> ScalarDriver:
    # Storing a NullDriver in a ScalarDriver slot slices it to a constant-0 ScalarDriver.
    s = ScalarDriver()
    for n in ScalarDriver.__slots__:
        setattr(s, n, getattr(d, n))
    return s


class AxialDriver:
    """core/drivers.hpp:362-461"""

    def __init__(self, *args):
        if len(args) == 0:
            self.drivers = [ScalarDriver(), ScalarDriver(), ScalarDriver()]
        elif len(args) == 3 and all(isinstance(a, ScalarDriver) for a in args):
            self.drivers = [_as_driver(a) for a in args]
        elif len(args) == 3:
            self.drivers = [ScalarDriver.getConstantDriver(float(a)) for a in args]
        elif len(args) == 1 and isinstance(args[0], CVector):
            v = args[0]
            self.drivers = [ScalarDriver.getConstantDriver(c) for c in (v.x, v.y, v.z)]
        elif len(args) == 1 and isinstance(args[0], (list, tuple)):
            lst = args[0]
            if len(lst) != 3:
                raise RuntimeError("The axial driver can only have 3 axes!")
            if all(isinstance(a, ScalarDriver) for a in lst):
                self.drivers = [_as_driver(a) for a in lst]
            else:
                self.drivers = [ScalarDriver.getConstantDriver(float(a)) for a in lst]
        else:
            raise TypeError("AxialDriver: incompatible constructor arguments")

    @staticmethod
    def getVectorAxialDriver(x: float, y: float, z: float) -> "AxialDriver":
        return AxialDriver(x, y, z)

    @staticmethod
    def getUniAxialDriver(driver: ScalarDriver, axis: Axis) -> "AxialDriver":
        n = NullDriver
        if axis == Axis.xaxis:
            return AxialDriver(driver, n(), n())
        if axis == Axis.yaxis:
            return AxialDriver(n(), driver, n())
        if axis == Axis.zaxis:
            return AxialDriver(n(), n(), driver)
        if axis == Axis.all:
            return AxialDriver(driver, driver, driver)
        return AxialDriver(n(), n(), n())

    def getCurrentAxialDrivers(self, time: float) -> CVector:
        return CVector(*(d.getCurrentScalarValue(time) for d in self.drivers))

    def getConstantValues(self) -> CVector:
        return CVector(*(d.constantValue for d in self.drivers))

    def applyMask(self, mask) -> None:
        m = [int(v) for v in (mask.tolist() if isinstance(mask, CVector) else mask)]
        assert len(m) == 3
        for i in range(3):
            if m[i] == 0:
                self.drivers[i] = ScalarDriver()
            elif m[i] != 1:
                raise RuntimeError("Invalid mask value, mask must be binary!")

    def _clone(self) -> "AxialDriver":
        a = AxialDriver.__new__(AxialDriver)
        a.drivers = [d._clone() for d in self.drivers]
        return a


def _as_axial(d) -> AxialDriver:
    if not isinstance(d, AxialDriver):
        raise TypeError(f"expected AxialDriver, got {type(d).__name__}")
    return d._clone()


# aliases (cmtj.cpp:41-98)
def constantDriver(value: float) -> ScalarDriver:
    return ScalarDriver.getConstantDriver(value)


def pulseDriver(constantValue: float, amplitude: float, period: float, cycle: float) -> ScalarDriver:
    return ScalarDriver.getPulseDriver(constantValue, amplitude, period, cycle)


def sineDriver(constantValue: float, amplitude: float, frequency: float, phase: float) -> ScalarDriver:
    return ScalarDriver.getSineDriver(constantValue, amplitude, frequency, phase)


def posSineDriver(constantValue: float, amplitude: float, frequency: float, phase: float) -> ScalarDriver:
    return ScalarDriver.getPosSineDriver(constantValue, amplitude, frequency, phase)


def stepDriver(constantValue: float, amplitude: float, timeStart: float, timeStop: float) -> ScalarDriver:
    return ScalarDriver.getStepDriver(constantValue, amplitude, timeStart, timeStop)


def trapezoidDriver(
    constantValue: float, amplitude: float, timeStart: float, edgeTime: float, steadyTime: float
) -> ScalarDriver:
    return ScalarDriver.getTrapezoidDriver(constantValue, amplitude, timeStart, edgeTime, steadyTime)


def gaussianImpulseDriver(constantValue: float, amplitude: float, t0: float, sigma: float) -> ScalarDriver:
    return ScalarDriver.getGaussianImpulseDriver(constantValue, amplitude, t0, sigma)


def gaussianStepDriver(constantValue: float, amplitude: float, t0: float, sigma: float) -> ScalarDriver:
    return ScalarDriver.getGaussianStepDriver(constantValue, amplitude, t0, sigma)


# --------------------------------------------------------------------------------------
# Layer
# --------------------------------------------------------------------------------------
class AdaptiveIntegrationParams:
    """core/junction.hpp:125-139"""

    def __init__(self):
        self.abs_tol = 1e-6
        self.rel_tol = 1e-3
        self.max_factor = 5.0
        self.min_factor = 0.1
        self.safety_factor = 0.9
        self.use_pid_control = False
        self.ki = 0.0
        self.kp = 0.2
        self.kd = 0.0
        self.prev_error_ratio = 1.0
        self.integral_error = 0.0


_AXIAL_SLOTS = {
    "Hext": _abi.SLOT_HEXT,
    "Hoe": _abi.SLOT_HOE,
    "Hdmi": _abi.SLOT_HDMI,
    "Hres": _abi.SLOT_HRES,
    "IDMItop": _abi.SLOT_IDMI_TOP,
    "IDMIbottom": _abi.SLOT_IDMI_BOTTOM,
}


class Layer:
    """One macrospin layer: the data of ``Layer<T>`` (core/junction.hpp:145-260)."""

    def __init__(
        self,
        id: str,
        mag,
        anis,
        Ms: float,
        thickness: float,
        cellSurface: float,
        demagTensor,
        damping: float = 0.011,
        *,
        _fieldLikeTorque: float = 0.0,
        _dampingLikeTorque: float = 0.0,
        _slonczewski: float = 0.0,
        _beta: float = 0.0,
        _spinPolarisation: float = 0.0,
    ):
        mag = CVector._coerce(mag).copy()
        anis = CVector._coerce(anis).copy()
        self._id = str(id)
        # the reference stores the UN-normalised input (junction.hpp:178-195, SURVEY note 10)
        self.mag = mag
        self.anis = anis
        self._Ms = float(Ms)
        self._thickness = float(thickness)
        self._cellSurface = float(cellSurface)
        self._demagTensor = _tensor(demagTensor)
        self._damping = float(damping)
        self.fieldLikeTorque = float(_fieldLikeTorque)
        self.dampingLikeTorque = float(_dampingLikeTorque)
        self.SlonczewskiSpacerLayerParameter = float(_slonczewski)
        self.beta = float(_beta)
        self.spinPolarisation = float(_spinPolarisation)
        if mag.length() == 0:
            raise RuntimeError("Initial magnetisation was set to a zero vector!")
        if anis.length() == 0:
            raise RuntimeError("Anisotropy was set to a zero vector!")
        self.kappa = 1.0
        self.includeSTT = False
        self.includeSOT = False
        self.dynamicSOT = True
        self.alternativeSTTSet = False
        self.temperatureSet = False
        self.nonStochasticTempSet = False
        self.pinkNoiseSet = False
        self.referenceType = Reference.none
        self.referenceLayer = CVector()
        self.secondaryReferenceLayer = CVector()
        self.dipoleBottom = [CVector(), CVector(), CVector()]
        self.dipoleTop = [CVector(), CVector(), CVector()]
        self.noiseAlpha = 1.0
        self.noiseScale = 0.0
        self.noiseStd = 0.0
        self.noiseAxis = Axis.all
        self.adaptiveParams = AdaptiveIntegrationParams()
        self._bfn = None
        # every driver slot starts as the default (constant 0) ScalarDriver
        self.drivers: List[ScalarDriver] = [ScalarDriver() for _ in range(_abi.NSLOTS)]

    # -- factories (cmtj.cpp:238-246) ---------------------------------------------------
    @staticmethod
    def createSOTLayer(
        id: str,
        mag,
        anis,
        Ms: float,
        thickness: float,
        cellSurface: float,
        demagTensor,
        damping: float = 0.011,
        fieldLikeTorque: float = 0.0,
        dampingLikeTorque: float = 0.0,
    ) -> "Layer":
        l = Layer(
            id,
            mag,
            anis,
            Ms,
            thickness,
            cellSurface,
            demagTensor,
            damping,
            _fieldLikeTorque=fieldLikeTorque,
            _dampingLikeTorque=dampingLikeTorque,
        )
        l.includeSTT = False
        l.includeSOT = True
        l.dynamicSOT = False
        return l

    @staticmethod
    def createSTTLayer(
        id: str,
        mag,
        anis,
        Ms: float,
        thickness: float,
        cellSurface: float,
        demagTensor,
        damping: float = 0.011,
        SlonczewskiSpacerLayerParameter: float = 1.0,
        beta: float = 0.0,
        spinPolarisation: float = 0.0,
    ) -> "Layer":
        l = Layer(
            id,
            mag,
            anis,
            Ms,
            thickness,
            cellSurface,
            demagTensor,
            damping,
            _slonczewski=SlonczewskiSpacerLayerParameter,
            _beta=beta,
            _spinPolarisation=spinPolarisation,
        )
        l.includeSTT = True
        l.includeSOT = False
        return l

    # -- read-only properties (cmtj.cpp:274-279) ------------------------------------------
    id = property(lambda self: self._id)
    Ms = property(lambda self: self._Ms)
    thickness = property(lambda self: self._thickness)
    damping = property(lambda self: self._damping)
    cellSurface = property(lambda self: self._cellSurface)
    demagTensor = property(lambda self: [r.copy() for r in self._demagTensor])

    def getId(self) -> str:
        return self._id

    # -- setters ---------------------------------------------------------------------------
    def _set_scalar(self, slot: int, driver: ScalarDriver) -> None:
        self.drivers[slot] = _as_driver(driver)

    def _set_axial(self, slot0: int, driver: AxialDriver) -> None:
        d = _as_axial(driver)
        for i in range(3):
            self.drivers[slot0 + i] = d.drivers[i]

    def setMagnetisation(self, mag) -> None:
        mag = CVector._coerce(mag).copy()
        if mag.length() == 0:
            raise RuntimeError("Initial magnetisation was set to a zero vector!")
        mag.normalize()
        self.mag = mag

    def setAnisotropyDriver(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_K1, driver)

    def setSecondOrderAnisotropyDriver(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_K2, driver)

    def setExternalFieldDriver(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_HEXT, driver)

    def setOerstedFieldDriver(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_HOE, driver)

    def setHdmiDriver(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_HDMI, driver)

    def setReservedInteractionField(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_HRES, driver)

    def setCurrentDriver(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_CURRENT, driver)

    def setIECDriverTop(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_IEC_TOP, driver)

    def setIECDriverBottom(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_IEC_BOTTOM, driver)

    def setQuadIECDriverTop(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_IECQ_TOP, driver)

    def setQuadIECDriverBottom(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_IECQ_BOTTOM, driver)

    def setIDMIDriverTop(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_IDMI_TOP, driver)

    def setIDMIDriverBottom(self, driver: AxialDriver) -> None:
        self._set_axial(_abi.SLOT_IDMI_BOTTOM, driver)

    def setReferenceLayer(self, reference) -> None:
        """Overloaded like the reference (junction.hpp:538-560): CVector or Reference enum."""
        if isinstance(reference, Reference):
            if reference == Reference.fixed and not self.referenceLayer.length():
                raise RuntimeError(
                    "Cannot set fixed polarisation layer to 0! Set reference to NONE to disable reference."
                )
            self.referenceType = reference
        else:
            self.referenceLayer = CVector._coerce(reference).copy()
            self.referenceType = Reference.fixed

    def setSecondaryReferenceLayer(self, reference) -> None:
        self.secondaryReferenceLayer = CVector._coerce(reference).copy()
        self.referenceType = Reference.fixed

    def getReferenceLayer(self) -> CVector:
        return self.referenceLayer.copy()

    def getSecondaryReferenceLayer(self) -> CVector:
        return self.secondaryReferenceLayer.copy()

    def getReferenceType(self) -> Reference:
        return self.referenceType

    def _setTorqueParameters(self) -> None:
        # junction.hpp:431-439 -- includeSOT is switched on before the checks throw
        self.includeSOT = True
        if self.includeSTT:
            raise RuntimeError("includeSTT was on and now setting SOT interaction!")
        if not self.dynamicSOT:
            raise RuntimeError("used a static SOT definition, now trying to set it dynamically!")

    def setFieldLikeTorqueDriver(self, driver: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_FL, driver)

    def setDampingLikeTorqueDriver(self, driver: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_DL, driver)

    def setSecondaryFieldLikeTorqueDriver(self, driver: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_FL2, driver)

    def setSecondaryDampingLikeTorqueDriver(self, driver: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_DL2, driver)

    def setPrimaryTorqueDrivers(self, fieldLikeTorque: ScalarDriver, dampingLikeTorque: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_FL, fieldLikeTorque)
        self._set_scalar(_abi.SLOT_DL, dampingLikeTorque)

    def setSecondaryTorqueDrivers(self, fieldLikeTorque: ScalarDriver, dampingLikeTorque: ScalarDriver) -> None:
        self._setTorqueParameters()
        self._set_scalar(_abi.SLOT_FL2, fieldLikeTorque)
        self._set_scalar(_abi.SLOT_DL2, dampingLikeTorque)

    def setTemperatureDriver(self, driver: ScalarDriver) -> None:
        self._set_scalar(_abi.SLOT_TEMPERATURE, driver)
        self.temperatureSet = True

    def setNonStochasticLangevinDriver(self, driver: ScalarDriver) -> None:
        # junction.hpp:390-395: the SDE flag is NOT set => no effect on dynamics (SURVEY note 6)
        self._set_scalar(_abi.SLOT_TEMPERATURE, driver)
        self.temperatureSet = False
        self.nonStochasticTempSet = True

    def hasTemperature(self) -> bool:
        return self.temperatureSet

    def setTopDipoleTensor(self, tensor) -> None:
        self.dipoleTop = _tensor(tensor)

    def setBottomDipoleTensor(self, tensor) -> None:
        self.dipoleBottom = _tensor(tensor)

    def setKappa(self, kappa: float) -> None:
        self.kappa = float(kappa)

    def setAlternativeSTT(self, alternativeSTT: bool) -> None:
        self.alternativeSTTSet = bool(alternativeSTT)

    def setAlphaNoise(self, alpha: float, std: float, scale: float, axis: Axis = Axis.all) -> None:
        if alpha < 0 or alpha > 2:
            raise RuntimeError("alpha must be between 0 and 2")
        self.noiseAlpha = float(alpha)
        self.noiseStd = float(std)
        self.noiseScale = float(scale)
        self.noiseAxis = axis
        self.pinkNoiseSet = True

    def setOneFNoise(self, sources: int, bias: float, scale: float) -> None:
        # constructed but never reaches the dynamics in the reference (SURVEY note 6)
        self.pinkNoiseSet = True

    def createBufferedAlphaNoise(self, bufferSize: int) -> None:
        from .noise import VectorAlphaNoise

        if self.noiseAlpha < 0:
            raise RuntimeError(
                "alpha must be set before creating the noise! Use setAlphaNoise function to set the alpha parameter."
            )
        self._bfn = VectorAlphaNoise(int(bufferSize), self.noiseAlpha, self.noiseStd, self.noiseScale, self.noiseAxis)

    def getOneFVector(self) -> CVector:
        if self.noiseScale != 0 and self._bfn is not None:
            return self._bfn.tickVector()
        return CVector()

    def setAdaptiveParams(self, params: AdaptiveIntegrationParams) -> None:
        self.adaptiveParams = copy.copy(params)

    def _clone(self) -> "Layer":
        l = copy.copy(self)
        l.mag = self.mag.copy()
        l.anis = self.anis.copy()
        l._demagTensor = [r.copy() for r in self._demagTensor]
        l.dipoleBottom = [r.copy() for r in self.dipoleBottom]
        l.dipoleTop = [r.copy() for r in self.dipoleTop]
        l.referenceLayer = self.referenceLayer.copy()
        l.secondaryReferenceLayer = self.secondaryReferenceLayer.copy()
        l.adaptiveParams = copy.copy(self.adaptiveParams)
        l.drivers = [d._clone() for d in self.drivers]
        return l

    # -- packing into the C-ABI record -------------------------------------------------------
    def _pack(self, out: _abi.Layer) -> None:
        out.mag[:] = self.mag.tolist()
        out.anis[:] = self.anis.tolist()
        out.Ms = self._Ms
        out.thickness = self._thickness
        out.cellSurface = self._cellSurface
        out.damping = self._damping
        for name, t in (
            ("demagTensor", self._demagTensor),
            ("dipoleBottom", self.dipoleBottom),
            ("dipoleTop", self.dipoleTop),
        ):
            arr = getattr(out, name)
            for r in range(3):
                for c in range(3):
                    arr[3 * r + c] = t[r][c]
        out.fieldLikeTorque = self.fieldLikeTorque
        out.dampingLikeTorque = self.dampingLikeTorque
        out.SlonczewskiSpacerLayerParameter = self.SlonczewskiSpacerLayerParameter
        out.beta = self.beta
        out.kappa = self.kappa
        out.spinPolarisation = self.spinPolarisation
        out.referenceLayer[:] = self.referenceLayer.tolist()
        out.secondaryReferenceLayer[:] = self.secondaryReferenceLayer.tolist()
        out.noiseAlpha = self.noiseAlpha
        out.noiseStd = self.noiseStd
        out.noiseScale = self.noiseScale
        p = self.adaptiveParams
        for n in (
            "abs_tol",
            "rel_tol",
            "max_factor",
            "min_factor",
            "safety_factor",
            "ki",
            "kp",
            "kd",
            "prev_error_ratio",
            "integral_error",
        ):
            setattr(out, n, float(getattr(p, n)))
        out.referenceType = int(self.referenceType)
        out.includeSTT = int(self.includeSTT)
        out.includeSOT = int(self.includeSOT)
        out.dynamicSOT = int(self.dynamicSOT)
        out.alternativeSTT = int(self.alternativeSTTSet)
        out.temperatureSet = int(self.temperatureSet)
        out.noiseAxis = int(self.noiseAxis)
        out.use_pid_control = int(bool(p.use_pid_control))
        for s in range(_abi.NSLOTS):
            self.drivers[s]._pack(out.drivers[s])


# --------------------------------------------------------------------------------------
# Junction
# --------------------------------------------------------------------------------------
MR_NONE, MR_CLASSIC, MR_STRIP = 0, 1, 2
_VEC = ("x", "y", "z")


def _is_number(v) -> bool:
    return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool)


class Junction:
    """Multilayer junction: ``Junction<T>`` of the reference (core/junction.hpp:1139-2057).

    ``runSimulation`` integrates on the GPU (a batch of one instance); see
    :class:`cmtj_b200.batch.JunctionBatch` for parameter sweeps, which is where the device
    is actually filled.
    """

    def __init__(self, layers: Sequence[Layer], *args, **kwargs):
        layers = list(layers)
        names3 = ("Rx0", "Ry0", "AMR_X", "AMR_Y", "SMR_X", "SMR_Y", "AHE")
        strip = None
        classic = None
        if args and not _is_number(args[0]):
            if len(args) + len(kwargs) != 7:
                raise TypeError("Junction(): incompatible constructor arguments")
            vals = list(args) + [kwargs[n] for n in names3[len(args) :]]
            strip = [[float(v) for v in lst] for lst in vals]
        elif any(k in kwargs for k in names3):
            strip = [[float(v) for v in kwargs[n]] for n in names3]
        elif args or kwargs:
            Rp = args[0] if len(args) > 0 else kwargs.get("Rp", 100)
            Rap = args[1] if len(args) > 1 else kwargs.get("Rap", 200)
            classic = (float(Rp), float(Rap))

        self._layers: List[Layer] = [l._clone() for l in layers]
        self.layerNo = len(self._layers)
        self.MR_mode = MR_NONE
        self.Rp = 0.0
        self.Rap = 0.0
        self.Rtag = "R"
        self._strip = None
        self._log: Dict[str, list] = {}
        self._logLength = 0
        if self.layerNo == 0:
            raise ValueError("Passed a zero length Layer vector!")
        if strip is None:
            ids = set()
            for l in self._layers:
                if l.id in ids:
                    raise ValueError("Layers must have unique ids!")
                ids.add(l.id)
        if classic is not None:
            if self.layerNo == 1 and not self._layers[0].referenceLayer.length():
                raise ValueError("MTJ with a single layer must have a pinning (referenceLayer) set!")
            if self.layerNo > 2:
                raise ValueError(
                    "This constructor supports only bilayers! Choose the other one with the strip resistance!"
                )
            self.Rp, self.Rap = classic
            self.MR_mode = MR_CLASSIC
            if self.layerNo == 2:
                self.Rtag = "R_" + self._layers[0].id + "_" + self._layers[1].id
        if strip is not None:
            if any(len(v) != self.layerNo for v in strip):
                raise ValueError("Layers and Rx0, Ry, AMR, AMR and SMR must be of the same size!")
            self._strip = dict(zip(names3, strip))
            self.MR_mode = MR_STRIP

    def _clone(self) -> "Junction":
        """Value semantics of ``std::vector<Junction>`` arguments (Stack / GroupInteraction copy
        their junctions, log included)."""
        j = copy.copy(self)
        j._layers = [l._clone() for l in self._layers]
        j._log = {k: list(v) for k, v in self._log.items()}
        j._strip = None if self._strip is None else {k: list(v) for k, v in self._strip.items()}
        return j

    def _append_samples(self, channels: np.ndarray, times: np.ndarray) -> None:
        """logLayerParams for a block of samples: channels [n_channels, n_samples]."""
        for name, row in zip(self._channel_names(), channels):
            self._log.setdefault(name, []).extend(row.tolist())
        self._log.setdefault("time", []).extend(np.asarray(times).tolist())
        self._logLength += len(times)

    # -- introspection -----------------------------------------------------------------------
    @property
    def layers(self) -> List[Layer]:
        return [l._clone() for l in self._layers]

    def getLayerIds(self) -> List[str]:
        return [l.id for l in self._layers]

    def getLayer(self, layerId: str) -> Layer:
        for l in self._layers:
            if l.id == layerId:
                return l
        raise RuntimeError("Failed to find a layer with a given id " + layerId + "!")

    # -- log -----------------------------------------------------------------------------------
    def clearLog(self) -> None:
        self._log = {}
        self._logLength = 0

    def getLog(self) -> Dict[str, List[float]]:
        return {k: list(v) for k, v in self._log.items()}

    def saveLog(self, filename: str) -> None:
        """core/junction.hpp:1638-1656: ';'-separated, one column per key."""
        if filename == "":
            raise RuntimeError("The filename may not be empty!")
        keys = list(self._log.keys())
        with open(filename, "w") as f:
            f.write("".join(k + ";" for k in keys) + "\n")
            for i in range(self._logLength):
                f.write("".join(f"{self._log[k][i]:g};" for k in keys) + "\n")

    # -- setter fan-out (junction.hpp:1269-1559) ---------------------------------------------------
    def _fan(self, layerID: str, fn: Callable[[Layer], None]) -> None:
        found = False
        for l in self._layers:
            if l.id == layerID or layerID == "all":
                fn(l)
                found = True
        if not found:
            raise RuntimeError("Failed to find a layer with a given id: " + layerID + "!")

    def _fan_or_get(self, layerID: str, fn: Callable[[Layer], None]) -> None:
        if layerID == "all":
            for l in self._layers:
                fn(l)
        else:
            fn(self.getLayer(layerID))

    def setLayerTemperatureDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setTemperatureDriver(driver))

    def setLayerNonStochasticLangevinDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setNonStochasticLangevinDriver(driver))

    def setLayerAnisotropyDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setAnisotropyDriver(driver))

    def setLayerSecondOrderAnisotropyDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setSecondOrderAnisotropyDriver(driver))

    def setLayerExternalFieldDriver(self, layerID: str, driver: AxialDriver) -> None:
        self._fan(layerID, lambda l: l.setExternalFieldDriver(driver))

    def setLayerOerstedFieldDriver(self, layerID: str, driver: AxialDriver) -> None:
        self._fan(layerID, lambda l: l.setOerstedFieldDriver(driver))

    def setLayerCurrentDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setCurrentDriver(driver))

    def setLayerDampingLikeTorqueDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setDampingLikeTorqueDriver(driver))

    def setLayerFieldLikeTorqueDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setFieldLikeTorqueDriver(driver))

    def setLayerSecondaryFieldLikeTorqueDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setSecondaryFieldLikeTorqueDriver(driver))

    def setLayerSecondaryDampingLikeTorqueDriver(self, layerID: str, driver: ScalarDriver) -> None:
        self._fan(layerID, lambda l: l.setSecondaryDampingLikeTorqueDriver(driver))

    def setLayerPrimaryTorqueDrivers(
        self, layerId: str, fieldLikeTorque: ScalarDriver, dampingLikeTorque: ScalarDriver
    ) -> None:
        self._fan(layerId, lambda l: l.setFieldLikeTorqueDriver(fieldLikeTorque))
        self._fan(layerId, lambda l: l.setDampingLikeTorqueDriver(dampingLikeTorque))

    def setLayerSecondaryTorqueDrivers(
        self, layerId: str, fieldLikeTorque: ScalarDriver, dampingLikeTorque: ScalarDriver
    ) -> None:
        self._fan(layerId, lambda l: l.setSecondaryFieldLikeTorqueDriver(fieldLikeTorque))
        self._fan(layerId, lambda l: l.setSecondaryDampingLikeTorqueDriver(dampingLikeTorque))

    def setLayerReservedInteractionField(self, layerID: str, driver: AxialDriver) -> None:
        self._fan(layerID, lambda l: l.setReservedInteractionField(driver))

    def setLayerHdmiDriver(self, layerID: str, driver: AxialDriver) -> None:
        self._fan(layerID, lambda l: l.setHdmiDriver(driver))

    def setLayerAlternativeSTT(self, layerID: str, alternative: bool) -> None:
        self._fan_or_get(layerID, lambda l: l.setAlternativeSTT(alternative))

    def setLayerOneFNoise(self, layerID: str, sources: int, bias: float, scale: float) -> None:
        self._fan_or_get(layerID, lambda l: l.setOneFNoise(sources, bias, scale))

    def setLayerMagnetisation(self, layerID: str, mag) -> None:
        self._fan(layerID, lambda l: l.setMagnetisation(mag))

    def getLayerMagnetisation(self, layerID: str) -> CVector:
        return self.getLayer(layerID).mag.copy()

    def getLayerReferenceType(self, layerID: str) -> Reference:
        return self.getLayer(layerID).referenceType

    def setLayerReferenceLayer(self, layerID: str, referenceLayer) -> None:
        v = CVector._coerce(referenceLayer)
        self._fan_or_get(layerID, lambda l: l.setReferenceLayer(v))

    def setLayerReferenceType(self, layerID: str, referenceType: Reference) -> None:
        if not isinstance(referenceType, Reference):
            raise TypeError("setLayerReferenceType expects a Reference enum")
        self._fan_or_get(layerID, lambda l: l.setReferenceLayer(referenceType))

    def _couple(self, bottomLayer: str, topLayer: str, set_top: Callable, set_bottom: Callable) -> None:
        """junction.hpp:1308-1371: coupling only between ADJACENT layers, either order."""
        for i in range(self.layerNo - 1):
            a, b = self._layers[i], self._layers[i + 1]
            if (a.id == bottomLayer and b.id == topLayer) or (a.id == topLayer and b.id == bottomLayer):
                set_top(a)
                set_bottom(b)
                return
        raise RuntimeError(
            "Failed to match the layer order or find layer ids: " + bottomLayer + " and " + topLayer + "!"
        )

    def setIECDriver(self, bottomLayer: str, topLayer: str, driver: ScalarDriver) -> None:
        self._couple(bottomLayer, topLayer, lambda l: l.setIECDriverTop(driver), lambda l: l.setIECDriverBottom(driver))

    def setQuadIECDriver(self, bottomLayer: str, topLayer: str, driver: ScalarDriver) -> None:
        self._couple(
            bottomLayer, topLayer, lambda l: l.setQuadIECDriverTop(driver), lambda l: l.setQuadIECDriverBottom(driver)
        )

    def setIDMIDriver(self, bottomLayer: str, topLayer: str, driver: AxialDriver) -> None:
        self._couple(
            bottomLayer, topLayer, lambda l: l.setIDMIDriverTop(driver), lambda l: l.setIDMIDriverBottom(driver)
        )

    # -- magnetoresistance (junction.hpp:1835-1889); host formula for the API getter only ----------------
    def getMagnetoresistance(self) -> List[float]:
        L = self._layers
        if self.MR_mode == MR_CLASSIC and self.layerNo == 2:
            return [self.Rp + (((self.Rap - self.Rp) / 2.0) * (1.0 - c_dot(L[0].mag, L[1].mag)))]
        if self.MR_mode == MR_CLASSIC and self.layerNo == 1:
            return [self.Rp + (((self.Rap - self.Rp) / 2.0) * (1.0 - c_dot(L[0].mag, L[0].referenceLayer)))]
        if self.MR_mode == MR_STRIP:
            s = self._strip
            rx = ry = 0.0
            for i, l in enumerate(L):
                m = l.mag
                Rx = s["Rx0"][i] + s["AMR_X"][i] * (m.x * m.x) + s["SMR_X"][i] * (m.y * m.y)
                Ry = s["Ry0"][i] + 0.5 * s["AHE"][i] * m.z + (s["AMR_Y"][i] + s["SMR_Y"][i]) * m.x * m.y
                rx += 1.0 / Rx
                ry += 1.0 / Ry
            return [1 / rx, 1 / ry, 0.0]
        raise RuntimeError("Magnetisation calculation is not supported for this structure!")

    # -- packing ---------------------------------------------------------------------------------------
    def _channel_names(self) -> List[str]:
        names = []
        for l in self._layers:
            names += [f"{l.id}_m{v}" for v in _VEC]
        if self.MR_mode == MR_CLASSIC:
            names.append("R" if self.layerNo == 1 else self.Rtag)
        elif self.MR_mode == MR_STRIP:
            names += ["Rx", "Ry", "Rz"]
        return names

    def _energy_channel_names(self) -> List[str]:
        """Keys added by calculateEnergies=True, in the order of the CMTJB_LOG_ENERGY channels
        (logLayerParams, junction.hpp:1584-1613)."""
        names = []
        for l in self._layers:
            names += [f"{l.id}_{k}" for k in ("K", "K2", "Jbottom", "Jtop", "I")]
            groups = ["Hext", "Hiec", "Hanis", "Hdemag", "Hth"] + (["Hfl", "Hdl"] if l.includeSOT else [])
            for g in groups:
                names += [f"{l.id}_{g}{v}" for v in _VEC]
        return names

    def _pack(self):
        if self.layerNo > _abi.MAX_LAYERS:
            raise NotImplementedError(f"junctions with more than {_abi.MAX_LAYERS} layers are not supported")
        jn = _abi.Junction()
        jn.n_layers = self.layerNo
        jn.mr_mode = self.MR_mode
        jn.Rp = self.Rp
        jn.Rap = self.Rap
        jn.MAGNETIC_PERMEABILITY = PhysicalConstants.magnetic_permeability()
        jn.GYRO = PhysicalConstants.gyromagnetic_ratio()
        jn.TtoAm = PhysicalConstants.TtoAm()
        jn.HBAR = PhysicalConstants.hbar()
        jn.ELECTRON_CHARGE = PhysicalConstants.elementary_charge()
        jn.BOLTZMANN_CONST = PhysicalConstants.boltzmann_constant()
        layers = (_abi.Layer * self.layerNo)()
        for i, l in enumerate(self._layers):
            l._pack(layers[i])
            if self._strip is not None:
                for n, vals in self._strip.items():
                    setattr(layers[i], n, vals[i])
        return jn, layers

    def _custom_table(self, layers, plan, timeStep: float) -> Optional[np.ndarray]:
        """``ScalarDriver.getCustomDriver`` callbacks (drivers.hpp:247-260, :294-303) evaluated where a
        fixed-step solver will ask for them: the three stage times ``i*dt, i*dt + 0.5*dt, i*dt + dt`` of
        every step (junction.hpp:1974, :1020-1033; Heun reads stages 0 and 2, Euler-Heun stage 0), in the
        reference's own floating-point expressions.  Assigns each custom driver its column in the packed
        ``layers`` and returns the ``[3 * total_iterations, n_custom]`` table (``cmtjb_run_args_t::
        custom_table``), or None when the model has no custom driver.  A raising callback surfaces as the
        reference's ``RuntimeError("Error in Python callback")``."""
        cols: List[ScalarDriver] = []
        for i, l in enumerate(self._layers):
            for sidx, d in enumerate(l.drivers):
                if d.update == UpdateType.custom:
                    layers[i].drivers[sidx].custom_column = len(cols)
                    cols.append(d)
        if not cols:
            return None
        n, dt = int(plan.total_iterations), float(timeStep)
        solver = int(plan.effective_solver)
        stages = (0, 1, 2) if solver == int(SolverMode.RK4) else ((0, 2) if solver == int(SolverMode.Heun) else (0,))
        tab = np.zeros((3 * max(n, 1), len(cols)), dtype=np.float64)
        for i in range(n):
            t0 = i * dt
            ts = (t0, t0 + 0.5 * dt, t0 + dt)
            for st in stages:
                for c, d in enumerate(cols):
                    tab[3 * i + st, c] = d.getCurrentScalarValue(ts[st])
        return tab

    # -- the hot path entry ---------------------------------------------------------------------------
    def runSimulation(
        self,
        totalTime: float,
        timeStep: float = 1e-13,
        writeFrequency: float = 1e-11,
        verbose: bool = False,
        calculateEnergies: bool = False,
        solverMode: SolverMode = SolverMode.RK4,
    ) -> None:
        """``Junction::runSimulation`` (core/junction.hpp:1948-2056) on the GPU."""
        import time as _time

        be = _abi.backend()
        t0 = _time.perf_counter()
        jn, layers = self._pack()
        args = _abi.RunArgs()
        args.thermal_normals = _thermal_normals
        args.totalTime = float(totalTime)
        args.timeStep = float(timeStep)
        args.writeFrequency = float(writeFrequency)
        args.solverMode = int(solverMode)
        args.log_mode = _abi.LOG_ENERGY if calculateEnergies else _abi.LOG_FULL
        args.seed = _next_seed()
        args.n_instances = 1
        args.instance_offset = 0
        args.start_iteration = 0
        args.onef_buffer = None  # 1/f^alpha noise (setAlphaNoise) is generated by the library on the device
        plan = be.plan(jn, layers, args)
        if plan.solver_overridden:
            print("[WARNING] Solver automatically changed to Euler Heun for stochastic calculation.")
        custom = self._custom_table(layers, plan, timeStep)
        if custom is not None:
            args.custom_table = custom.ctypes.data
            args.n_custom = custom.shape[1]
        n_ch, n_s = int(plan.n_channels), int(plan.n_samples)
        m_state = np.array([c for l in self._layers for c in l.mag.tolist()], dtype=np.float64).reshape(-1, 1)
        log = np.zeros((n_ch, max(n_s, 1), 1), dtype=np.float64)
        status = np.zeros(1, dtype=np.int32)
        nsamp = np.zeros(1, dtype=np.int32) if plan.variable_samples else None
        tlog = np.zeros((max(n_s, 1), 1), dtype=np.float64) if plan.variable_samples else None
        rc = be.run_host(jn, layers, None, 0, args, m_state, log, status, nsamp, tlog)
        # state persists between runs: write back even when the run failed
        for i, l in enumerate(self._layers):
            l.mag = CVector(m_state[3 * i, 0], m_state[3 * i + 1, 0], m_state[3 * i + 2, 0])
        if plan.variable_samples:
            n_s = int(nsamp[0])
            times = tlog[:n_s, 0]
        else:
            we = int(plan.write_every)
            times = np.arange(0, int(plan.total_iterations), we, dtype=np.float64) * float(timeStep)
        if rc == _abi.OK or rc == _abi.ENAN:
            names = self._channel_names() + (self._energy_channel_names() if calculateEnergies else [])
            # the reference pushes "<id>_I" twice per sample for STT / SOT layers (junction.hpp:1593, :1611-1612)
            twice = {f"{l.id}_I" for l in self._layers if l.includeSTT or l.includeSOT} if calculateEnergies else set()
            for name, ch in zip(names, range(n_ch)):
                vals = log[ch, :n_s, 0]
                self._log.setdefault(name, []).extend((np.repeat(vals, 2) if name in twice else vals).tolist())
            self._log.setdefault("time", []).extend(times[:n_s].tolist())
            self._logLength += n_s
        be.check(rc)
        if verbose:
            print(f"Steps in simulation: {int(plan.total_iterations)}")
            print(f"Write every: {int(plan.write_every)}")
            print(f"Simulation time = {int(_time.perf_counter() - t0)}[s]")


_seed_rng = np.random.default_rng()  # OS entropy, like std::random_device


def seed(value: Optional[int] = None) -> None:
    """Re-seed the private generator the Philox keys of stochastic runs are drawn from
    (``None`` = OS entropy).  An addition: the reference cannot repeat a stochastic run."""
    global _seed_rng
    _seed_rng = np.random.default_rng(value)


_thermal_normals = 0


def thermal_normals(precision: Optional[str] = None) -> str:
    """Precision of the Langevin normals of all subsequent stochastic runs of this process: ``"float32"`` (default: Philox +
    Box-Muller on the SFU, absolute error 2^-21, tail to 6.7 sigma) or ``"float64"`` (53-bit uniforms, double-precision
    Box-Muller, tail to 8.6 sigma; the reference draws ``std::normal_distribution<double>``).  Returns the current setting.
    An addition (``cmtjb_run_args_t::thermal_normals``)."""
    global _thermal_normals
    if precision is not None:
        if precision not in ("float32", "float64"):
            raise ValueError("precision must be 'float32' or 'float64'")
        _thermal_normals = 1 if precision == "float64" else 0
    return "float64" if _thermal_normals else "float32"


def _next_seed() -> int:
    """Philox key for the next stochastic run.  The reference seeds every Layer from
    ``std::random_device`` (junction.hpp:174-175), i.e. runs are unrepeatable and never touch
    numpy's global generator: scripts that ``np.random.seed`` and then draw their own disturbances
    between runs (cmtj/utils/procedures.py:131-140) must see the same stream with us, so the keys
    come from a private generator (``cmtj_b200.seed`` makes runs reproducible)."""
    return int(_seed_rng.integers(0, 2**63 - 1, dtype=np.int64))
