"""The checker for the sweep procedures: runs ONE sweep point the way the reference's
PIMM_procedure / VSD_procedure do (cmtj/utils/procedures.py:120-141, :253-282) through whatever
backend the facade is on (tests put it on the CPU oracle), then reduces the full-rate log with
oracle/post_oracle.py."""
import os
import sys

import numpy as np

import cmtj_b200 as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
import post_oracle  # noqa: E402


def _axial(axis, driver):
    n = cm.NullDriver
    if axis == cm.Axis.zaxis:
        return "z", cm.AxialDriver(n(), n(), driver)
    if axis == cm.Axis.yaxis:
        return "y", cm.AxialDriver(n(), driver, n())
    return "x", cm.AxialDriver(driver, n(), n())


def _tuples(params):
    return [(p.Rxx0, p.Rxy0, p.Ramr, p.Rahe, p.Rsmr, p.w, p.l) for p in params]


def pimm_point(junction, H, int_step, resistance_params, Hoe_direction=cm.Axis.zaxis, Hoe_excitation=50, Hoe_duration=3,
               simulation_duration=5e-9, wait_time=0.0, max_frequency=80e9, parallel=False, take_last_n=100):
    comp, oe = _axial(Hoe_direction, cm.ScalarDriver.getStepDriver(0, Hoe_excitation, 0, int_step * Hoe_duration))
    junction.clearLog()
    junction.setLayerExternalFieldDriver("all", cm.AxialDriver(*[cm.ScalarDriver.getConstantDriver(float(h)) for h in H]))
    junction.setLayerOerstedFieldDriver("all", oe)
    junction.runSimulation(simulation_duration, int_step, int_step)
    layers = junction.layers
    return post_oracle.pimm_point(junction.getLog(), junction.getLayerIds(), [l.thickness for l in layers], [l.Ms for l in layers],
                                  comp, int_step, _tuples(resistance_params), parallel, wait_time, max_frequency, take_last_n)


def vsd_point(junction, H, frequency, int_step, resistance_params, Hoe_direction=cm.Axis.yaxis, Hoe_excitation=50,
              simulation_duration=30e-9, Rtype="Rz", parallel=False):
    _, oe = _axial(Hoe_direction, cm.ScalarDriver.getSineDriver(0, Hoe_excitation, float(frequency), 0))
    junction.clearLog()
    junction.setLayerExternalFieldDriver("all", cm.AxialDriver(*[cm.ScalarDriver.getConstantDriver(float(h)) for h in H]))
    junction.setLayerOerstedFieldDriver("all", oe)
    junction.runSimulation(simulation_duration, int_step, int_step)
    return post_oracle.vsd_point(junction.getLog(), junction.getLayerIds(), float(frequency), int_step, Rtype,
                                 _tuples(resistance_params), parallel)


def golden(name):
    path = os.path.join(ROOT, "tests", "golden", "post_" + name + ".npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
