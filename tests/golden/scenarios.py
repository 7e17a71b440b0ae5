"""Scenario functions shared by make_golden.py (run against the UNMODIFIED reference module,
oracle/_ref) and tests/test_golden.py (run against cmtj_b200 on the oracle or on the GPU).
Each takes a module exposing the reference API and returns {name: float64 array}.  They restate
the numerical cores of the reference's curated examples (tests/test_curated_examples.py) at
reduced size, using only the core API so that nothing from /root/reference is needed at run time."""
import math

import numpy as np

import helpers


def _arr(log, keys, tail=None):
    out = {}
    for k in keys:
        v = np.asarray(log[k], dtype=np.float64)
        out[k] = v if tail is None else v[-tail:]
    return out


def sto(m):
    """tests/test_curated_examples.py:885-948 (test_sto) at 5 ns: 1-layer STT oscillator, RK4."""
    demag = [m.CVector(0, 0, 0), m.CVector(0, 0, 0), m.CVector(0, 0, 1)]
    l1 = m.Layer.createSTTLayer("free", m.CVector(0, 0, 1), m.CVector(0, 0, 1), 1.0, 1.4e-9, 7e-10 * 7e-10, demag, 0.3, 1.0, 1, 1.0)
    l1.setReferenceLayer(m.CVector(0, 1.0, 1.0))
    j = m.Junction([l1], 100, 200)
    j.setLayerAnisotropyDriver("free", m.constantDriver(350e3))
    j.setLayerCurrentDriver("free", m.constantDriver(-6e9))
    j.runSimulation(5e-9, 1e-13, 1e-13)
    return _arr(j.getLog(), ["free_mx", "free_my", "free_mz", "R"], tail=100)


def field_sweep_carried(m):
    """PIMM/VSD-style sweep (cmtj/utils/procedures.py:120-141): 2-layer MTJ, the relaxed state of
    field point k seeds point k+1, step Oersted pulse, every step logged; last 20 samples kept."""
    j = helpers.two_layer_mtj(m)
    oe = m.ScalarDriver.getStepDriver(0, 50, 0, 3e-13)
    j.setLayerOerstedFieldDriver("all", m.AxialDriver(m.NullDriver(), m.NullDriver(), oe))
    out = {}
    for k, H in enumerate(np.linspace(-200e3, 200e3, 5)):
        j.clearLog()
        j.setLayerExternalFieldDriver("all", m.AxialDriver(float(H), 10e3, 1e3))
        j.runSimulation(0.2e-9, 1e-13, 1e-13)
        for name, v in _arr(j.getLog(), ["free_mx", "free_mz", "bottom_my", "R_free_bottom"], tail=20).items():
            out[f"{k}/{name}"] = v
    return out


def sot_hysteresis(m):
    """CIMS-hysteresis-like (tests/test_curated_examples.py:337-439): 1-layer SOT, step torque
    drivers whose amplitude is swept with carried state, STRIP resistance, Heun."""
    j = helpers.sot_layer(m, dynamic=True)
    out = {}
    for k, a in enumerate([-9e3, -3e3, 3e3, 9e3]):
        j.clearLog()
        j.setLayerFieldLikeTorqueDriver("free", m.ScalarDriver.getStepDriver(0, a / 3, 0.05e-9, 0.25e-9))
        j.setLayerDampingLikeTorqueDriver("free", m.ScalarDriver.getStepDriver(0, a, 0.05e-9, 0.25e-9))
        j.runSimulation(0.3e-9, 1e-13, 1e-12, solverMode=m.SolverMode.Heun)
        for name, v in _arr(j.getLog(), ["free_mz", "Rx", "Ry"], tail=30).items():
            out[f"{k}/{name}"] = v
    return out


def dormand_prince(m):
    """cims-stability style (tests/test_curated_examples.py:951-1083): 1 layer, K2, Dormand-Prince."""
    j = helpers.single_layer_mtj(m, mag=(0.2, 0.1, 0.9))
    j.setLayerSecondOrderAnisotropyDriver("free", m.constantDriver(3e3))
    j.setLayerExternalFieldDriver("free", m.AxialDriver(30e3, 0.0, 5e3))
    j.runSimulation(2e-9, 1e-12, 1e-11, solverMode=m.SolverMode.DormandPrince)
    return _arr(j.getLog(), ["time", "free_mx", "free_my", "free_mz", "R"])


def tri_layer_heun(m):
    j = helpers.tri_layer_mtj(m)
    j.setLayerExternalFieldDriver("all", m.AxialDriver(m.sineDriver(5e3, 2e3, 8e9, 0.0), m.NullDriver(), m.constantDriver(-2e4)))
    j.runSimulation(0.3e-9, 1e-13, 2e-12, solverMode=m.SolverMode.Heun)
    return _arr(j.getLog(), ["free_mx", "middle_my", "bottom_mz"])


def series_stack(m):
    import test_coupled_oracle_vs_ref as T

    st = m.stack if hasattr(m, "stack") else m.stack
    out = {}
    for kcl in (True, False):
        js = [T._stt_two_layer(m), T._stt_two_layer(m, (0.5, 0.1, 0.8)), T._stt_two_layer(m, (-0.3, 0.3, 0.9))]
        s = st.SeriesStack(js, "free", "bottom", 0.1, kcl)
        s.setCouplingStrength(0.25)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, 2e10, 5e9, 0.0))
        s.runSimulation(0.3e-9, 1e-13, 3e-12)
        tag = "kcl" if kcl else "chain"
        for name, v in _arr(s.getLog(), ["Resistance", "I_0", "I_2", "C_1_2"]).items():
            out[f"{tag}/{name}"] = v
        out[f"{tag}/2/free_mx"] = np.asarray(s.getLog(2)["free_mx"], dtype=np.float64)
    return out


def dipole_group(m):
    import test_coupled_oracle_vs_ref as T

    # Ms = 1.2 T: a regular (non-chaotic) array, so a 1e-9 comparison of trajectories is meaningful
    g = T._group(m, m.reservoir, n=6, Ms=1.2)
    g.runSimulation(0.2e-9, 1e-13, 2e-12)
    out = {}
    for k in (0, 3, 5):
        for name, v in _arr(g.getLog(k), ["free_mx", "free_mz"]).items():
            out[f"{k}/{name}"] = v
    return out


SCENARIOS = {
    "sto": sto,
    "field_sweep_carried": field_sweep_carried,
    "sot_hysteresis": sot_hysteresis,
    "dormand_prince": dormand_prince,
    "tri_layer_heun": tri_layer_heun,
    "series_stack": series_stack,
    "dipole_group": dipole_group,
}
