"""Pins the oracle (oracle/llgs_oracle.cpp, driven through the cmtj_b200 facade and its
packing code) BIT-FOR-BIT against the unmodified reference extension (oracle/_ref), on
deterministic runs.  CPU only.  If these pass, facade + packing + restatement reproduce
core/junction.hpp exactly, and the restatement can stand in for the reference on the GPU box.
"""
import math

import numpy as np
import pytest

import cmtj_b200
import helpers

MODES = ["RK4", "Heun", "EulerHeun"]


def _run_both(ref, build, T=1e-9, dt=1e-13, wf=1e-12, mode="RK4", tweak=None, **kw):
    jr = build(ref, **kw)
    jo = build(cmtj_b200, **kw)
    if tweak:
        tweak(ref, jr)
        tweak(cmtj_b200, jo)
    jr.runSimulation(T, dt, wf, solverMode=getattr(ref.SolverMode, mode))
    jo.runSimulation(T, dt, wf, solverMode=getattr(cmtj_b200.SolverMode, mode))
    return jr, jo


def _assert_same_logs(jr, jo):
    lr, lo = helpers.log_arrays(jr.getLog()), helpers.log_arrays(jo.getLog())
    assert set(lr) == set(lo)
    for k in lr:
        assert lr[k].shape == lo[k].shape, k
        assert np.array_equal(lr[k], lo[k]), (k, helpers.max_rel_err(lo[k], lr[k]))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("build", ["single_layer_mtj", "two_layer_mtj", "tri_layer_mtj", "stt_layer_mtj"])
def test_fixtures_bit_exact(ref, on_oracle, build, mode):
    jr, jo = _run_both(ref, getattr(helpers, build), mode=mode)
    _assert_same_logs(jr, jo)
    for lid in jr.getLayerIds():
        assert jr.getLayerMagnetisation(lid).tolist() == jo.getLayerMagnetisation(lid).tolist()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dynamic", [True, False])
def test_sot_strip_bit_exact(ref, on_oracle, dynamic, mode):
    jr, jo = _run_both(ref, helpers.sot_layer, T=2e-9, mode=mode, dynamic=dynamic)
    _assert_same_logs(jr, jo)
    assert jr.getMagnetoresistance() == jo.getMagnetoresistance()


def test_default_write_frequency_truncates_to_99(ref, on_oracle):
    # SURVEY 8a note 2: 1e-11/1e-13 -> 99
    jr = helpers.two_layer_mtj(ref)
    jo = helpers.two_layer_mtj(cmtj_b200)
    jr.runSimulation(2e-9)
    jo.runSimulation(2e-9)
    _assert_same_logs(jr, jo)
    t = np.asarray(jo.getLog()["time"])
    assert len(t) == 203 and t[1] == 99 * 1e-13


def test_unnormalised_initial_mag_and_carried_state(ref, on_oracle):
    # SURVEY 8a note 10 + section 5: state persists and the log appends until clearLog
    jr = helpers.two_layer_mtj(ref, mag=(0.3, -0.2, 1.0))
    jo = helpers.two_layer_mtj(cmtj_b200, mag=(0.3, -0.2, 1.0))
    assert jo.getLayerMagnetisation("free").tolist() == [0.3, -0.2, 1.0]
    for H in (0.0, 50e3, 100e3):
        for j, m in ((jr, ref), (jo, cmtj_b200)):
            j.setLayerExternalFieldDriver("all", m.AxialDriver(H, 10e3, 0.0))
            j.runSimulation(0.5e-9, 1e-13, 1e-13)
    _assert_same_logs(jr, jo)
    assert len(jo.getLog()["time"]) == 3 * 5000
    jo.clearLog()
    assert jo.getLog() == {}


def _all_driver_kinds(m):
    S = m.ScalarDriver
    return [
        S.getConstantDriver(3.5),
        S.getPulseDriver(1.0, 2.0, 1e-10, 0.3),
        S.getSineDriver(10, 20, 7e9, 0.4),
        S.getPosSineDriver(-3, 20, 7e9, 0.4),
        S.getHalfSineDriver(-3, 20, 7e9, 0.4) if hasattr(S, "getHalfSineDriver") else None,
        S.getStepDriver(0.5, 4.0, 1e-10, 3e-10),
        S.getTrapezoidDriver(0.5, 4.0, 1e-10, 0.5e-10, 2e-10),
        S.getGaussianImpulseDriver(0.1, 5.0, 3e-10, 0.4e-10),
        S.getGaussianStepDriver(0.1, 5.0, 3e-10, 0.4e-10),
    ]


def test_driver_values_match_reference(ref, oracle):
    """drivers.hpp:262-305 -- facade host evaluation, oracle evaluation and the reference agree."""
    import ctypes as C

    from cmtj_b200 import _abi

    times = np.linspace(0, 6e-10, 977)
    f = oracle.lib.oracle_eval_driver
    f.restype = C.c_int
    for dr, do in zip(_all_driver_kinds(ref), _all_driver_kinds(cmtj_b200)):
        if do is None:
            continue
        want = np.array([dr.getCurrentScalarValue(float(t)) for t in times]) if dr is not None else None
        got_py = np.array([do.getCurrentScalarValue(float(t)) for t in times])
        rec = _abi.Driver()
        do._pack(rec)
        out = np.zeros_like(times)
        f(C.byref(rec), times.ctypes.data_as(C.c_void_p), C.c_int64(len(times)), out.ctypes.data_as(C.c_void_p))
        if want is not None:
            assert np.array_equal(out, want), do
            # the pure-Python API evaluation may differ from libm-in-C++ by an ulp
            assert np.allclose(got_py, want, rtol=1e-14, atol=0), do
        else:
            assert np.array_equal(out, got_py)


@pytest.mark.parametrize("kind", range(9))
def test_time_dependent_drivers_in_dynamics(ref, on_oracle, kind):
    """Every driver type on Hoe_z, K1 and IEC of a 2-layer junction."""

    def tweak(m, j):
        d = _all_driver_kinds(m)[kind]
        if d is None:
            pytest.skip("half-sine factory not bound in the reference")
        j.setLayerOerstedFieldDriver("all", m.AxialDriver(m.NullDriver(), m.NullDriver(), d * 1e3))
        j.setLayerAnisotropyDriver("free", d * 300.0)
        j.setIECDriver("free", "bottom", d * -1e-4)

    jr, jo = _run_both(ref, helpers.two_layer_mtj, T=0.7e-9, wf=1e-13, tweak=tweak)
    _assert_same_logs(jr, jo)


def test_idmi_dipole_hdmi_k2_terms(ref, on_oracle):
    def tweak(m, j):
        j.setIDMIDriver("free", "bottom", m.AxialDriver(1e-4, -2e-4, 3e-4))
        j.setLayerHdmiDriver("free", m.AxialDriver(1e3, 2e3, -3e3))
        j.setLayerSecondOrderAnisotropyDriver("bottom", m.ScalarDriver.getConstantDriver(7e3))
        t = [m.CVector(1e-3, 2e-4, 0), m.CVector(2e-4, -3e-3, 1e-4), m.CVector(0, 1e-4, 2e-3)]
        j.getLayer("free").setBottomDipoleTensor(t)
        j.getLayer("bottom").setTopDipoleTensor(t)

    for mode in MODES:
        jr, jo = _run_both(ref, helpers.two_layer_mtj, T=0.5e-9, wf=1e-13, mode=mode, tweak=tweak)
        _assert_same_logs(jr, jo)


def test_reference_types_top_bottom(ref, on_oracle):
    def build(m):
        d = helpers._demag(m)
        a = m.Layer.createSTTLayer("free", m.CVector(0.1, 0.2, 1), m.CVector(0, 0, 1), 1.2, 1.5e-9, 1e-14, d, 0.02, 1.3, 0.1, 0.7)
        b = m.Layer.createSTTLayer("bottom", m.CVector(1, 0.1, 0.1), m.CVector(1, 0, 0), 1.5, 3e-9, 1e-14, d, 0.02, 1.3, 0.0, 0.7)
        a.setReferenceLayer(m.Reference.bottom)
        b.setReferenceLayer(m.Reference.top)
        j = m.Junction([a, b], 100, 200)
        j.setLayerAnisotropyDriver("free", m.ScalarDriver.getConstantDriver(4e5))
        j.setLayerAnisotropyDriver("bottom", m.ScalarDriver.getConstantDriver(1e5))
        j.setLayerCurrentDriver("all", m.ScalarDriver.getSineDriver(2e10, 1e10, 4e9, 0))
        j.setLayerAlternativeSTT("bottom", True)
        j.setIECDriver("free", "bottom", m.ScalarDriver.getConstantDriver(2e-5))
        return j

    for mode in MODES:
        jr, jo = _run_both(ref, build, T=0.5e-9, wf=1e-13, mode=mode)
        _assert_same_logs(jr, jo)


def test_dormand_prince_single_layer(ref, on_oracle):
    jr, jo = _run_both(ref, helpers.sot_layer, T=3e-9, dt=1e-12, wf=1e-11, mode="DormandPrince", dynamic=True)
    _assert_same_logs(jr, jo)
    assert len(jo.getLog()["time"]) > 10


def test_dormand_prince_custom_tolerances(ref, on_oracle):
    # use_pid_control=True is not pinned: the unmodified reference segfaults on it in this
    # toolchain (NaN step -> "Invalid timestep detected" -> crash), so there is nothing to compare.
    def tweak(m, j):
        p = m.AdaptiveIntegrationParams()
        p.abs_tol = 1e-7
        p.rel_tol = 1e-4
        p.max_factor = 3.0
        j.getLayer("free").setAdaptiveParams(p)

    jr, jo = _run_both(ref, helpers.single_layer_mtj, T=2e-9, dt=1e-12, wf=2e-11, mode="DormandPrince", tweak=tweak)
    _assert_same_logs(jr, jo)


def test_constants_snapshot(ref, on_oracle):
    # constants.hpp:39-44 are mutable globals read at run time
    try:
        ref.constants.PhysicalConstants.set_gyromagnetic_ratio(2.0e5)
        cmtj_b200.PhysicalConstants.set_gyromagnetic_ratio(2.0e5)
        ref.constants.PhysicalConstants.set_magnetic_permeability(1.2e-6)
        cmtj_b200.PhysicalConstants.set_magnetic_permeability(1.2e-6)
        jr, jo = _run_both(ref, helpers.two_layer_mtj, T=0.3e-9, wf=1e-13)
        _assert_same_logs(jr, jo)
    finally:
        ref.constants.PhysicalConstants.resetToDefaults()
        cmtj_b200.PhysicalConstants.resetToDefaults()
    assert cmtj_b200.PhysicalConstants.gyromagnetic_ratio() == 220880.0


def test_error_contract(ref, on_oracle):
    """Same exception types / message substrings as the reference (SURVEY 8a note 11)."""
    for m in (ref, cmtj_b200):
        d = helpers._demag(m)
        l = m.Layer("free", m.CVector(0, 0, 1), m.CVector(1, 0, 0), 1.0, 1e-9, 1e-14, d, 0.01)
        with pytest.raises(ValueError, match="must have a pinning"):
            m.Junction([l], 100, 200)
        with pytest.raises(ValueError, match="zero length"):
            m.Junction([])
        with pytest.raises(ValueError, match="unique ids"):
            m.Junction([l, l])
        with pytest.raises(RuntimeError, match="zero vector"):
            m.Layer("x", m.CVector(0, 0, 0), m.CVector(1, 0, 0), 1.0, 1e-9, 1e-14, d, 0.01)
        j = m.Junction([l])
        with pytest.raises(RuntimeError, match="cannot be larger than write frequency"):
            j.runSimulation(1e-9, 1e-12, 1e-13)
        with pytest.raises(RuntimeError, match="Failed to find a layer"):
            j.setLayerAnisotropyDriver("nope", m.ScalarDriver.getConstantDriver(1.0))
        with pytest.raises(RuntimeError, match="Start time cannot be later"):
            m.ScalarDriver.getStepDriver(0, 1, 2e-9, 1e-9)
        sot = m.Layer.createSOTLayer("s", m.CVector(0, 0, 1), m.CVector(1, 0, 0), 1.0, 1e-9, 1e-14, d, 0.01, 1.0, 1.0)
        with pytest.raises(RuntimeError, match="static SOT"):
            sot.setFieldLikeTorqueDriver(m.ScalarDriver.getConstantDriver(1.0))


def test_reference_is_ill_conditioned_for_euler_heun(ref):
    """Documents why GPU parity uses a sensitivity-scaled bound on some runs: the REFERENCE's own
    Euler-Heun trajectory of the tri-layer fixture amplifies a 1e-15 change of the initial
    magnetisation by > 1e9 within 2 ns, while RK4 on the same system stays below 1e-10."""

    def run(mode, eps):
        j = helpers.tri_layer_mtj(ref, mag=(0.1 * (1 + eps), 0.1, 1.0))
        j.runSimulation(2e-9, 1e-13, 1e-12, solverMode=getattr(ref.SolverMode, mode))
        return np.asarray(j.getLog()["free_mx"])

    d_eh = np.abs(run("EulerHeun", 0.0) - run("EulerHeun", 1e-15)).max()
    d_rk = np.abs(run("RK4", 0.0) - run("RK4", 1e-15)).max()
    assert d_eh > 1e-6
    assert d_rk < 1e-10


@pytest.mark.parametrize("mode", MODES + ["DormandPrince"])
@pytest.mark.parametrize("build", ["two_layer_mtj", "stt_layer_mtj", "sot_dynamic", "sot_static"])
def test_calculate_energies_bit_exact(ref, on_oracle, build, mode):
    """calculateEnergies=True (junction.hpp:1584-1613): the cached values of the LAST right-hand-side
    evaluation of each step, incl. the doubled "<id>_I" key for STT / SOT layers."""
    if mode == "DormandPrince" and build == "two_layer_mtj":
        pytest.skip("multi-layer Dormand-Prince diverges in the reference")
    logs = []
    for m in (ref, cmtj_b200):
        if build.startswith("sot"):
            j = helpers.sot_layer(m, dynamic=build == "sot_dynamic")
        else:
            j = getattr(helpers, build)(m)
        if build == "two_layer_mtj":
            j.setLayerAnisotropyDriver("free", m.ScalarDriver.getSineDriver(1e3, 300.0, 5e9, 0.1))
            j.setIECDriver("free", "bottom", m.ScalarDriver.getStepDriver(-1.78e-3, 1e-4, 1e-10, 3e-10))
            j.setLayerExternalFieldDriver("all", m.AxialDriver(m.ScalarDriver.getSineDriver(1e4, 2e3, 7e9, 0), m.NullDriver(), m.NullDriver()))
        dt = 1e-12 if mode == "DormandPrince" else 1e-13
        j.runSimulation(0.5e-9, dt, 10 * dt, False, True, getattr(m.SolverMode, mode))
        logs.append(j.getLog())
    lr, lo = helpers.log_arrays(logs[0]), helpers.log_arrays(logs[1])
    assert set(lr) == set(lo), sorted(set(lr) ^ set(lo))
    for k in lr:
        assert lr[k].shape == lo[k].shape, (k, lr[k].shape, lo[k].shape)
        assert np.array_equal(lr[k], lo[k]), (k, helpers.max_rel_err(lo[k], lr[k]))
    assert any(k.endswith("_Hanisx") for k in lo)
    if build != "two_layer_mtj":
        assert len(lo["free_I"]) == 2 * len(lo["time"])


def _custom_tweak(mod, j):
    """Python callbacks on a field component, the anisotropy and (SOT layer) a torque driver."""
    cb_h = lambda t: 2e3 * math.sin(2 * math.pi * 7e9 * t) + 1e12 * t  # noqa: E731
    cb_k = lambda t: 0.8e6 * (1 + 0.05 * math.cos(2 * math.pi * 3e9 * t))  # noqa: E731
    ids = j.getLayerIds()
    j.setLayerOerstedFieldDriver(ids[0], mod.AxialDriver(mod.ScalarDriver.getCustomDriver(cb_h), mod.NullDriver(),
                                                        mod.ScalarDriver.getConstantDriver(100.0)))
    j.setLayerAnisotropyDriver(ids[-1], mod.ScalarDriver.getCustomDriver(cb_k))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("build", ["single_layer_mtj", "two_layer_mtj"])
def test_custom_python_callback_drivers_bit_exact(ref, on_oracle, build, mode):
    """ScalarDriver.getCustomDriver in the dynamics (drivers.hpp:294-303): the reference calls the Python
    function at every stage time; the facade pre-evaluates it there (Junction._custom_table) and the
    restatement reads the table.  Same doubles in, bit-identical trajectories out."""
    jr, jo = _run_both(ref, getattr(helpers, build), T=0.3e-9, wf=5e-13, mode=mode, tweak=_custom_tweak)
    _assert_same_logs(jr, jo)


@pytest.mark.parametrize("mode", MODES)
def test_custom_python_callback_drivers_on_a_six_layer_junction_bit_exact(ref, on_oracle, mode):
    """the same for a junction beyond four layers (the lane-per-layer kernel's case on the GPU)"""
    jr, jo = _run_both(ref, helpers.multi_layer, T=0.2e-9, wf=4e-13, mode=mode, tweak=_custom_tweak, n=6)
    _assert_same_logs(jr, jo)


def test_custom_driver_errors(on_oracle):
    j = helpers.single_layer_mtj(cmtj_b200)
    j.setLayerCurrentDriver("free", cmtj_b200.ScalarDriver.getCustomDriver(lambda t: 1 / 0))
    with pytest.raises(RuntimeError, match="Error in Python callback"):
        j.runSimulation(1e-11, 1e-13, 1e-13)
    j = helpers.single_layer_mtj(cmtj_b200)
    j.setLayerAnisotropyDriver("free", cmtj_b200.ScalarDriver.getCustomDriver(lambda t: 1e5))
    with pytest.raises(NotImplementedError, match="fixed-step"):
        j.runSimulation(1e-11, 1e-13, 1e-13, solverMode=cmtj_b200.SolverMode.DormandPrince)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,strip", [(4, False), (5, True), (8, False)])
def test_many_layer_junctions_bit_exact(ref, on_oracle, n, strip, mode):
    """runMultiLayerSolver for L = 4, 5, 8 layers (junction.hpp:1682-1707): the restatement is generic in the
    layer count; on the GPU, L > 4 runs the lane-per-layer kernel (tests/test_gpu_parity.py)."""
    jr, jo = _run_both(ref, helpers.multi_layer, T=0.2e-9, wf=4e-13, mode=mode, n=n, strip=strip)
    _assert_same_logs(jr, jo)


@pytest.mark.parametrize("mode", MODES)
def test_two_layer_stt_junction_bit_exact(ref, on_oracle, mode):
    """Reference = bottom layer (referenceType BOTTOM, junction.hpp:727-740), constant and sine current."""
    jr, jo = _run_both(ref, helpers.stt_two_layer_mtj, T=0.5e-9, wf=5e-13, mode=mode)
    _assert_same_logs(jr, jo)

    def sine(mod, j):
        j.setLayerCurrentDriver("free", mod.ScalarDriver.getSineDriver(0, 5e10, 2e9, 0.0))

    jr, jo = _run_both(ref, helpers.stt_two_layer_mtj, T=0.5e-9, wf=5e-13, mode=mode, tweak=sine)
    _assert_same_logs(jr, jo)
