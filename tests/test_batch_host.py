"""Host logic of the batched sweep API (JunctionBatch), checked on CPU with the oracle
standing in for the device library: instance i of a batch must equal a sequential
reference run with instance i's parameters (same initial m, same drivers)."""
import numpy as np
import pytest

import cmtj_b200
import helpers


def _ref_run(ref, H, m0, amp=None, mode="RK4", T=0.4e-9, wf=1e-12):
    j = helpers.two_layer_mtj(ref)
    j.setLayerExternalFieldDriver("all", ref.AxialDriver(*[float(h) for h in H]))
    oe = ref.ScalarDriver.getSineDriver(0.0, 50.0 if amp is None else float(amp), 9e9, 0.1)
    j.setLayerOerstedFieldDriver("free", ref.AxialDriver(ref.NullDriver(), ref.NullDriver(), oe))
    j.setLayerMagnetisation("free", ref.CVector(*m0[0]))
    j.setLayerMagnetisation("bottom", ref.CVector(*m0[1]))
    j.runSimulation(T, 1e-13, wf, solverMode=getattr(ref.SolverMode, mode))
    return helpers.log_arrays(j.getLog()), j


@pytest.mark.parametrize("mode", ["RK4", "Heun", "EulerHeun"])
def test_batch_equals_sequential_reference(ref, on_oracle, mode):
    rng = np.random.default_rng(0)
    B = 5
    H = np.stack([np.linspace(-300e3, 300e3, B), np.full(B, 10e3), rng.normal(0, 1e3, B)], axis=1)
    m_free = rng.normal(0, 1, (B, 3)) + np.array([0, 0, 3.0])
    m_bot = rng.normal(0, 1, (B, 3)) + np.array([3.0, 0, 0])
    amps = np.linspace(10.0, 500.0, B)

    j = helpers.two_layer_mtj(cmtj_b200)
    oe = cmtj_b200.ScalarDriver.getSineDriver(0.0, 50.0, 9e9, 0.1)
    j.setLayerOerstedFieldDriver("free", cmtj_b200.AxialDriver(cmtj_b200.NullDriver(), cmtj_b200.NullDriver(), oe))
    b = cmtj_b200.JunctionBatch(j, B)
    b.set_external_field("all", H)
    b.set_magnetisation("free", m_free)
    b.set_magnetisation("bottom", m_bot)
    b.set_override("free", "Hoe_z", "amplitude", amps)
    log = b.run(0.4e-9, 1e-13, 1e-12, solverMode=getattr(cmtj_b200.SolverMode, mode))
    assert log["free_mx"].shape == (400, B)
    for i in range(B):
        want, jr = _ref_run(ref, H[i], (m_free[i], m_bot[i]), amps[i], mode)
        got = log.instance(i)
        assert set(got) == set(want)
        for k in want:
            assert np.array_equal(got[k], want[k]), (i, k)
        assert np.array_equal(b.get_magnetisation("free")[i], jr.getLayerMagnetisation("free").tolist())


def test_batch_state_carries_and_log_none(on_oracle):
    j = helpers.two_layer_mtj(cmtj_b200)
    b = cmtj_b200.JunctionBatch(j, 3)
    b.set_external_field("all", np.array([[1e5, 0, 0], [0, 1e5, 0], [0, 0, 1e5]]))
    full = b.run(0.2e-9, 1e-13, 1e-13)
    m_after = b.get_magnetisation("free")
    assert np.array_equal(m_after[:, 0], full["free_mx"][-1])
    # second leg continues from the carried state; LOG_NONE returns no samples but moves m
    empty = b.run(0.1e-9, 1e-13, 1e-13, log=False)
    assert len(empty) == 0
    assert not np.array_equal(b.get_magnetisation("free"), m_after)
    # same thing in one go
    b2 = cmtj_b200.JunctionBatch(j, 3)
    b2.set_external_field("all", np.array([[1e5, 0, 0], [0, 1e5, 0], [0, 0, 1e5]]))
    b2.run(0.2e-9, 1e-13, 1e-13, log=False)
    b2.run(0.1e-9, 1e-13, 1e-13, log=False)
    assert np.array_equal(b2.get_magnetisation("free"), b.get_magnetisation("free"))


def test_batch_argument_checks(on_oracle):
    j = helpers.two_layer_mtj(cmtj_b200)
    b = cmtj_b200.JunctionBatch(j, 4)
    with pytest.raises(ValueError):
        b.set_override("free", "K1", "constantValue", np.zeros(3))
    with pytest.raises(RuntimeError, match="Failed to find a layer"):
        b.set_override("nope", "K1", "constantValue", np.zeros(4))
    with pytest.raises(RuntimeError, match="zero vector"):
        b.set_magnetisation("free", np.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="cannot be larger than write frequency"):
        b.run(1e-9, 1e-12, 1e-13)
    empty = cmtj_b200.JunctionBatch(j, 0)
    assert empty.run(1e-10, 1e-13, 1e-13)["free_mx"].shape == (1000, 0)


def test_cpulist_parsing_and_affinity_helpers_are_safe_without_a_gpu():
    from cmtj_b200 import _numa

    assert _numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _numa._parse_cpulist("") == []
    # no CUDA device here: nothing is known, nothing is changed
    import os

    before = os.sched_getaffinity(0)
    assert _numa.bind_host_to_gpu(0) is None or os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_kernel_variant_selection_rules():
    """Which compiled variant covers which model (cabi.cu: needed_features per layer + pick_variant).  Host
    logic only: the library is dlopen'ed, no CUDA call is made."""
    import cmtj_b200 as cm

    dt = 1e-13
    # S1: two-layer PIMM sweep, per-instance field, step pulse on all layers -> lean IEC + table
    j = helpers.two_layer_mtj(cm)
    j.setLayerOerstedFieldDriver("all", cm.AxialDriver(cm.NullDriver(), cm.NullDriver(), cm.ScalarDriver.getStepDriver(0, 50, 0, 3 * dt)))
    b = cm.JunctionBatch(j, 4)
    b.set_external_field("all", np.zeros((4, 3)))
    assert b.selected_variant(1e-9, dt, 1e-11) == "CMTJB_RK4/L2/F_IEC | F_DYN"
    # no time-dependent driver: the lean variant
    b0 = cm.JunctionBatch(helpers.two_layer_mtj(cm), 4)
    assert b0.selected_variant(1e-9, dt, 1e-11) == "CMTJB_RK4/L2/F_IEC"
    # per-instance frequency: the shared stage table cannot hold it -> the fast-field variant, or the general
    # inline-evaluating one when overrides reach beyond that driver
    j = helpers.two_layer_mtj(cm)
    j.setLayerOerstedFieldDriver("all", cm.AxialDriver(cm.NullDriver(), cm.ScalarDriver.getSineDriver(0, 50, 5e9, 0), cm.NullDriver()))
    b = cm.JunctionBatch(j, 4)
    b.set_override("all", "Hoe_y", "frequency", np.linspace(1e9, 4e9, 4))
    assert b.selected_variant(1e-9, dt, 1e-11) == "CMTJB_RK4/L2/F_IEC | F_FAST"  # one field driver per layer: resolved once
    b.set_override("free", "K1", "constantValue", np.linspace(1e3, 2e3, 4))       # a per-instance SCALAR slot as well
    assert "F_ALL" in b.selected_variant(1e-9, dt, 1e-11) and "F_FAST" not in b.selected_variant(1e-9, dt, 1e-11)
    # CIMS shape: torque drivers and per-instance amplitudes on the free layer only -> split variant
    j = helpers.tri_layer_mtj(cm)
    j.setLayerFieldLikeTorqueDriver("free", cm.stepDriver(0, 1e3, 0, 1e-9))
    j.setLayerDampingLikeTorqueDriver("free", cm.stepDriver(0, 16e3, 0, 1e-9))
    j.getLayer("free").setReferenceLayer(cm.CVector(0, 1, 0))
    b = cm.JunctionBatch(j, 4)
    b.set_override("free", "dampingLike", "amplitude", np.ones(4))
    b.set_external_field("all", np.zeros((4, 3)))
    v = b.selected_variant(2e-9, 1e-12, 1e-11, cm.SolverMode.Heun)
    assert v.startswith("CMTJB_HEUN/L3/") and "CMTJB_SPLIT(1)" in v, v
    # the same drive on the BOTTOM layer: no split variant names that layer -> a plain variant
    j = helpers.tri_layer_mtj(cm)
    j.setLayerFieldLikeTorqueDriver("bottom", cm.stepDriver(0, 1e3, 0, 1e-9))
    j.getLayer("bottom").setReferenceLayer(cm.CVector(0, 1, 0))
    v = cm.JunctionBatch(j, 4).selected_variant(2e-9, 1e-12, 1e-11, cm.SolverMode.Heun)
    assert "SPLIT" not in v, v
    # two-layer STT junction: constant current, sine current (batch-uniform waveform), at temperature
    sv = lambda j, *a: cm.JunctionBatch(j, 4).selected_variant(1e-9, dt, 1e-11, *a)  # noqa: E731
    assert sv(helpers.stt_two_layer_mtj(cm)) == "CMTJB_RK4/L2/F_IEC | F_STT"
    assert sv(helpers.stt_two_layer_mtj(cm, current=cm.sineDriver(0, 5e10, 2e9, 0))) == "CMTJB_RK4/L2/F_IEC | F_STT | F_DYNS"
    assert sv(helpers.stt_two_layer_mtj(cm, temperature=300.0)) == "CMTJB_EULER_HEUN/L2/F_IEC | F_STT | F_THERMAL"
    assert sv(helpers.stt_two_layer_mtj(cm, temperature=300.0), cm.SolverMode.Heun) == "CMTJB_HEUN/L2/F_IEC | F_STT | F_THERMAL"
    # thermal single layer -> Euler-Heun lean thermal variant; many layers -> the lane-per-layer kernel
    assert "F_THERMAL" in cm.JunctionBatch(helpers.stt_layer_mtj(cm, temperature=300.0), 4).selected_variant(1e-9, dt, 1e-11)
    assert cm.JunctionBatch(helpers.multi_layer(cm, 6), 4).selected_variant(1e-9, dt, 1e-11) == "wide/L6/F_ALL"
