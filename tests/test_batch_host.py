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
