"""GPU parity of the coupled-system kernels (stack_kernel, group_block_kernel, group_grid_kernel)
against the oracle restatement (itself bit-exact to the reference, test_coupled_oracle_vs_ref.py).
Deterministic (RK4) runs: <= 1e-9 relative on every logged channel, with the same
conditioning-aware bound as test_gpu_parity.py."""
import math

import numpy as np
import pytest

import cmtj_b200
import helpers
import test_coupled_oracle_vs_ref as T
from parity_report import record

pytestmark = pytest.mark.gpu
TOL = 1e-9
EPS = 1e-12
m = cmtj_b200


def _three(fn):
    """CUDA, oracle, oracle with junction 0's m perturbed by EPS (sensitivity of the reference itself)."""
    out = []
    for backend, eps in ((None, 0.0), (helpers.oracle_backend(), 0.0), (helpers.oracle_backend(), EPS)):
        with helpers.facade_on(backend):
            out.append(fn(eps))
    return out


def _check(g, w, p, tol=TOL):
    assert set(g) == set(w)
    worst, worst_bound, widened = 0.0, tol, False
    for k in w:
        assert g[k].shape == w[k].shape, k
        if k == "time":
            assert np.array_equal(g[k], w[k])
            continue
        scale = max(np.max(np.abs(w[k])), 1e-300)
        err = np.abs(g[k] - w[k]) / scale
        sens = np.maximum.accumulate(np.abs(p[k] - w[k]) / scale, axis=0)
        bound = np.maximum(tol, 10 * sens)
        worst, worst_bound = max(worst, float(err.max())), max(worst_bound, float(bound.max()))
        widened = widened or bool((err > tol).any())
        bad = err > bound
        assert not bad.any(), (k, float(err.max()), float(sens.max()))
    record("coupled system", max_rel_err=worst, bound_1e9=tol, max_bound_applied=worst_bound, widened=widened)


def _perturb(j, eps):
    l0 = j.getLayerIds()[0]
    v = j.getLayer(l0).mag
    j.getLayer(l0).mag = m.CVector(v.x * (1 + eps) + eps, v.y, v.z)


def _collect(system, n):
    out = {}
    try:
        out.update({"S/" + k: v for k, v in helpers.log_arrays(system.getLog()).items()})
    except RuntimeError:
        pass
    for k in range(n):
        out.update({f"{k}/{name}": v for name, v in helpers.log_arrays(system.getLog(k)).items()})
    out["time"] = out.pop("0/time")
    return {k: v for k, v in out.items() if not k.endswith("/time")} | {"time": out["time"]}


@pytest.mark.parametrize("kind", ["SeriesStack", "ParallelStack"])
@pytest.mark.parametrize("mode", ["kcl", "delayed", "instant"])
def test_stt_stack(kind, mode):
    def fn(eps):
        js = [T._stt_two_layer(m), T._stt_two_layer(m, (0.5, 0.1, 0.8)), T._stt_two_layer(m, (-0.3, 0.3, 0.9)),
              T._stt_two_layer(m, (0.0, 0.6, 0.6)), T._stt_two_layer(m, (0.2, 0.2, 0.9))]
        if eps:
            _perturb(js[0], eps)
        s = getattr(m.stack, kind)(js, "free", "bottom", 0.1, mode == "kcl")
        s.setCouplingStrength([0.25, -0.1, 0.3, 0.05])
        if mode == "instant":
            s.setDelayed(False)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, 2e10, 5e9, 0.0))
        s.runSimulation(0.5e-9, 1e-13, 5e-13)
        return _collect(s, 5)

    g, w, p = _three(fn)
    _check(g, w, p)


@pytest.mark.parametrize("kcl", [True, False])
def test_single_layer_heterogeneous_stack(kcl):
    def fn(eps):
        js = T._hetero_pair(m) + T._hetero_pair(m)
        if eps:
            _perturb(js[0], eps)
        s = m.stack.SeriesStack(js, "free", "bottom", 0.2, kcl)
        s.setCouplingStrength(0.3)
        s.setCoupledCurrentDriver(m.ScalarDriver.getPulseDriver(1e10, 4e10, 1e-10, 0.4))
        s.setExternalFieldDriver(m.AxialDriver(20e3, 5e3, -3e3))
        s.runSimulation(0.4e-9, 1e-13, 4e-13)
        return _collect(s, 4)

    g, w, p = _three(fn)
    _check(g, w, p)


def test_stack_of_64_batched():
    """BASELINE config 5 shape: a series stack of 64 MTJs; 6 copies in one launch with
    per-instance current amplitude; every copy equals the oracle."""
    rng = np.random.default_rng(5)
    mags = rng.normal(0, 0.3, (64, 3)) + np.array([0.2, 0.1, 1.0])
    amps = np.linspace(1e10, 6e10, 6)

    def fn(eps):
        js = [T._stt_two_layer(m, tuple(v)) for v in mags]
        if eps:
            _perturb(js[0], eps)
        s = m.stack.SeriesStack(js, useKCL=True)
        s.setCouplingStrength(0.1)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(1e10, 1e10, 6e9, 0.0))
        out = s.run_batch(len(amps), 0.2e-9, 1e-13, 4e-13, current={"amplitude": amps})
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)


@pytest.mark.parametrize("inter", [None, "nullDipoleInteraction", "computeDipoleInteractionNoumra"])
def test_group_small(inter):
    def fn(eps):
        g = T._group(m, m.reservoir, n=7, inter=inter)
        if eps:
            _perturb(g.junctionList[0], eps)
        g.runSimulation(0.3e-9, 1e-13, 3e-13)
        return _collect(g, 7)

    g, w, p = _three(fn)
    _check(g, w, p)


def _lattice_group(n_side, eps=0.0, pitch=100e-9, seed=11):
    rng = np.random.default_rng(seed)
    js, pos = [], []
    for k in range(n_side * n_side):
        mag = rng.normal(0, 0.2, 3) + np.array([0.0, 0.0, 1.0])
        # Ms = 1.2 T: dipole fields of a few kA/m at this pitch, i.e. a regular (non-chaotic) array,
        # so the 1e-9 bar is meaningful for every element
        l = m.Layer("free", m.CVector(*mag), m.CVector(0, 0, 1), 1.2, 2e-9, math.pi * (20e-9) ** 2, [m.CVector(0, 0, 0)] * 3, 0.01)
        j = m.Junction([l])
        j.setLayerAnisotropyDriver("free", m.ScalarDriver.getConstantDriver(0.3e6))
        js.append(j)
        pos.append(m.CVector(pitch * (k % n_side), pitch * (k // n_side), 0.0))
    if eps:
        _perturb(js[0], eps)
    return m.reservoir.GroupInteraction(pos, js, "free")


def test_group_grid_kernel_lattice():
    """20x20 lattice = 400 elements > one block: cooperative grid kernel, one grid barrier per step."""

    def fn(eps):
        g = _lattice_group(20, eps)
        out = g.run_batch(1, 0.05e-9, 1e-13, 1e-12)
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)
    assert np.ptp(w["free_mx"][:, 0, 200]) > 1e-3  # the array does move


def test_group_batch_of_blocks():
    """several copies of a 6x6 array in one launch (block-per-array kernel) with per-element Hext."""
    H = np.random.default_rng(2).normal(0, 5e3, (5, 36))

    def fn(eps):
        g = _lattice_group(6, eps)
        out = g.run_batch(5, 0.1e-9, 1e-13, 1e-12, overrides={("free", "Hext_x", "constantValue"): H})
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)


def test_thermal_stack_and_group_run_euler_heun():
    """300 K: getSolver turns the requested RK4 into Euler-Heun for every junction; different RNG
    streams, so only sanity is checked here (|m| = 1, finite, moved)."""
    js = [helpers.stt_layer_mtj(m, temperature=300.0) for _ in range(3)]
    s = m.stack.SeriesStack(js)
    s.setCouplingStrength(0.1)
    s.setCoupledCurrentDriver(m.ScalarDriver.getConstantDriver(-1e10))
    s.runSimulation(0.2e-9, 1e-13, 1e-12)
    log = helpers.log_arrays(s.getLog(1))
    n = np.sqrt(log["free_mx"] ** 2 + log["free_my"] ** 2 + log["free_mz"] ** 2)
    assert np.allclose(n, 1.0, atol=1e-12) and np.ptp(log["free_mx"]) > 0
    g = T._group(m, m.reservoir, n=4, temperature=300.0)
    g.runSimulation(0.1e-9, 1e-13, 1e-12)
    lg = helpers.log_arrays(g.getLog(2))
    assert np.all(np.isfinite(lg["free_mz"])) and len(lg["time"]) == 100


@pytest.fixture
def general_variant(monkeypatch):
    """CMTJB_FORCE_VARIANT=full: the fully general term set (F_ALL, models read from the device array)
    instead of the lean / uniform-model variants the systems above select."""
    monkeypatch.setenv("CMTJB_FORCE_VARIANT", "full")


@pytest.mark.parametrize("mode", ["kcl", "delayed", "instant"])
def test_stack_general_variant(general_variant, mode):
    def fn(eps):
        js = [T._stt_two_layer(m), T._stt_two_layer(m, (0.5, 0.1, 0.8)), T._stt_two_layer(m, (-0.3, 0.3, 0.9))]
        if eps:
            _perturb(js[0], eps)
        s = m.stack.SeriesStack(js, "free", "bottom", 0.1, mode == "kcl")
        s.setCouplingStrength([0.25, -0.1])
        if mode == "instant":
            s.setDelayed(False)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, 2e10, 5e9, 0.0))
        s.runSimulation(0.3e-9, 1e-13, 5e-13)
        return _collect(s, 3)

    g, w, p = _three(fn)
    _check(g, w, p)


def test_group_general_variant(general_variant):
    def fn(eps):
        g = _lattice_group(20, eps)
        out = g.run_batch(1, 0.03e-9, 1e-13, 1e-12)
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)

    def fn2(eps):
        g = T._group(m, m.reservoir, n=7, inter=None)
        if eps:
            _perturb(g.junctionList[0], eps)
        g.runSimulation(0.2e-9, 1e-13, 3e-13)
        return _collect(g, 7)

    g, w, p = _three(fn2)
    _check(g, w, p)


@pytest.mark.parametrize("n", [3, 8, 32, 64, 96])
def test_stack_kcl_reduction_shapes(n):
    """The KCL sum takes the shuffle path for warp-aligned stacks (power of two <= 32, multiples of 32)
    and the block scan otherwise; a batch of 3 stacks so that blocks hold several / partial stacks."""
    rng = np.random.default_rng(n)
    mags = rng.normal(0, 0.3, (n, 3)) + np.array([0.2, 0.1, 1.0])

    def fn(eps):
        js = [T._stt_two_layer(m, tuple(v)) for v in mags]
        if eps:
            _perturb(js[0], eps)
        s = m.stack.SeriesStack(js, useKCL=True)
        s.setCouplingStrength(0.05)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(1e10, 1e10, 6e9, 0.0))
        out = s.run_batch(3, 0.05e-9, 1e-13, 5e-13, current={"amplitude": np.array([1e10, 2e10, 3e10])})
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)


def test_group_two_kernel_experiment_matches_the_oracle(monkeypatch):
    """CMTJB_GROUP_TWO_KERNEL=1: pair-sum kernel + per-element step kernel per time step (off by default)."""
    monkeypatch.setenv("CMTJB_GROUP_TWO_KERNEL", "1")

    def fn(eps):
        g = _lattice_group(20, eps)
        out = g.run_batch(1, 0.03e-9, 1e-13, 1e-12)
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)


@pytest.mark.parametrize("n", [7, 144])
def test_group_custom_interaction_table(n):
    """A user-supplied interaction callable (probed into a tensor table, CMTJB_INT_TABLE; group_pair_table_kernel +
    group_step_kernel) against the oracle evaluating the same table."""

    def halved(r1, r2, l1, l2):
        return m.reservoir.computeDipoleInteraction(r1, r2, l1, l2) * 0.5

    def fn(eps):
        if n == 7:
            g = T._group(m, m.reservoir, n=7)
            if eps:
                _perturb(g.junctionList[0], eps)
        else:
            g = _lattice_group(12, eps)  # 144 elements: several warp tiles and partner slices
        g.setInteractionFunction(halved)
        out = g.run_batch(1, 0.03e-9, 1e-13, 1e-12)
        out.pop("m_state")
        return out

    g, w, p = _three(fn)
    _check(g, w, p)
