"""Pins the oracle's Stack / GroupInteraction restatements (oracle/llgs_oracle.cpp, driven through
cmtj_b200.stack / cmtj_b200.reservoir) BIT-FOR-BIT against the unmodified reference extension
(oracle/_ref).  CPU only."""
import math

import numpy as np
import pytest

import cmtj_b200
import helpers


def _mods(ref):
    return ((ref, ref.stack, ref.reservoir), (cmtj_b200, cmtj_b200.stack, cmtj_b200.reservoir))


def _same(a, b):
    a, b = helpers.log_arrays(a), helpers.log_arrays(b)
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), (k, helpers.max_rel_err(b[k], a[k]))


def _hetero_pair(m):
    """tests/test_stack.py:141-180 of the reference: two different single-layer junctions."""
    demag = helpers._demag(m)
    out = []
    for mag, anis, Ms, K in (((0, 0, 1), (1, 0, 0), 1.65, 1.05e3), ((0.3, 0.2, 0.9), (1, 0.1, 0.3), 1.35, 1.05e2)):
        l = m.Layer("free", m.CVector(*mag), m.CVector(*anis), Ms, 3e-9, helpers.SURFACE, demag, 0.005)
        l.setReferenceLayer(m.CVector(1, 0, 0))
        j = m.Junction([l], 100, 200)
        j.setLayerAnisotropyDriver("free", m.ScalarDriver.getConstantDriver(K))
        out.append(j)
    return out


@pytest.mark.parametrize("kind", ["SeriesStack", "ParallelStack"])
@pytest.mark.parametrize("kcl", [True, False])
@pytest.mark.parametrize("build", ["single", "two", "hetero"])
def test_stack_bit_exact(ref, on_oracle, kind, kcl, build):
    stacks = []
    for m, st, _ in _mods(ref):
        if build == "hetero":
            js = _hetero_pair(m) + _hetero_pair(m)[:1]
        else:
            b = helpers.single_layer_mtj if build == "single" else helpers.two_layer_mtj
            js = [b(m), b(m, mag=(0.4, -0.1, 0.8)), b(m, mag=(-0.2, 0.5, 0.7))]
        if build != "two":
            # a plain Layer has no STT term; make the current matter for one junction
            pass
        s = getattr(st, kind)(js, "free", "bottom", 0.3, kcl)
        s.setCouplingStrength([0.3, -0.2])
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(1e10, 5e9, 7e9, 0.2))
        s.setExternalFieldDriver(m.AxialDriver(20e3, 5e3, -3e3))
        s.runSimulation(0.3e-9, 1e-13, 3e-13)
        stacks.append(s)
    r, o = stacks
    _same(r.getLog(), o.getLog())
    for k in range(3):
        _same(r.getLog(k), o.getLog(k))
        assert r.getMagnetisation(k, "free").tolist() == o.getMagnetisation(k, "free").tolist()


def _stt_two_layer(m, mag=(0.1, 0.2, 1.0)):
    """2-layer junction whose layers actually feel the stack current (STT, reference = neighbour)."""
    d = helpers._demag(m)
    a = m.Layer.createSTTLayer("free", m.CVector(*mag), m.CVector(0, 0, 1), 1.2, 1.5e-9, 1e-14, d, 0.02, 1.3, 0.1, 0.7)
    b = m.Layer.createSTTLayer("bottom", m.CVector(1, 0.1, 0.1), m.CVector(1, 0, 0), 1.5, 3e-9, 1e-14, d, 0.02, 1.3, 0.0, 0.7)
    a.setReferenceLayer(m.Reference.bottom)
    b.setReferenceLayer(m.Reference.top)
    j = m.Junction([a, b], 100, 200)
    j.setLayerAnisotropyDriver("free", m.ScalarDriver.getConstantDriver(4e5))
    j.setLayerAnisotropyDriver("bottom", m.ScalarDriver.getConstantDriver(1e5))
    j.setIECDriver("free", "bottom", m.ScalarDriver.getConstantDriver(2e-5))
    return j


@pytest.mark.parametrize("kind", ["SeriesStack", "ParallelStack"])
@pytest.mark.parametrize("mode", ["kcl", "delayed", "instant"])
def test_stt_stack_bit_exact(ref, on_oracle, kind, mode):
    stacks = []
    for m, st, _ in _mods(ref):
        js = [_stt_two_layer(m), _stt_two_layer(m, (0.5, 0.1, 0.8)), _stt_two_layer(m, (-0.3, 0.3, 0.9)), _stt_two_layer(m, (0.0, 0.6, 0.6))]
        s = getattr(st, kind)(js, "free", "bottom", 0.1, mode == "kcl")
        s.setCouplingStrength(0.25)
        if mode == "instant":
            s.setDelayed(False)
        s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, 2e10, 5e9, 0.0))
        s.runSimulation(0.4e-9, 1e-13, 2e-13)
        s.runSimulation(0.1e-9, 1e-13, 1e-13)  # logs append, state carries
        stacks.append(s)
    r, o = stacks
    _same(r.getLog(), o.getLog())
    for k in range(4):
        _same(r.getLog(k), o.getLog(k))


def _group(m, res, n=5, inter=None, Ms=1e6, temperature=None):
    rng = np.random.default_rng(3)
    js, pos = [], []
    for k in range(n):
        mag = rng.normal(0, 1, 3)
        l = m.Layer("free", m.CVector(*mag), m.CVector(0, 0, 1), Ms, 2e-9, math.pi * (20e-9) ** 2, [m.CVector(0, 0, 0)] * 3, 0.01 + 0.002 * k)
        j = m.Junction([l])
        j.setLayerAnisotropyDriver("free", m.ScalarDriver.getConstantDriver(0.9e6 + 1e4 * k))
        j.setLayerExternalFieldDriver("free", m.AxialDriver(1e3, 2e3 * k, -1e3))
        if temperature is not None:
            j.setLayerTemperatureDriver("free", m.ScalarDriver.getConstantDriver(temperature))
        js.append(j)
        pos.append(m.CVector(100e-9 * (k % 3), 110e-9 * (k // 3), 10e-9 * k))
    g = res.GroupInteraction(pos, js, "free")
    if inter is not None:
        g.setInteractionFunction(getattr(res, inter))
    return g


@pytest.mark.parametrize("inter", [None, "nullDipoleInteraction", "computeDipoleInteractionNoumra"])
def test_group_bit_exact(ref, on_oracle, inter):
    gs = []
    for m, _, res in _mods(ref):
        # Ms = 1e6 "Tesla" like the reference's own test (tests/test_group_interaction.py:11-27):
        # a dipole field large enough to matter at 100 nm pitch
        g = _group(m, res, inter=inter)
        g.runSimulation(0.2e-9, 1e-13, 2e-13)
        gs.append(g)
    r, o = gs
    for k in range(5):
        _same(r.getLog(k), o.getLog(k))
    moved = helpers.log_arrays(o.getLog(0))
    assert np.ptp(moved["free_mx"]) > 1e-6


def test_interaction_functions_match_reference(ref):
    for args in (((0, 0, 0), (100e-9, 0, 0)), ((1e-8, -2e-8, 3e-8), (-5e-8, 4e-8, 1e-9))):
        vals = []
        for m, _, res in _mods(ref):
            l1 = m.Layer("free", m.CVector(0.2, 0.3, 1), m.CVector(0, 0, 1), 1.2, 2e-9, 1e-15, [m.CVector(0, 0, 0)] * 3)
            l2 = m.Layer("free", m.CVector(-0.5, 0.1, 0.7), m.CVector(0, 0, 1), 0.8, 3e-9, 2e-15, [m.CVector(0, 0, 0)] * 3)
            p1, p2 = m.CVector(*args[0]), m.CVector(*args[1])
            vals.append(
                [f(p1, p2, l1, l2).tolist() for f in (res.nullDipoleInteraction, res.computeDipoleInteraction, res.computeDipoleInteractionNoumra)]
            )
        assert np.allclose(vals[0], vals[1], rtol=1e-13, atol=0)


def test_error_contract(ref, on_oracle):
    """Same exception types / message substrings as the reference (tests/test_stack.py:11-37,
    tests/test_group_interaction.py:30-47, :126-135)."""
    for m, st, res in _mods(ref):
        j = helpers.single_layer_mtj(m)
        for cls in (st.SeriesStack, st.ParallelStack):
            with pytest.raises(RuntimeError, match="must have at least 2 junctions"):
                cls([j])
            with pytest.raises(RuntimeError, match="must have at least 2 junctions"):
                cls([])
            with pytest.raises(RuntimeError, match="Asking for id of a non-existing junction!"):
                cls([j, j]).getLog(2)
            with pytest.raises(TypeError):
                cls([j, j]).getLog(-1)
            s = cls([j, j])
            with pytest.raises(RuntimeError, match="size of stack size - 1"):
                s.setCouplingStrength([0.1, 0.2, 0.3])
            with pytest.raises(RuntimeError, match="only supported for 2 layer stacks"):
                s.setDelayed(False)
            with pytest.raises(RuntimeError, match="cannot be larger than write frequency"):
                s.runSimulation(1e-9, 1e-12, 1e-13)
        l = m.Layer("free", m.CVector(0, 0, 1), m.CVector(0, 0, 1), 1.0, 2e-9, 1e-15, [m.CVector(0, 0, 0)] * 3)
        with pytest.raises(RuntimeError, match="non-classic"):
            st.SeriesStack([m.Junction([l]), m.Junction([l])])
        j1, j2 = m.Junction([l]), m.Junction([l])
        p1, p2 = m.CVector(0, 0, 0), m.CVector(1e-7, 0, 0)
        with pytest.raises(RuntimeError):
            res.GroupInteraction([p1], [j1, j2])
        with pytest.raises(RuntimeError):
            res.GroupInteraction([], [])
        with pytest.raises(RuntimeError):
            res.GroupInteraction([p1, p1], [j1, j2])
        g = res.GroupInteraction([p1, p2], [j1, j2])
        with pytest.raises(RuntimeError):
            g.getLog(2)
        with pytest.raises((RuntimeError, IndexError, TypeError)):
            g.getLog(-1)
        g.runSimulation(1e-11, 1e-12)
        assert len(g.getLog(0)["time"]) > 0
        g.clearLogs()
        with pytest.raises(KeyError):
            g.getLog(1)["time"]


def test_stack_batch_equals_single_runs(on_oracle):
    """run_batch instance i == a single stack with instance i's current amplitude."""
    m = cmtj_b200
    js = [_stt_two_layer(m), _stt_two_layer(m, (0.5, 0.1, 0.8)), _stt_two_layer(m, (-0.3, 0.3, 0.9))]
    amps = np.array([1e10, 2e10, 4e10, -3e10])
    s = m.stack.SeriesStack(js, useKCL=False)
    s.setCouplingStrength(0.2)
    s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, 1e10, 5e9, 0.0))
    out = s.run_batch(len(amps), 0.2e-9, 1e-13, 1e-13, current={"amplitude": amps})
    for i, a in enumerate(amps):
        s1 = m.stack.SeriesStack(js, useKCL=False)
        s1.setCouplingStrength(0.2)
        s1.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(3e10, float(a), 5e9, 0.0))
        s1.runSimulation(0.2e-9, 1e-13, 1e-13)
        log = helpers.log_arrays(s1.getLog())
        for k in ("Resistance", "I_0", "I_2", "C_0_1", "C_1_2"):
            assert np.array_equal(out[k][:, i], log[k]), (i, k)
        j1 = helpers.log_arrays(s1.getLog(1))
        assert np.array_equal(out["1/free_mx"][:, i], j1["free_mx"])


def _scaled_cutoff(res):
    """A user-written interaction in the only form the reference's Python API allows (its Layer exposes no `mag`): a
    composition of the built-in functions -- here Kanao's dipole field, halved, cut off beyond 250 nm."""

    def f(r1, r2, layer1, layer2):
        if (r1 - r2).length() > 2.5e-7:
            return res.nullDipoleInteraction(r1, r2, layer1, layer2)
        return res.computeDipoleInteraction(r1, r2, layer1, layer2) * 0.5

    return f


def test_custom_python_interaction_matches_reference(ref, on_oracle):
    """setInteractionFunction with an arbitrary callable (reservoir.hpp:181, cmtj.cpp:438): the reference calls it N^2
    times per step; here it is probed once per ordered pair and evaluated as a tabulated affine map (CMTJB_INT_TABLE).
    Not the same floating-point expression, so: 1e-10 relative instead of bit-exact."""
    gs = []
    for m, _, res in _mods(ref):
        g = _group(m, res, n=7, Ms=1.2)  # kA/m dipole fields: a regular array, so a tolerance means something
        g.setInteractionFunction(_scaled_cutoff(res))
        g.runSimulation(0.2e-9, 1e-13, 2e-13)
        gs.append(g)
    r, o = gs
    for k in range(7):
        a, b = helpers.log_arrays(r.getLog(k)), helpers.log_arrays(o.getLog(k))
        for name in a:
            assert helpers.max_rel_err(b[name], a[name]) <= 1e-10, (k, name)
    # the cutoff matters (element 6 is beyond it from element 0) and so does the interaction
    plain = _group(cmtj_b200, cmtj_b200.reservoir, n=7, Ms=1.2)
    plain.setInteractionFunction(cmtj_b200.reservoir.nullDipoleInteraction)
    plain.runSimulation(0.2e-9, 1e-13, 2e-13)
    assert helpers.max_rel_err(helpers.log_arrays(plain.getLog(0))["free_mx"], helpers.log_arrays(o.getLog(0))["free_mx"]) > 1e-6


def test_non_affine_interaction_is_rejected(on_oracle):
    m = cmtj_b200

    def f(r1, r2, layer1, layer2):
        return m.reservoir.computeDipoleInteraction(r1, r2, layer1, layer2) * (1.0 + layer2.mag.x ** 2)

    g = _group(m, m.reservoir, n=4)
    g.setInteractionFunction(f)
    with pytest.raises(NotImplementedError):
        g.runSimulation(1e-12, 1e-13, 1e-13)
    with pytest.raises(TypeError):
        g.setInteractionFunction(3)
