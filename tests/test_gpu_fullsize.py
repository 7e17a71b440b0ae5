"""Parity at BASELINE.json's full sizes.  configs[1] (2-layer MTJ, 1e5 field points x 2e5 RK4 steps, log every
100 steps), configs[3] (3-layer CIMS grid, 2048 x 2048 x 20 000 Heun steps) and configs[4] (64 x 64 dipole-coupled
macrospins; series stack of 64 MTJs, 4096 stacks x 2000 steps) through size-independent properties:

* every logged magnetisation has unit length;
* batch-position invariance: an instance integrated inside the 1e5-instance sweep (SMSP-balanced persistent
  kernel, time-window specialised launches) is BIT-IDENTICAL to the same instance integrated in a batch of 48;
* those 48 instances, over the FULL 2e5 steps, agree with the CPU oracle within the north_star bound
  (1e-9 relative, widened only where the reference arithmetic itself amplifies a 1e-12 perturbation)."""
import numpy as np
import pytest
import torch

import cmtj_b200
import helpers
import workloads as wl
from parity_report import compare_with_sensitivity, record

pytestmark = pytest.mark.gpu
DT, STEPS, WE, B = 1e-13, 200000, 100, 100000
T, WF = STEPS * DT * (1 + 1e-9), WE * DT * (1 + 1e-9)


def _junction():
    j = helpers.two_layer_mtj(cmtj_b200, mag=(0, 0, 1))
    oe = cmtj_b200.ScalarDriver.getStepDriver(0, 50, 0, 3 * DT)
    j.setLayerOerstedFieldDriver("all", cmtj_b200.AxialDriver(cmtj_b200.NullDriver(), cmtj_b200.NullDriver(), oe))
    return j


def _inputs(start="hard_axis"):
    """hard_axis: every point released 1e-3 from z, an unstable equilibrium (the reference itself amplifies 1e-12 to
    order one over the run: the ill-conditioned case); along_field: saturated along +-x like bench.py (well conditioned)."""
    rng = np.random.default_rng(0)
    H = np.zeros((B, 3))
    H[:, 0] = np.linspace(-500e3, 500e3, B)
    base = np.array([[0, 0, 1.0]]) if start == "hard_axis" else np.sign(H[:, :1]) * np.array([[1.0, 0, 0]])
    mf = rng.normal(0, 1e-3, (B, 3)) + base
    mb = rng.normal(0, 1e-3, (B, 3)) + base
    return H, mf, mb


def _run(H, mf, mb, device_resident, eps=0.0, axis=0):
    b = cmtj_b200.JunctionBatch(_junction(), H.shape[0])
    b.set_external_field("all", H)
    mf = mf.copy()
    mf[:, axis] = mf[:, axis] * (1 + eps) + eps
    b.set_magnetisation("free", mf)
    b.set_magnetisation("bottom", mb)
    if device_resident:
        return b.run_device(T, DT, WF)
    return b.run(T, DT, WF)


@pytest.mark.parametrize("start", ["hard_axis", "along_field"])
def test_s1_full_size_properties_and_sampled_oracle_parity(start):
    H, mf, mb = _inputs(start)
    big = _run(H, mf, mb, True)
    torch.cuda.synchronize()
    for lid in ("free", "bottom"):
        n2 = big[f"{lid}_mx"] ** 2 + big[f"{lid}_my"] ** 2 + big[f"{lid}_mz"] ** 2
        assert float((n2 - 1).abs().max()) < 1e-12
    sel = np.sort(np.random.default_rng(7).choice(B, size=48, replace=False))
    keys = [k for k in big if k not in ("time", "_log")]
    got = {k: big[k][:, torch.from_numpy(sel).cuda()].cpu().numpy() for k in keys}
    del big
    small = _run(H[sel], mf[sel], mb[sel], False)
    for k in keys:
        assert np.array_equal(got[k], small[k]), k  # position in the batch / kernel form do not matter
    with helpers.facade_on(helpers.oracle_backend()):
        want = _run(H[sel], mf[sel], mb[sel], False)
        pert = [_run(H[sel], mf[sel], mb[sel], False, eps=1e-12, axis=a) for a in range(3)]
    _compare(f"S1 full size ({start} start): 48 of 1e5 instances, all 2e5 RK4 steps", keys, got, want, pert)


_compare = compare_with_sensitivity


def test_s4_cims_full_grid_properties_and_sampled_oracle_parity():
    """BASELINE configs[3] at its named size: 2048 x 2048 current x field grid, 3-layer stack, 20 000 Heun steps with
    SOT step drivers (per-instance amplitudes).  |m| = 1 on every final state; 64 instances drawn from the grid are
    bit-identical to the same 64 integrated as their own batch; those 64, logged every 100 steps over the whole
    run, agree with the oracle."""
    n, steps, dt = wl.S4_GRID, wl.S4_STEPS, wl.S4_DT
    T = wl.exact(steps, dt)
    mode = cmtj_b200.SolverMode.Heun
    big = wl.s4_batch(cmtj_b200, n)
    big.run_device(T, dt, T, solverMode=mode, log=False)
    torch.cuda.synchronize()
    mfin = big._m_dev  # [9, B]
    for l in range(3):
        n2 = (mfin[3 * l : 3 * l + 3] ** 2).sum(dim=0)
        assert float((n2 - 1).abs().max()) < 1e-12
    sel = np.sort(np.random.default_rng(11).choice(n * n, size=64, replace=False))
    got_final = mfin[:, torch.from_numpy(sel).cuda()].cpu().numpy()
    del big, mfin
    small = wl.s4_batch(cmtj_b200, n, select=sel)
    small.run(T, dt, T, solverMode=mode, log=False)
    assert np.array_equal(small.m_state, got_final)  # position in the batch / kernel form do not matter
    wf = wl.exact(100, dt)
    logs = []
    for be, axis in ((None, None), (helpers.oracle_backend(), None), (helpers.oracle_backend(), 0), (helpers.oracle_backend(), 1),
                     (helpers.oracle_backend(), 2)):
        with helpers.facade_on(be):
            b = wl.s4_batch(cmtj_b200, n, select=sel)
            if axis is not None:
                mf = b.get_magnetisation("free")
                mf[:, axis] = mf[:, axis] * (1 + 1e-12) + 1e-12
                b.set_magnetisation("free", mf, normalise=False)
            logs.append(b.run(T, dt, wf, solverMode=mode))
    got, want = logs[:2]
    keys = [k for k in want if k != "time"]
    _compare("S4 full size: 64 of 2048^2 instances, all 20000 Heun steps", keys, got, want, logs[2:])


def test_s5_lattice_64x64_against_the_oracle():
    """BASELINE configs[4]: 4096 dipole-coupled macrospins (GroupInteraction, reservoir.hpp:129-152), cooperative grid
    kernel, 50 RK4 steps with every step logged against the oracle's O(N^2) loop (1.7e7 pair fields per step)."""
    steps = 50
    logs = []
    for be in (None, helpers.oracle_backend()):
        with helpers.facade_on(be):
            g = wl.s5_lattice(cmtj_b200, wl.S5G_SIDE)
            out = g.run_batch(1, wl.exact(steps), 1e-13, 1e-13)
            out.pop("m_state")
            logs.append(out)
    got, want = logs
    worst = 0.0
    for k in want:
        if k == "time":
            continue
        assert got[k].shape == want[k].shape == (steps, 1, wl.S5G_SIDE**2)
        worst = max(worst, helpers.max_rel_err(got[k], want[k]))
    record("S5 full size: 64x64 dipole lattice, 50 RK4 steps, every element", max_rel_err=worst, bound_1e9=1e-9, max_bound_applied=1e-9, widened=False)
    assert worst <= 1e-9, worst
    assert np.ptp(want["free_mx"][:, 0, 2000]) > 1e-6  # the array does move


def test_s5_series_stack_of_64_full_run_against_the_oracle():
    """BASELINE configs[4], electrical half: 4096 series stacks of 64 STT MTJs (KCL), per-stack current amplitude,
    2000 RK4 steps.  Final states of 8 stacks drawn from the batch are bit-identical to the same stacks run as their
    own batch of 8, and those 8 agree with the oracle over the whole run (stack resistance, currents, couplings and
    every junction's m, logged every 20 steps)."""
    steps, n_stacks, N = wl.S5S_STEPS, wl.S5S_STACKS, wl.S5S_JUNCTIONS
    T, wf = wl.exact(steps), wl.exact(20)
    amps = wl.s5_stack_amplitudes(n_stacks)
    big = wl.s5_stack(cmtj_b200).run_batch(n_stacks, T, 1e-13, T, current={"amplitude": amps})
    mbig = big["m_state"]  # [3, n_stacks * N]
    assert np.allclose((mbig**2).sum(axis=0), 1.0, atol=1e-12)
    sel = np.sort(np.random.default_rng(3).choice(n_stacks, size=8, replace=False))
    cols = (sel[:, None] * N + np.arange(N)[None, :]).ravel()
    logs = []
    for be, eps in ((None, 0.0), (helpers.oracle_backend(), 0.0), (helpers.oracle_backend(), 1e-12)):
        with helpers.facade_on(be):
            s = wl.s5_stack(cmtj_b200)
            if eps:
                l0 = s.junctionList[0].getLayer("free")
                l0.mag = cmtj_b200.CVector(l0.mag.x * (1 + eps) + eps, l0.mag.y, l0.mag.z)
            logs.append(s.run_batch(8, T, 1e-13, wf, current={"amplitude": amps[sel]}))
    got, want, pert = logs
    assert np.array_equal(got.pop("m_state"), mbig[:, cols])
    want.pop("m_state")
    pert.pop("m_state")
    keys = [k for k in want if k != "time"]
    for k in keys:
        assert want[k].shape[1] == 8, (k, want[k].shape)  # [samples, stack, ...]
    _compare("S5 full size: 8 of 4096 series stacks of 64 MTJs, all 2000 RK4 steps", keys, got, want, pert)

