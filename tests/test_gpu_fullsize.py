"""Parity at BASELINE.json's full size (configs[1]: 2-layer MTJ, 1e5 field points x 2e5 RK4 steps, log every
100 steps) through size-independent properties:

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

pytestmark = pytest.mark.gpu
DT, STEPS, WE, B = 1e-13, 200000, 100, 100000
T, WF = STEPS * DT * (1 + 1e-9), WE * DT * (1 + 1e-9)


def _junction():
    j = helpers.two_layer_mtj(cmtj_b200, mag=(0, 0, 1))
    oe = cmtj_b200.ScalarDriver.getStepDriver(0, 50, 0, 3 * DT)
    j.setLayerOerstedFieldDriver("all", cmtj_b200.AxialDriver(cmtj_b200.NullDriver(), cmtj_b200.NullDriver(), oe))
    return j


def _inputs():
    rng = np.random.default_rng(0)
    H = np.zeros((B, 3))
    H[:, 0] = np.linspace(-500e3, 500e3, B)
    mf = rng.normal(0, 1e-3, (B, 3)) + [0, 0, 1.0]
    mb = rng.normal(0, 1e-3, (B, 3)) + [0, 0, 1.0]
    return H, mf, mb


def _run(H, mf, mb, device_resident, eps=0.0):
    b = cmtj_b200.JunctionBatch(_junction(), H.shape[0])
    b.set_external_field("all", H)
    mf = mf.copy()
    mf[:, 0] = mf[:, 0] * (1 + eps) + eps
    b.set_magnetisation("free", mf)
    b.set_magnetisation("bottom", mb)
    if device_resident:
        return b.run_device(T, DT, WF)
    return b.run(T, DT, WF)


def test_s1_full_size_properties_and_sampled_oracle_parity():
    H, mf, mb = _inputs()
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
        pert = _run(H[sel], mf[sel], mb[sel], False, eps=1e-12)
    for k in keys:
        scale = np.maximum(np.max(np.abs(want[k]), axis=0, keepdims=True), 1e-300)
        err = np.abs(got[k] - want[k]) / scale
        sens = np.maximum.accumulate(np.abs(pert[k] - want[k]) / scale, axis=0)
        assert not (err > np.maximum(1e-9, 10 * sens)).any(), (k, float(err.max()), float(sens.max()))
