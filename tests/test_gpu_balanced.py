"""The SMSP-balanced persistent form of the fixed-step kernels (llgs_balanced_kernel: the run cut into
(round, instance-warp) tasks drawn from a ticket counter, state handed over through m_state) must
be BIT-IDENTICAL to the plain one-thread-per-instance form: same arithmetic, only the assignment
of work to warps changes.  CMTJB_BALANCE=1/0 forces either form; forced runs use >= 3 rounds whose
length is deliberately not a multiple of the log period."""
import os

import numpy as np
import pytest

import cmtj_b200 as cm
import helpers
from cmtj_b200.batch import PostSpec

pytestmark = pytest.mark.gpu
DT = 1e-13


def _forced(flag, fn):
    prev = os.environ.get("CMTJB_BALANCE")
    os.environ["CMTJB_BALANCE"] = flag
    try:
        return fn()
    finally:
        if prev is None:
            del os.environ["CMTJB_BALANCE"]
        else:
            os.environ["CMTJB_BALANCE"] = prev


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and np.array_equal(x, y), k


def _batch(build, B, seed=0):
    rng = np.random.default_rng(seed)
    j = build(cm)
    oe = cm.ScalarDriver.getStepDriver(0, 50, 0, 3 * DT)
    j.setLayerOerstedFieldDriver("all", cm.AxialDriver(cm.NullDriver(), cm.NullDriver(), oe))
    b = cm.JunctionBatch(j, B)
    H = rng.uniform(-300e3, 300e3, (B, 3))
    b.set_external_field("all", H)
    for lid in j.getLayerIds():
        b.set_magnetisation(lid, rng.normal(0, 0.2, (B, 3)) + [0, 0, 1.0])
    return b


@pytest.mark.parametrize("build,B,steps,we", [
    (helpers.two_layer_mtj, 1000, 1000, 7),    # ragged last warp, rounds not aligned with the log period
    (helpers.two_layer_mtj, 33, 257, 1),       # log every step
    (helpers.single_layer_mtj, 4097, 300, 100),
    (helpers.tri_layer_mtj, 250, 600, 11),
])
def test_balanced_rk4_is_bit_identical(on_cuda, build, B, steps, we):
    def run():
        b = _batch(build, B)
        out = b.run(steps * DT * (1 + 1e-9), DT, we * DT * (1 + 1e-9))
        out = dict(out)
        out["_final"] = b.m_state.copy()
        return out

    _same(_forced("1", run), _forced("0", run))


@pytest.mark.parametrize("mode", ["Heun", "EulerHeun"])
def test_balanced_stochastic_is_bit_identical(on_cuda, mode):
    """Thermal noise is Philox keyed by (seed, instance, layer, step): independent of which warp
    integrates which sub-range."""
    def run():
        j = helpers.single_layer_mtj(cm)
        j.setLayerTemperatureDriver("all", cm.constantDriver(300.0))
        b = cm.JunctionBatch(j, 777)
        out = dict(b.run(500 * DT * (1 + 1e-9), DT, 13 * DT * (1 + 1e-9), solverMode=getattr(cm.SolverMode, mode), seed=42))
        out["_final"] = b.m_state.copy()
        return out

    _same(_forced("1", run), _forced("0", run))


def test_balanced_no_log_final_state_and_carry(on_cuda):
    """LOG_NONE launches and a continued run (start_iteration) through the balanced form."""
    import torch

    def run():
        b = _batch(helpers.two_layer_mtj, 500, seed=3)
        b.run_device(400 * DT * (1 + 1e-9), DT, DT, log=False)
        res = b.run_device(600 * DT * (1 + 1e-9), DT, 9 * DT * (1 + 1e-9), start_iteration=400)
        torch.cuda.synchronize()
        return {"log": res["_log"].cpu().numpy(), "final": b._m_dev.cpu().numpy()}

    _same(_forced("1", run), _forced("0", run))


def test_balanced_fused_post_is_bit_identical(on_cuda):
    """Projection log and the spin-diode IIR carry (sd_state) across task boundaries."""
    import torch

    from scipy.signal import butter

    def run():
        B = 300
        b = _batch(helpers.two_layer_mtj, B, seed=5)
        n = 700
        dev = torch.device("cuda", 0)
        proj = np.zeros((1, 6))
        proj[0, 2], proj[0, 5] = 0.6, 0.4
        proj_log = torch.zeros((1, n, B), dtype=torch.float64, device=dev)
        bb, aa = butter(3, 10e6 / (0.5 / DT), btype="low")
        state = torch.zeros((4, B), dtype=torch.float64, device=dev)
        omega = torch.full((B,), 2 * np.pi * 5e9, dtype=torch.float64, device=dev)
        from cmtj_b200 import _abi

        post = PostSpec(proj=proj, proj_log=proj_log, proj_sample0=0, sd_mode=_abi.SD_RZ, res=np.zeros((2, _abi.RES_NFIELDS)),
                        sd_b=bb, sd_a=aa / aa[0], sd_omega=omega, sd_state=state)
        b.run_device(n * DT * (1 + 1e-9), DT, DT, log=False, post=post)
        torch.cuda.synchronize()
        return {"proj": proj_log.cpu().numpy(), "sd": state.cpu().numpy(), "final": b._m_dev.cpu().numpy()}

    _same(_forced("1", run), _forced("0", run))


@pytest.mark.parametrize("ring", [None, "2", "3"])
def test_host_entry_streams_large_logs_in_time_chunks(on_cuda, monkeypatch, ring):
    """cmtjb_run_host cuts logs above 64 MB into time chunks that integrate into a ring of device buffers
    while finished chunks are copied out; the assembled host log must equal the device-resident log of the
    same run bit for bit (sample rows that do not divide evenly into chunks; ring forced down to 2 and 3
    buffers so that buffers are reused while their previous content is still being copied)."""
    import torch

    if ring:
        monkeypatch.setenv("CMTJB_HOST_RING", ring)

    B, steps, we = 5000, 3030, 10  # 7 x 303 x 5000 doubles = 85 MB -> 32 chunks of 9.5 samples
    b = _batch(helpers.two_layer_mtj, B, seed=9)
    host = dict(b.run(steps * DT * (1 + 1e-9), DT, we * DT * (1 + 1e-9)))
    final_host = b.m_state.copy()
    b2 = _batch(helpers.two_layer_mtj, B, seed=9)
    dev = b2.run_device(steps * DT * (1 + 1e-9), DT, we * DT * (1 + 1e-9))
    torch.cuda.synchronize()
    for k, v in host.items():
        if k != "time":
            assert np.array_equal(dev[k].cpu().numpy(), v), k
    assert np.array_equal(b2._m_dev.cpu().numpy(), final_host)
