"""Stochastic parity (north_star: "stochastic runs must agree in distribution (switching
probability, PSD) within stated confidence bounds, since the RNG streams differ").

GPU side: Philox4x32-10 + Box-Muller keyed by (seed, instance, layer, step).  Reference side:
std::normal_distribution over mt19937(random_device), through the UNMODIFIED reference
extension (oracle/_ref travels to the GPU box).  Every bound below is a >= 4-sigma band of the
estimator it applies to, so a correct implementation fails a given assertion with p < 1e-4."""
import ctypes as C
import math

import numpy as np
import pytest

import cmtj_b200
import helpers
from cmtj_b200 import _abi

pytestmark = pytest.mark.gpu
m = cmtj_b200


# ---- the generator itself ----------------------------------------------------------------------------
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)."""
    lib = _abi.load_library()
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        assert lib.cmtjb_philox_host(c, k, o, 0) == 0
        assert list(o) == want


def _normals(seed, inst0, n_inst, layer, step0, n_steps):
    lib = _abi.load_library()
    out = np.zeros((n_steps, n_inst, 3))
    f = lib.cmtjb_normals_host
    f.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
    assert f(seed, inst0, n_inst, layer, step0, n_steps, out.ctypes.data, 0) == 0
    return out


def test_normal_stream_moments_and_independence():
    z = _normals(12345, 0, 4096, 0, 0, 256)  # 3.1e6 normals
    n = z.size
    flat = z.ravel()
    assert abs(flat.mean()) < 4 / math.sqrt(n)
    assert abs(flat.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((flat**3).mean()) < 4 * math.sqrt(15 / n)
    assert abs((flat**4).mean() - 3) < 4 * math.sqrt(96 / n)
    # tails: P(|z| > 3) = 2.6998e-3
    p3 = 2.699796e-3
    assert abs((np.abs(flat) > 3).mean() - p3) < 4 * math.sqrt(p3 / n)
    # Kolmogorov-Smirnov against N(0,1) on a sub-sample (critical value at alpha=1e-4: 2.23/sqrt(n))
    from scipy import special

    sub = np.sort(flat[:: 16])
    cdf = 0.5 * (1 + special.erf(sub / math.sqrt(2)))
    d = np.max(np.abs(cdf - (np.arange(1, len(sub) + 1) - 0.5) / len(sub)))
    assert d < 2.23 / math.sqrt(len(sub)) + 0.5 / len(sub)
    # independence: across components, neighbouring instances, consecutive steps, layers
    s4 = 4 / math.sqrt(z.shape[0] * z.shape[1])
    assert abs((z[..., 0] * z[..., 1]).mean()) < s4 and abs((z[..., 0] * z[..., 2]).mean()) < s4
    assert abs((z[:, :-1, 0] * z[:, 1:, 0]).mean()) < s4
    assert abs((z[:-1, :, 1] * z[1:, :, 1]).mean()) < s4
    z1 = _normals(12345, 0, 4096, 1, 0, 256)
    assert abs((z[..., 0] * z1[..., 0]).mean()) < s4
    # sharding: instance_offset continues the same streams; another seed gives another stream
    assert np.array_equal(_normals(12345, 1000, 8, 0, 5, 3), z[5:8, 1000:1008])
    assert not np.array_equal(_normals(12346, 0, 8, 0, 0, 3), z[:3, :8])


# ---- sLLGS dynamics vs the reference ---------------------------------------------------------------
def _pbit(mod, Hx=0.0, T=300.0, mag=(1, 0, 0.02)):
    """p-bit / thermal-switching junction of curated-examples/p-bit.py:188-227 (BASELINE config 3)
    with a 7 nm radius free layer so that 300 K matters within 0.5 ns."""
    demag = [(0.00024164288391924, 0, 0), (0, 0.000160046006320031, 0), (0, 0, 0.999598310229469)]
    l = mod.Layer.createSTTLayer("free", mod.CVector(*mag), mod.CVector(1, 0, 0), 1.6, 2e-9, math.pi * (0.7e-9) ** 2 * 1e2,
                                 [mod.CVector(*r) for r in demag], 0.1, 1.0, 0.9, 1.0)
    l.setReferenceLayer(mod.CVector(1, 0, 0))
    j = mod.Junction([l], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(350e3))
    j.setLayerCurrentDriver("free", mod.ScalarDriver.getConstantDriver(-1e10))
    j.setLayerExternalFieldDriver("free", mod.AxialDriver(Hx, 0.0, 0.0))
    j.setLayerTemperatureDriver("free", mod.ScalarDriver.getConstantDriver(T))
    return j


FIELDS = [-260e3, -280e3, -300e3, -320e3, -340e3]


@pytest.mark.parametrize("mode", ["Heun", "EulerHeun"])
def test_switching_probability_curve(ref, mode):
    """P(m_x < 0 after 0.5 ns) vs opposing field at 300 K: 20000 GPU trials per field in ONE batch
    against 500 reference trials per field; two-proportion z-test at 4 sigma per point."""
    n_gpu, n_ref, steps = 20000, 500, 5000
    j = _pbit(m)
    b = m.JunctionBatch(j, n_gpu * len(FIELDS))
    H = np.zeros((b.B, 3))
    H[:, 0] = np.repeat(FIELDS, n_gpu)
    b.set_external_field("free", H)
    b.run(steps * 1e-13, 1e-13, steps * 1e-13, solverMode=getattr(m.SolverMode, mode), log=False, seed=2024)
    assert "F_THERMAL" in b.last_plan.variant.decode()
    mx = b.get_magnetisation("free")[:, 0].reshape(len(FIELDS), n_gpu)
    p_gpu = (mx < 0).mean(axis=1)
    p_ref = []
    for Hx in FIELDS:
        k = 0
        for _ in range(n_ref):
            jr = _pbit(ref, Hx)
            jr.runSimulation(steps * 1e-13, 1e-13, steps * 1e-13, solverMode=getattr(ref.SolverMode, mode))
            k += jr.getLayerMagnetisation("free").x < 0
        p_ref.append(k / n_ref)
    p_ref = np.array(p_ref)
    pooled = (p_gpu * n_gpu + p_ref * n_ref) / (n_gpu + n_ref)
    sigma = np.sqrt(np.maximum(pooled * (1 - pooled), 1e-4) * (1 / n_gpu + 1 / n_ref))
    assert np.all(np.abs(p_gpu - p_ref) <= 4 * sigma), (p_gpu, p_ref, sigma)
    # the curve is a curve: monotone in the field and spanning the transition
    assert p_gpu[0] < 0.35 and p_gpu[-1] > 0.65 and np.all(np.diff(p_gpu) > 0), p_gpu


def test_short_time_diffusion_variance(ref):
    """50 Euler-Heun steps from m = z at 300 K: Var(m_x), Var(m_y) against 4000 reference trials
    (relative standard error of a variance estimate: sqrt(2/n))."""
    n_gpu, n_ref, steps = 100000, 4000, 50
    j = _pbit(m, mag=(0, 0, 1))
    b = m.JunctionBatch(j, n_gpu)
    b.run(steps * 1e-13, 1e-13, steps * 1e-13, log=False, seed=7)
    g = b.get_magnetisation("free")
    r = np.zeros((n_ref, 3))
    for k in range(n_ref):
        jr = _pbit(ref, mag=(0, 0, 1))
        jr.runSimulation(steps * 1e-13, 1e-13, steps * 1e-13)
        r[k] = jr.getLayerMagnetisation("free").tolist()
    for c in (0, 1):
        ratio = g[:, c].var() / r[:, c].var()
        assert abs(ratio - 1) < 4 * math.sqrt(2 / n_ref + 2 / n_gpu), (c, ratio)
        assert abs(g[:, c].mean() - r[:, c].mean()) < 4 * math.sqrt(r[:, c].var() / n_ref), c
    assert np.allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-12)


def test_thermal_fmr_psd(ref):
    """Thermally excited precession: Welch-averaged PSD of m_x(t) around the resonance line, GPU
    (2048 trajectories) vs reference (48 trajectories); the ratio of the band-integrated power must
    be within 4 sigma of 1 and the line must sit at the same frequency (band centroid within 4 standard errors)."""
    steps, n_gpu, n_ref = 8192, 2048, 48

    def build(mod):
        j = _pbit(mod, Hx=0.0, mag=(1, 0, 0.0))
        j.setLayerCurrentDriver("free", mod.ScalarDriver.getConstantDriver(0.0))
        j.setLayerExternalFieldDriver("free", mod.AxialDriver(200e3, 0.0, 0.0))
        return j

    b = m.JunctionBatch(build(m), n_gpu)
    log = b.run(steps * 1e-13, 1e-13, 1e-13, solverMode=m.SolverMode.Heun, seed=99)
    yg = log["free_my"]  # [steps, n_gpu]
    yr = np.zeros((steps, n_ref))
    for k in range(n_ref):
        jr = build(ref)
        jr.runSimulation(steps * 1e-13, 1e-13, 1e-13, solverMode=ref.SolverMode.Heun)
        yr[:, k] = jr.getLog()["free_my"]
    win = np.hanning(steps)[:, None]

    def psd(y):
        return (np.abs(np.fft.rfft((y - y.mean(axis=0)) * win, axis=0)) ** 2)

    Pg, Pr = psd(yg), psd(yr)
    pg, pr = Pg.mean(axis=1), Pr.mean(axis=1)
    kg = int(np.argmax(pg[1:])) + 1  # 2048-trajectory average: the line position is well resolved
    lo, hi = max(kg - 6, 1), kg + 7
    # line position: power-weighted centroid of the band per trajectory, means compared at 4 standard errors
    # (the argmax of the 48-trajectory reference average alone wanders by a couple of bins: each bin of it
    # is a chi-square estimate with a 14 % relative spread, and the reference cannot be seeded)
    k = np.arange(lo, hi)[:, None]
    cg, cr = (k * Pg[lo:hi]).sum(axis=0) / Pg[lo:hi].sum(axis=0), (k * Pr[lo:hi]).sum(axis=0) / Pr[lo:hi].sum(axis=0)
    se_c = math.sqrt(cg.var() / n_gpu + cr.var() / n_ref)
    assert abs(cg.mean() - cr.mean()) < max(4 * se_c, 0.25), (cg.mean(), cr.mean(), se_c)
    assert abs(int(np.argmax(pr[lo:hi])) + lo - kg) <= 3, (kg, int(np.argmax(pr[lo:hi])) + lo)
    # band power per trajectory -> mean and standard error on both sides
    bg, br = Pg[lo:hi].sum(axis=0), Pr[lo:hi].sum(axis=0)
    se = math.sqrt(bg.var() / n_gpu / bg.mean() ** 2 + br.var() / n_ref / br.mean() ** 2)
    assert abs(bg.mean() / br.mean() - 1) < 4 * se, (bg.mean() / br.mean(), se)
    # total fluctuation power (equipartition level) as well
    tg, tr = yg.var(axis=0), yr.var(axis=0)
    se_t = math.sqrt(tg.var() / n_gpu / tg.mean() ** 2 + tr.var() / n_ref / tr.mean() ** 2)
    assert abs(tg.mean() / tr.mean() - 1) < 4 * se_t, (tg.mean() / tr.mean(), se_t)


def test_alpha_noise_buffer_reaches_the_dynamics(ref):
    """setAlphaNoise -> VectorAlphaNoise buffer added to dW under Euler-Heun (junction.hpp:1130-1136):
    with no temperature the trajectory spread must match the reference's in distribution."""
    steps, n = 400, 300

    def run(mod):
        out = np.zeros(n)
        for k in range(n):
            j = helpers.single_layer_mtj(mod, mag=(0, 0, 1))
            j.getLayer("free").setAlphaNoise(1.0, 1e-2, 1e-2, mod.Axis.all)
            j.runSimulation(steps * 1e-13, 1e-13, steps * 1e-13)
            out[k] = j.getLayerMagnetisation("free").x
        return out

    g, r = run(m), run(ref)
    assert g.std() > 0
    ratio = g.var() / r.var()
    # the increments are 1/f-correlated within a run but runs are independent: variance-ratio bound
    assert abs(ratio - 1) < 4 * math.sqrt(4 / n), ratio
