"""Stochastic parity (north_star: "stochastic runs must agree in distribution (switching
probability, PSD) within stated confidence bounds, since the RNG streams differ").

GPU side: Philox4x32-10 + Box-Muller keyed by (seed, instance, layer, step).  Reference side:
std::normal_distribution over mt19937(random_device), through the UNMODIFIED reference
extension (oracle/_ref travels to the GPU box), >= 1024 trajectories per assertion from a process pool
(tests/ref_pool.py).  Every comparison with the reference is a two-sample z-test on the means of a
per-trajectory statistic with the pooled standard error, bounded at stat_helpers.Z_BOUND = 5.3 sigma:
false-failure probability < 1e-6 per assertion (derivation and bootstrap calibration in
tests/stat_helpers.py).  The achieved z values are recorded in the parity report (conftest.py)."""
import ctypes as C
import math

import numpy as np
import pytest

import cmtj_b200
import helpers
import ref_pool
import stat_helpers as sh
from cmtj_b200 import _abi
from parity_report import record

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


def test_float64_normal_stream():
    """thermal_normals = 1: 53-bit uniforms, double-precision Box-Muller (cmtjb_normals_host_ex).  Same distributional
    checks as the float32 stream; the values are full-precision doubles (not representable in float32), a different stream
    from the default one, a pure function of (seed, instance, layer, step), and the radius tail is not cut at 6.7 sigma."""
    lib = _abi.load_library()
    f = lib.cmtjb_normals_host_ex
    f.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int]

    def normals64(seed, inst0, n_inst, layer, step0, n_steps, prec=1):
        out = np.zeros((n_steps, n_inst, 3))
        assert f(seed, inst0, n_inst, layer, step0, n_steps, prec, out.ctypes.data, 0) == 0
        return out

    z = normals64(777, 0, 4096, 0, 0, 256)
    flat = z.ravel()
    n = flat.size
    assert abs(flat.mean()) < 4 / math.sqrt(n)
    assert abs(flat.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((flat**3).mean()) < 4 * math.sqrt(15 / n)
    assert abs((flat**4).mean() - 3) < 4 * math.sqrt(96 / n)
    p3 = 2.699796e-3
    assert abs((np.abs(flat) > 3).mean() - p3) < 4 * math.sqrt(p3 / n)
    from scipy import special

    sub = np.sort(flat[::16])
    cdf = 0.5 * (1 + special.erf(sub / math.sqrt(2)))
    d = np.max(np.abs(cdf - (np.arange(1, len(sub) + 1) - 0.5) / len(sub)))
    assert d < 2.23 / math.sqrt(len(sub)) + 0.5 / len(sub)
    s4 = 4 / math.sqrt(z.shape[0] * z.shape[1])
    assert abs((z[..., 0] * z[..., 1]).mean()) < s4 and abs((z[..., 0] * z[..., 2]).mean()) < s4
    assert abs((z[:-1, :, 1] * z[1:, :, 1]).mean()) < s4
    # full double precision: (almost) none of the values survives a round trip through float32
    assert (flat.astype(np.float32).astype(np.float64) == flat).mean() < 1e-3
    assert (_normals(777, 0, 64, 0, 0, 4).astype(np.float32).astype(np.float64) == _normals(777, 0, 64, 0, 0, 4)).all()  # the default stream is float32
    # a different stream from the float32 one, reproducible, continued by instance_offset
    assert not np.array_equal(z[:4, :64], _normals(777, 0, 64, 0, 0, 4))
    assert np.array_equal(normals64(777, 1000, 8, 0, 5, 3), z[5:8, 1000:1008])
    assert np.array_equal(normals64(777, 0, 8, 0, 0, 3, prec=0), _normals(777, 0, 8, 0, 0, 3))


# ---- sLLGS dynamics vs the reference ---------------------------------------------------------------
@pytest.fixture(scope="module", autouse=True)
def _pool(ref):
    """reference workers for the whole module (spawned once, before the first GPU call of the module)"""
    ref_pool.pool()
    yield
    ref_pool.close()


FIELDS = [-260e3, -280e3, -300e3, -320e3, -340e3]


@pytest.fixture
def normals_precision(request):
    prev = m.thermal_normals()
    m.thermal_normals(request.param)
    yield request.param
    m.thermal_normals(prev)


@pytest.mark.parametrize("mode,normals_precision", [("Heun", "float32"), ("EulerHeun", "float32"), ("Heun", "float64")],
                         indirect=["normals_precision"])
def test_switching_probability_curve(ref, mode, normals_precision):
    """P(m_x < 0 after 0.5 ns) vs opposing field at 300 K: 20000 GPU trials per field in ONE batch against
    4096 reference trials per field; two-proportion z-test per point.  Both precisions of the Langevin normals."""
    n_gpu, n_ref, steps = 20000, 4096, 5000
    j = ref_pool.pbit(m)
    b = m.JunctionBatch(j, n_gpu * len(FIELDS))
    H = np.zeros((b.B, 3))
    H[:, 0] = np.repeat(FIELDS, n_gpu)
    b.set_external_field("free", H)
    b.run(steps * 1e-13, 1e-13, steps * 1e-13, solverMode=getattr(m.SolverMode, mode), log=False, seed=2024)
    assert "F_THERMAL" in b.last_plan.variant.decode()
    mx = b.get_magnetisation("free")[:, 0].reshape(len(FIELDS), n_gpu)
    p_gpu = (mx < 0).mean(axis=1)
    zs = []
    for k, Hx in enumerate(FIELDS):
        r = ref_pool.run(ref_pool.switching_trials, n_ref, Hx, steps, mode)
        z, d, se = sh.z_proportions(int((mx[k] < 0).sum()), n_gpu, int((r < 0).sum()), n_ref)
        zs.append(z)
        record(f"switching probability {mode}{'' if normals_precision == 'float32' else ' (float64 normals)'} H={Hx:g}", z=z, bound=sh.Z_BOUND, estimate=d, se=se, n_gpu=n_gpu, n_ref=n_ref)
    assert np.all(np.abs(zs) <= sh.Z_BOUND), (zs, p_gpu)
    # the curve is a curve: monotone in the field and spanning the transition
    assert p_gpu[0] < 0.35 and p_gpu[-1] > 0.65 and np.all(np.diff(p_gpu) > 0), p_gpu


def test_short_time_diffusion_variance(ref):
    """50 Euler-Heun steps from m = z at 300 K: mean and variance of m_x, m_y against 16384 reference trials
    (per-trajectory statistic for the variance: the squared deviation from the pooled mean)."""
    n_gpu, n_ref, steps = 100000, 16384, 50
    j = ref_pool.pbit(m, mag=(0, 0, 1))
    b = m.JunctionBatch(j, n_gpu)
    b.run(steps * 1e-13, 1e-13, steps * 1e-13, log=False, seed=7)
    g = b.get_magnetisation("free")
    r = ref_pool.run(ref_pool.diffusion_trials, n_ref, steps)
    for c in (0, 1):
        z, d, se = sh.z_two_sample(g[:, c], r[:, c])
        record(f"diffusion mean m_{'xy'[c]}", z=z, bound=sh.Z_BOUND, estimate=d, se=se, n_gpu=n_gpu, n_ref=n_ref)
        assert abs(z) <= sh.Z_BOUND, (c, z)
        z, d, se = sh.z_ratio((g[:, c] - g[:, c].mean()) ** 2, (r[:, c] - r[:, c].mean()) ** 2)
        record(f"diffusion variance ratio m_{'xy'[c]}", z=z, bound=sh.Z_BOUND, estimate=d, se=se, n_gpu=n_gpu, n_ref=n_ref)
        assert abs(z) <= sh.Z_BOUND, (c, z, d)
    assert np.allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-12)


def test_thermal_fmr_psd(ref):
    """Thermally excited precession: Hann-windowed periodogram of m_y(t) around the resonance line, 4096 GPU
    trajectories against 1024 reference trajectories.  Per-trajectory statistics: the power-weighted centroid of
    the 13-bin band around the line (line position), the band-integrated power (line strength) and Var(m_y)
    (equipartition level).  One bin = 1.22 GHz: the centroid bound of 5.3 standard errors is 0.18 bins = 0.6 %
    of the 37.7 GHz line."""
    steps, n_gpu, n_ref, n_bins = 8192, 4096, 1024, 128
    b = m.JunctionBatch(ref_pool.fmr(m), n_gpu)
    log = b.run(steps * 1e-13, 1e-13, 1e-13, solverMode=m.SolverMode.Heun, seed=99)
    yg = log["free_my"]  # [steps, n_gpu]
    win = np.hanning(steps)[:, None]
    Pg = (np.abs(np.fft.rfft((yg - yg.mean(axis=0)) * win, axis=0)[:n_bins]) ** 2).T  # [n_gpu, n_bins]
    Vg = yg.var(axis=0)
    S = ref_pool.run(ref_pool.fmr_spectra, n_ref, steps, n_bins)
    Pr, Vr = S[:, :n_bins], S[:, n_bins]
    kg = int(np.argmax(Pg.mean(axis=0)[1:])) + 1  # 4096-trajectory average: the line position is well resolved
    lo, hi = max(kg - 6, 1), kg + 7
    assert hi <= n_bins
    k = np.arange(lo, hi)
    bg, br = Pg[:, lo:hi].sum(axis=1), Pr[:, lo:hi].sum(axis=1)
    cg, cr = (Pg[:, lo:hi] * k).sum(axis=1) / bg, (Pr[:, lo:hi] * k).sum(axis=1) / br
    z_c, d_c, se_c = sh.z_two_sample(cg, cr)
    z_b, d_b, se_b = sh.z_ratio(bg, br)
    z_v, d_v, se_v = sh.z_ratio(Vg, Vr)
    print(f"FMR line: peak bin {kg}; centroid GPU {cg.mean():.4f} vs reference {cr.mean():.4f} bins "
          f"(difference {d_c:+.4f} +- {se_c:.4f}, z = {z_c:+.2f}); band power ratio - 1 = {d_b:+.4f} +- {se_b:.4f}; "
          f"variance ratio - 1 = {d_v:+.4f} +- {se_v:.4f}")
    record("thermal FMR line centroid [bins]", z=z_c, bound=sh.Z_BOUND, estimate=d_c, se=se_c, n_gpu=n_gpu, n_ref=n_ref)
    record("thermal FMR band power ratio", z=z_b, bound=sh.Z_BOUND, estimate=d_b, se=se_b, n_gpu=n_gpu, n_ref=n_ref)
    record("thermal FMR Var(m_y) ratio", z=z_v, bound=sh.Z_BOUND, estimate=d_v, se=se_v, n_gpu=n_gpu, n_ref=n_ref)
    assert abs(z_c) <= sh.Z_BOUND, (cg.mean(), cr.mean(), se_c)
    assert abs(z_b) <= sh.Z_BOUND, (d_b, se_b)
    assert abs(z_v) <= sh.Z_BOUND, (d_v, se_v)


# ---- 1/f^alpha noise (setAlphaNoise; junction.hpp:406-425, :1130-1136; noise.hpp:190-240, :291-301) ------------
@pytest.mark.parametrize("alpha", [0.0, 1.0, 2.0])
def test_alpha_noise_generator_matches_reference(ref, alpha):
    """The device generator (Philox + cuFFT restatement of BufferedAlphaNoise::fillBuffer) against the reference's
    own buffers: per-series variance and per-series spectral slope, 2048 GPU series vs 1024 reference buffers of
    4096 samples.  Shards reproduce the batch; different seeds differ."""
    n, std, n_gpu, n_ref = 4096, 0.7, 2048, 1024
    g = m.noise.alpha_noise(n, alpha, std, seed=5, n_series=n_gpu).astype(np.float64)
    r = ref_pool.run(ref_pool.alpha_buffers, n_ref, n, alpha, std)
    assert g.shape == (n_gpu, n) and np.all(np.isfinite(g))
    z_v, d_v, se_v = sh.z_ratio(g.var(axis=1), r.var(axis=1))
    z_s, d_s, se_s = sh.z_two_sample(sh.psd_slopes(g), sh.psd_slopes(r))
    z_m, d_m, se_m = sh.z_two_sample(g.mean(axis=1), r.mean(axis=1))
    record(f"1/f generator alpha={alpha:g}: variance ratio", z=z_v, bound=sh.Z_BOUND, estimate=d_v, se=se_v, n_gpu=n_gpu, n_ref=n_ref)
    record(f"1/f generator alpha={alpha:g}: PSD slope", z=z_s, bound=sh.Z_BOUND, estimate=d_s, se=se_s, n_gpu=n_gpu, n_ref=n_ref)
    record(f"1/f generator alpha={alpha:g}: series mean", z=z_m, bound=sh.Z_BOUND, estimate=d_m, se=se_m, n_gpu=n_gpu, n_ref=n_ref)
    assert abs(z_v) <= sh.Z_BOUND and abs(z_s) <= sh.Z_BOUND and abs(z_m) <= sh.Z_BOUND, (z_v, z_s, z_m)
    assert abs(sh.psd_slopes(g).mean() + alpha) < 0.1  # and it is 1/f^alpha noise
    # series are keyed by their global index: a shard continues the batch; another seed gives other series
    assert np.array_equal(m.noise.alpha_noise(n, alpha, std, seed=5, n_series=4, series0=100), g[100:104].astype(np.float32))
    assert not np.array_equal(m.noise.alpha_noise(n, alpha, std, seed=6, n_series=4), g[:4].astype(np.float32))
    # independence across series
    assert abs(np.mean(g[:-1, n // 2] * g[1:, n // 2])) < sh.Z_BOUND * g[:, n // 2].var() / math.sqrt(n_gpu)


@pytest.mark.parametrize("axis_name,alpha", [("all", 1.0), ("yaxis", 1.5)])
def test_alpha_noise_batch_dynamics_match_reference(ref, axis_name, alpha, monkeypatch):
    """setAlphaNoise in a sweep: 4096 instances in one batch (buffers generated on the device, here in forced
    sub-batches) against 1024 reference runs, no temperature.  Per-trajectory statistics of m_z(t) sampled every 8
    steps: its final-time and time-averaged squares (spread), and the spectral slope of the trajectory."""
    steps, every, n_gpu, n_ref = 4096, 8, 4096, 1024
    std, scale = 0.5, 2e-2
    monkeypatch.setenv("CMTJB_ONEF_BUDGET_MB", "64")  # 3 series x 4096 x 4 B per instance: sub-batches of 1344 instances
    j = ref_pool.onef_junction(m, alpha, std, scale, axis_name)
    b = m.JunctionBatch(j, n_gpu)
    log = b.run(steps * 1e-13, 1e-13, every * 1e-13 * (1 + 1e-9), solverMode=m.SolverMode.EulerHeun, seed=31)
    assert "F_ONEF" in b.last_plan.variant.decode()
    g = log["free_mz"].T  # [n_gpu, samples]
    r = ref_pool.run(ref_pool.onef_trajectories_z, n_ref, steps, alpha, std, scale, axis_name, every)
    assert g.shape[1] == r.shape[1] == steps // every and g.std() > 0
    for what, fg, fr in (("final m_z^2", g[:, -1] ** 2, r[:, -1] ** 2), ("mean m_z^2", (g**2).mean(axis=1), (r**2).mean(axis=1))):
        z, d, se = sh.z_ratio(fg, fr)
        record(f"1/f dynamics axis={axis_name} alpha={alpha:g}: {what} ratio", z=z, bound=sh.Z_BOUND, estimate=d, se=se, n_gpu=n_gpu, n_ref=n_ref)
        assert abs(z) <= sh.Z_BOUND, (what, z, d, se)
    z, d, se = sh.z_two_sample(sh.psd_slopes(g, 32, 2), sh.psd_slopes(r, 32, 2))
    record(f"1/f dynamics axis={axis_name} alpha={alpha:g}: PSD slope of m_z(t)", z=z, bound=sh.Z_BOUND, estimate=d, se=se, n_gpu=n_gpu, n_ref=n_ref)
    assert abs(z) <= sh.Z_BOUND, (z, d, se)
    # the split into sub-batches is invisible: one instance range rerun alone reproduces its rows bit for bit
    monkeypatch.delenv("CMTJB_ONEF_BUDGET_MB")
    b2 = m.JunctionBatch(j, 64, instance_offset=2000)
    log2 = b2.run(steps * 1e-13, 1e-13, every * 1e-13 * (1 + 1e-9), solverMode=m.SolverMode.EulerHeun, seed=31)
    assert np.array_equal(log2["free_mz"], log["free_mz"][:, 2000:2064])


def test_alpha_noise_single_junction_and_stack(ref):
    """Junction.runSimulation with setAlphaNoise (the reference's own entry), and the same layer inside a Stack."""
    j = ref_pool.onef_junction(m, 1.0, 0.5, 2e-2, "all")
    j.runSimulation(2e-10, 1e-13, 1e-12)
    mz = np.asarray(j.getLog()["free_mz"])
    assert mz.std() > 0 and np.all(np.isfinite(mz))
    jj = [ref_pool.onef_junction(m, 1.0, 0.5, 2e-2, "all") for _ in range(2)]
    st = m.stack.SeriesStack(jj)
    st.setCoupledCurrentDriver(m.ScalarDriver.getConstantDriver(0.0))
    st.setCouplingStrength(0.0)
    st.runSimulation(2e-10, 1e-13, 1e-12)
    a, c = np.asarray(st.getLog(0)["free_mz"]), np.asarray(st.getLog(1)["free_mz"])
    assert a.std() > 0 and c.std() > 0 and not np.array_equal(a, c)  # every junction has its own series
