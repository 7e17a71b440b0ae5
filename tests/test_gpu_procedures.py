"""GPU parity of the batched sweep procedures (cmtj_b200.utils.procedures)
against goldens generated from the UNMODIFIED reference procedures (tests/golden/post_*.npz) and
against the CPU checker (C oracle + oracle/post_oracle.py) on seeded inputs.

Tolerances.  PIMM quantities are linear / smooth functions of trajectories that agree to <= 1e-9
(north_star), so they are held to 1e-9 relative to their scale.  The spin-diode voltage goes through
the reference's order-3 Butterworth at 2e-6 of Nyquist, whose poles sit 6e-6 from the unit circle: a
1e-12 relative perturbation of R moves the REFERENCE's own result by ~1e-4.  The IIR itself is
therefore pinned separately and tightly (same inputs -> same output as scipy.lfilter), and the
end-to-end value is held to 10x the measured sensitivity of the reference computation."""
import os
import sys

import numpy as np
import pytest

import cmtj_b200 as cm
import helpers
import post_helpers as ph

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import post_scenarios as ps  # noqa: E402

P = cm.utils.procedures
TOL = 1e-9


def _close(got, want, tol=TOL, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = helpers.max_rel_err(got, want)
    assert err <= tol, (what, err)


@pytest.mark.parametrize("name", ["pimm_fresh", "pimm_fresh_full", "pimm_carried", "pimm_static"])
def test_pimm_matches_reference_golden(on_cuda, name):
    got, want = ps.SCENARIOS[name](cm, P, True), ph.golden(name)
    assert set(got) == set(want)
    for k in want:
        _close(got[k], want[k], what=f"{name}/{k}")


def test_pimm_batched_vs_oracle_random_sweep(on_cuda, oracle):
    """64 random field points, single-layer junction, x pulse, instance chunking forced."""
    rng = np.random.default_rng(3)
    H = rng.uniform(-400e3, 400e3, (64, 3))
    rp = ps.rparams(P, 1)
    kw = dict(Hoe_direction=cm.Axis.xaxis, Hoe_excitation=300, simulation_duration=0.3e-9, max_frequency=60e9, take_last_n=10)
    spectrum, freqs, out = P.PIMM_procedure(helpers.single_layer_mtj(cm), H, 1e-13, rp, disturbance=0, memory_budget=2e6, carry_state=False, **kw)
    assert spectrum.shape[0] == 64 and freqs.shape == (spectrum.shape[1], 1)
    with helpers.facade_on(oracle):
        for i in rng.choice(64, 8, replace=False):
            ref = ph.pimm_point(helpers.single_layer_mtj(cm), H[i], 1e-13, rp, **kw)
            _close(spectrum[i], ref["spectrum"].ravel(), what=f"spectrum[{i}]")
            _close(out["Rx"][i], ref["Rx"], what="Rx")
            _close(out["Ry"][i], ref["Ry"], what="Ry")
            _close(out["m_avg"][i], ref["m_avg"], what="m_avg")
            np.testing.assert_array_equal(freqs, ref["freqs"])


def test_sd_filter_kernel_is_scipy_lfilter(on_cuda):
    """Same inputs -> the device IIR reproduces scipy's lfilter (operation order, no FMA)."""
    from scipy.signal import butter, lfilter

    rng = np.random.default_rng(0)
    n, B, dt = 20000, 33, 1e-13
    t = np.arange(n) * dt
    R = 150 + 20 * np.cos(2 * np.pi * 5e9 * t[:, None] + rng.uniform(0, 6, B)) + 1e-3 * rng.normal(size=(n, B))
    I = np.sin(2 * np.pi * 5e9 * t)
    got = cm.utils.compute_sd(R, I, dt)
    b, a = butter(3, 10e6 / (0.5 / dt), btype="low")
    want = lfilter(b, a, -I[:, None] * R, axis=0).mean(axis=0)
    _close(got, want, tol=1e-11, what="compute_sd [n,B]")
    assert abs(cm.utils.compute_sd(R[:, 0], I, dt) - want[0]) <= 1e-11 * abs(want[0])
    Ib = np.sin(2 * np.pi * t[:, None] * rng.uniform(3e9, 9e9, B))
    want2 = lfilter(b, a, -Ib * R, axis=0).mean(axis=0)
    _close(cm.utils.compute_sd(R, Ib, dt), want2, tol=1e-11, what="per-series current")


def _vsd_bound(junction_builder, n_layers, H, f, kw, rng):
    """|vmix| of the checker and the sensitivity of the REFERENCE computation to a 1e-12 relative
    perturbation of R (the conditioning of compute_sd), for one sweep point."""
    j = junction_builder(cm)
    v = ph.vsd_point(j, H, f, ps.DT, ps.rparams(P, n_layers), Hoe_excitation=ps.VSD_KW["Hoe_excitation"],
                     simulation_duration=ps.VSD_KW["simulation_duration"], **kw)
    log = j.getLog()
    ids = j.getLayerIds()
    pert = {k: (np.asarray(x) * (1 + 1e-12 * rng.normal(size=len(x))) if k != "time" else x) for k, x in log.items()}
    params = ph._tuples(ps.rparams(P, n_layers))
    v2 = ph.post_oracle.vsd_point(pert, ids, float(f), ps.DT, kw.get("Rtype", "Rz"), params, kw.get("parallel", False))
    return v, abs(v2 - v)


@pytest.mark.parametrize("name,build,n,kw", [
    ("vsd_rz_fresh", helpers.two_layer_mtj, 2, dict(Rtype="Rz")),
    ("vsd_rx_fresh", helpers.two_layer_mtj, 2, dict(Rtype="Rx", Hoe_direction=cm.Axis.xaxis)),
    ("vsd_ry_parallel_single", helpers.single_layer_mtj, 1, dict(Rtype="Ry", parallel=True)),
])
def test_vsd_matches_reference_golden(on_cuda, oracle, name, build, n, kw):
    got = ps.SCENARIOS[name](cm, P, True)["spectrum"]
    want = ph.golden(name)["spectrum"]
    assert got.shape == want.shape == (len(ps.H_VSD), len(ps.F_VSD))
    rng = np.random.default_rng(1)
    with helpers.facade_on(oracle):
        for i, H in enumerate(ps.H_VSD):
            for k, f in enumerate(ps.F_VSD):
                v, sens = _vsd_bound(build, n, H, f, kw, rng)
                assert v == want[i, k]  # the checker IS the reference here (bit-exact on the CPU)
                bound = max(TOL * abs(v), 10 * sens)
                assert abs(got[i, k] - want[i, k]) <= bound, (i, k, got[i, k], want[i, k], bound)
                assert bound <= 5e-3 * abs(v)  # the bound is a statement about conditioning, not a free pass


def test_vsd_carried_matches_reference_golden(on_cuda):
    got = ps.vsd_carried(cm, P, True)["spectrum"]
    want = ph.golden("vsd_carried")["spectrum"]
    # same bound family as above: 5e-3 relative would be the conditioning ceiling; observed ~1e-5
    assert helpers.max_rel_err(got, want) <= 2e-3


def test_vsd_fused_equals_standalone_kernels(on_cuda):
    """The fused in-loop reduction against the same reduction done by the stand-alone kernels on the
    full-rate device log of an identical batched run (same R to the last bit)."""
    import torch

    j = helpers.two_layer_mtj(cm)
    H = np.repeat(ps.H_VSD, 3, axis=0)
    f = np.tile(ps.F_VSD, 2)
    got = P.VSD_procedure(helpers.two_layer_mtj(cm), ps.H_VSD, ps.F_VSD, disturbance=0, resistance_params=ps.rparams(P, 2),
                          Rtype="Rz", carry_state=False, **ps.VSD_KW).ravel()
    oe = cm.ScalarDriver.getSineDriver(0, ps.VSD_KW["Hoe_excitation"], float(f[0]), 0)
    j.setLayerOerstedFieldDriver("all", cm.AxialDriver(cm.NullDriver(), oe, cm.NullDriver()))
    b = cm.JunctionBatch(j, 6)
    b.set_external_field("all", H)
    b.set_override("all", "Hoe_y", "frequency", f)
    res = b.run_device(ps.VSD_KW["simulation_duration"], ps.DT, ps.DT)
    R = res["R_free_bottom"]
    t = res["time"]
    I = torch.sin(torch.from_numpy(2 * np.pi * f).cuda()[None, :] * t[:, None])
    want = cm.utils.resistance.compute_sd_device(R, I, ps.DT).cpu().numpy()
    _close(got, want, tol=1e-6, what="fused vs stand-alone")


def test_resistance_kernels_match_numpy(on_cuda):
    rng = np.random.default_rng(2)
    m = rng.normal(size=(3, 3, 50))
    m /= np.sqrt((m * m).sum(axis=1, keepdims=True))
    params = ph._tuples(ps.rparams(P, 3))
    args = ([p[0] for p in params], [p[1] for p in params], [p[2] for p in params], [p[3] for p in params], [p[4] for p in params])
    lw = dict(l=[p[6] for p in params], w=[p[5] for p in params])
    for par, fn in ((False, cm.utils.calculate_resistance_series), (True, cm.utils.calculate_resistance_parallel)):
        want = ph.post_oracle.calculate_resistance(params, m, par)
        got = fn(*args, m, **lw)
        _close(got[0], want[0], tol=1e-14, what="Rx")
        _close(got[1], want[1], tol=1e-14, what="Ry")
        g0 = fn(*args, m[:, :, 0], **lw)
        assert abs(g0[0] - want[0][0]) <= 1e-13 * abs(want[0][0])
    Sx, Sy = cm.utils.compute_resistance(*args, m, **lw)
    wx, wy = ph.post_oracle.compute_resistance(*args, m, **lw)
    _close(Sx, wx, tol=1e-14)
    _close(Sy, wy, tol=1e-14)
    want = 100 + 0.5 * (200 - 100) * (1 - np.sum(m[0] * m[1], axis=0))
    _close(cm.utils.compute_gmr(100, 200, m[0], m[1]), want, tol=1e-14)
    _close(cm.utils.calculate_magnetoresistance(100, 200, m[:2]), 100 + 50 * np.sum(m[0] * m[1], axis=0), tol=1e-13)


def test_spectrum_strip_matches_scipy(on_cuda):
    rng = np.random.default_rng(4)
    x = rng.normal(size=5001)
    yf, fr = cm.utils.compute_spectrum_strip(x, 1e-13, 80e9)
    wy, wf = ph.post_oracle.compute_spectrum_strip(x, 1e-13, 80e9)
    np.testing.assert_array_equal(fr, wf)
    _close(yf, wy, tol=1e-12)


@pytest.mark.parametrize("n,B,K_freq,pitch,mode", [
    (200000, 9, 80e9, 0, ""),        # PIMM shape, K = 1601 of 100 001 bins: Chebyshev moments of 100-sample blocks, 2000-point transforms
    (200000, 70, 80e9, 2, "moments:250"),  # blocks of 250: transforms of 800 points < K (aliased index), D = 32, more than one 32-series tile
    (200000, 5, 10e9, 0, "moments:40:nows"),  # K = 201, short blocks; the library allocates the moment slab itself
    (6561, 3, 1e11, 0, "moments:27"),      # odd block length (middle sample), radix-3 transforms
    (6561, 40, 1e11, 0, "moments:27"),     # the same through the bulk-copy kernel (one whole 32-series tile) + 8 ragged series
    (200000, 33, 80e9, 1, "moments"),      # odd pitch: rows are not 16-byte aligned, register-staged kernel only
    (200000, 9, 80e9, 0, "grouped"), # single pass, n1 = 2000 >= K = 1601, four instances per block, ragged last group
    (8000, 5, 400e9, 3, ""),         # a row view with a pitch, K = 320
    (4096, 4, 4e12, 0, ""),          # K = 1639 of a power of two
    (10000, 3, 2.0995e12, 0, ""),    # K = 2100 > n1 = 2000: aliased first-stage index
    (6000, 3, 4.9e12, 0, ""),        # K = 2941: too many bins for four instances' sums, one instance per block
    (6561, 2, 1e12, 0, ""),          # 3^8: radix-3 stages only
    (10007, 3, 1e12, 0, ""),         # prime length: library path
    (200000, 3, 80e9, 0, "single"),  # one instance per block
    (5000, 6, 80e9, 0, "cufft"),
])
def test_spectrum_device_shapes(on_cuda, monkeypatch, n, B, K_freq, pitch, mode):
    """|rfft| of time-major [n, B] series, K lowest bins: pruned single-pass kernels and the cuFFT path vs numpy"""
    import torch

    if mode:
        monkeypatch.setenv("CMTJB_SPECTRUM", mode.split(":")[0])
        if ":" in mode:
            monkeypatch.setenv("CMTJB_SPECTRUM_N2", mode.split(":")[1])
            if mode.endswith(":nows"):
                monkeypatch.setenv("CMTJB_SPECTRUM_NO_WS", "1")
            monkeypatch.setenv("CMTJB_SPECTRUM_SLAB_MB", "8")  # several slabs
    rng = np.random.default_rng(n + B)
    t = np.arange(n)[:, None]
    x = rng.normal(size=(n, B)) + 3.0 * np.sin(2 * np.pi * t * (7.25 + np.arange(B)[None, :]) / n) + 0.5
    full = torch.zeros((n, B + pitch), dtype=torch.float64, device="cuda")
    full[:, :B] = torch.from_numpy(x).cuda()
    yf, fr = P.spectrum_device(full[:, :B], 1e-13, K_freq)
    want = np.abs(np.fft.rfft(x, axis=0))[: fr.size].T
    assert yf.shape == want.shape and fr.size > 0
    scale = np.abs(want).max(axis=1, keepdims=True)
    assert np.max(np.abs(yf.cpu().numpy() - want) / scale) < 1e-12


@pytest.mark.parametrize("B,chunk_mb,wait,last_n,overlap", [
    (40, 0.2, 0.0, 20, True),      # 40 series x 5000 steps in chunks of a few blocks (one full 32-series tile + 8 ragged)
    (70, 0.5, 0.1e-9, 100, True),  # wait_time > 0: the series starts inside the run; last-n longer than a block (merged tail chunk)
    (40, 0.2, 0.0, 20, False),     # the default: the same chunks on one stream, foreground kernels only
    (3, 8192, 0.0, 20, True),      # one chunk
])
def test_pimm_streamed_spectrum_equals_resident_log(on_cuda, monkeypatch, B, chunk_mb, wait, last_n, overlap):
    """PIMM_batch reduces the projected signal chunk by chunk (_SpectrumStream: block moments, no resident log); the
    result must equal the whole-log path (CMTJB_PIMM_RESIDENT_LOG=1) to rounding, and numpy's rfft of that log."""
    rng = np.random.default_rng(B)
    H = rng.uniform(-300e3, 300e3, (B, 3))
    kw = dict(int_step=1e-13, resistance_params=ps.rparams(P, 2), simulation_duration=0.5e-9, take_last_n=last_n, max_frequency=80e9,
              Hoe_excitation=500, wait_time=wait, disturbance=0, disable_tqdm=True)
    monkeypatch.setenv("CMTJB_PIMM_CHUNK_MB", str(chunk_mb))
    calls = []
    orig = P._SpectrumStream.add
    monkeypatch.setattr(P._SpectrumStream, "add",
                        lambda self, chunk, q0, nq, background=False: (calls.append((q0, nq, background)), orig(self, chunk, q0, nq, background))[1])
    if overlap:
        monkeypatch.setenv("CMTJB_PIMM_OVERLAP", "1")  # the two-stream experiment: background reduction kernel
    s1, f1, o1 = P.PIMM_batch(helpers.two_layer_mtj(cm), H, **kw)
    assert calls and sum(nq for _, nq, _ in calls) > 0  # the streamed path ran
    if chunk_mb < 1:
        assert len(calls) > 1, calls
        # every chunk but the last is reduced by the background kernel while the next one integrates
        assert [bg for _, _, bg in calls] == [overlap] * (len(calls) - 1) + [False]
    monkeypatch.setenv("CMTJB_PIMM_RESIDENT_LOG", "1")
    n_calls = len(calls)
    s2, f2, o2 = P.PIMM_batch(helpers.two_layer_mtj(cm), H, **kw)
    assert len(calls) == n_calls  # the resident path did not stream
    np.testing.assert_array_equal(f1, f2)
    scale = np.abs(s2).max(axis=1, keepdims=True)
    assert np.max(np.abs(s1 - s2) / scale) < 1e-12
    for k in ("Rx", "Ry", "m_avg"):
        np.testing.assert_array_equal(np.asarray(o1[k]), np.asarray(o2[k]))  # the trajectories themselves are the same launches' work


def test_spectrum_stream_chunks_in_any_order(on_cuda):
    import torch

    n, B = 6000, 37
    rng = np.random.default_rng(5)
    x = rng.normal(size=(n, B)) + 2.0 * np.sin(2 * np.pi * np.arange(n)[:, None] * (3.5 + np.arange(B)[None, :]) / n)
    st = P._SpectrumStream(n, B, 1e-13, 1e11, 0)
    assert st.block_len > 0 and st.n_blocks * st.block_len == n
    st.allocate()
    xd = torch.from_numpy(x).cuda()
    L = st.block_len
    cuts = [0, 2, 6, st.n_blocks - (st.n_blocks % 2 or 2), st.n_blocks]
    spans = [(a, b - a) for a, b in zip(cuts, cuts[1:]) if b > a]
    for i, (q0, nq) in enumerate(reversed(spans)):
        st.add(xd[q0 * L : (q0 + nq) * L].contiguous(), q0, nq, background=bool(i % 2))  # both forms of the reduction
    got = st.finish().cpu().numpy()
    assert np.array_equal(st.finish_to_host(n_slabs=3), got)  # slab-wise transform with the copies overlapped: same numbers
    want = np.abs(np.fft.rfft(x, axis=0))[: st.K].T
    assert np.max(np.abs(got - want) / np.abs(want).max(axis=1, keepdims=True)) < 1e-12


def test_distribute_covers_the_product(on_cuda, oracle):
    j = helpers.two_layer_mtj(cm)
    Hx, Hz = [-100e3, 50e3, 200e3], [0.0, 30e3]
    seen = {}
    for idx, out in cm.utils.distribute(j, [("all", "Hext_x", "constantValue"), ("free", "Hext_z", "constantValue")], [Hx, Hz],
                                        2e-11, 1e-13, 1e-12, reduce_fn=lambda log: log["R_free_bottom"][-1]):
        seen[tuple(idx)] = out
    assert sorted(seen) == [(a, b) for a in range(3) for b in range(2)]
    with helpers.facade_on(oracle):
        for (a, b), v in seen.items():
            jj = helpers.two_layer_mtj(cm)
            jj.setLayerExternalFieldDriver("all", cm.AxialDriver(Hx[a], 0.0, 0.0))
            jj.setLayerExternalFieldDriver("free", cm.AxialDriver(Hx[a], 0.0, Hz[b]))
            jj.runSimulation(2e-11, 1e-13, 1e-12)
            assert abs(jj.getLog()["R_free_bottom"][-1] - v) <= TOL * abs(v)
