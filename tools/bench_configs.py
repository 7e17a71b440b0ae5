#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (S2..S5) on one GPU, device-
resident, CUDA events.  Not the headline (bench.py is); these are the per-config numbers quoted in
DESIGN.md / profiles/.  One JSON line per configuration.

usage: bench_configs.py [s2] [s3] [s4] [s5g] [s5s] [scale]   (default: all, at the sizes below)
"""
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmtj_b200 as m  # noqa: E402
import helpers  # noqa: E402
from cmtj_b200 import _abi, _coupled  # noqa: E402

lib = _abi.load_library()
fl, ms = C.c_double(), C.c_double()
lib.cmtjb_fp64_peak(0, C.c_int64(100000), C.byref(fl), C.byref(ms))
PEAK = fl.value
HBM = 6650.0
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps=3):
    # warm up for at least 0.5 s of wall time: the short configurations (a few ms per call) would
    # otherwise be timed while the SM clock is still ramping up from idle
    t0 = time.perf_counter()
    while True:
        fn()
        torch.cuda.synchronize()
        if time.perf_counter() - t0 > 0.5:
            break
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def emit(name, layer_steps, secs, flop_per_layer_step, log_bytes, extra):
    line = {
        "config": name,
        "layer_timesteps_per_s": layer_steps / secs,
        "seconds": secs,
        "fp64_frac_of_measured_peak": layer_steps * flop_per_layer_step / secs / PEAK,
        "algorithmic_flop_per_layer_step": flop_per_layer_step,
        "log_GBps": log_bytes / secs / 1e9,
        "hbm_frac": log_bytes / secs / 1e9 / HBM,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def s2(scale):
    """vsd-basic-field-scan (curated-examples/vsd-basic-field-scan.py:57-60,102-128): 1 layer, sine
    Oersted drive with per-instance frequency, logged EVERY step (HBM-write bound)."""
    B, steps = int(262144 * scale), 5000  # 52 GB of log when every step is written; ~15 ms per call
    demag = [m.CVector(0, 0, 0), m.CVector(0, 0, 0), m.CVector(0, 0, 1)]
    l = m.Layer("free", m.CVector(0, 0, 1), m.CVector(0, 0, 1), 0.525, 1.45e-9, 0.0, demag, 0.03)
    l.setReferenceLayer(m.CVector(0, 1, 0))
    j = m.Junction([l], 100, 200)
    j.setLayerAnisotropyDriver("free", m.constantDriver(1.54e5))
    j.setLayerOerstedFieldDriver("free", m.AxialDriver(m.NullDriver(), m.NullDriver(), m.sineDriver(0, -5000, 14e9, 0)))
    b = m.JunctionBatch(j, B)
    th = math.radians(91)
    Hm = np.linspace(350e3, 630e3, B)
    b.set_external_field("free", np.stack([Hm * math.sin(th), np.zeros(B), Hm * math.cos(th)], axis=1))
    T = steps * 1e-13 * (1 + 1e-9)
    for tag, we in (("log every step", 1), ("log every 100", 100)):
        out = b.run_device(T, 1e-13, we * 1e-13 * (1 + 1e-9))["_log"]
        secs = timed(lambda: b.run_device(T, 1e-13, we * 1e-13 * (1 + 1e-9), out=out), reps=7)
        emit(f"S2 vsd 1-layer sine Oersted RK4, {tag}", B * steps, secs, 225.0, out.numel() * 8,
             {"instances": B, "steps": steps, "variant": _abi.last_variant()})
        del out


def s3(scale):
    """p-bit thermal switching (curated-examples/p-bit.py:188-227): 1-layer STT, 300 K, Heun and
    Euler-Heun, final state only."""
    B, steps = int(1_000_000 * scale), 5000
    j = helpers.stt_layer_mtj(m, temperature=300.0)
    b = m.JunctionBatch(j, B)
    H = np.zeros((B, 3))
    H[:, 0] = np.repeat([-500e3, -50e3, 0, 50e3, 500e3], B // 5 + 1)[:B]
    b.set_external_field("free", H)
    T = steps * 1e-13 * (1 + 1e-9)
    for mode, flop in (("Heun", 2 * (3 + 9 + 6 + 27 + 41) + 33 + 15 + 54), ("EulerHeun", (3 + 9 + 6 + 27 + 41) + 24 + 15 + 54)):
        sm = getattr(m.SolverMode, mode)
        secs = timed(lambda: b.run_device(T, 1e-13, T, solverMode=sm, log=False, seed=1))
        emit(f"S3 p-bit 1-layer STT 300 K {mode}, final state only", B * steps, secs, float(flop), 0,
             {"instances": B, "steps": steps, "variant": _abi.last_variant()})


def s4(scale):
    """CIMS stability diagram (curated-examples/cims-stability-stability-diagram.py): current x field
    grid, 3-layer stack (tri_layer_mtj couplings), Heun, SOT step drivers with per-instance amplitude."""
    n = int(2048 * math.sqrt(scale))
    B, steps = n * n, 20000
    j = helpers.tri_layer_mtj(m)
    j.setLayerFieldLikeTorqueDriver("free", m.stepDriver(0, 1e3, 0, 10e-9))
    j.setLayerDampingLikeTorqueDriver("free", m.stepDriver(0, 16e3, 0, 10e-9))
    j.getLayer("free").setReferenceLayer(m.CVector(0, 1, 0))
    b = m.JunctionBatch(j, B)
    cur, fld = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-400e3, 400e3, n))
    b.set_override("free", "fieldLike", "amplitude", 1e3 * cur.ravel() ** 2)
    b.set_override("free", "dampingLike", "amplitude", 16e3 * cur.ravel())
    b.set_external_field("all", np.stack([fld.ravel(), np.zeros(B), np.zeros(B)], axis=1))
    T = steps * 1e-12 * (1 + 1e-9)
    secs = timed(lambda: b.run_device(T, 1e-12, T, solverMode=m.SolverMode.Heun, log=False), reps=2)
    emit("S4 CIMS 3-layer Heun, SOT step drivers, final state only", 3 * B * steps, secs, 2 * (3 + 9 + 20 + 27 + 34 + 8) + 33, 0,
         {"instances": B, "grid": [n, n], "steps": steps, "variant": _abi.last_variant()})


def dp(scale):
    """Dormand-Prince (adaptive RK5(4), junction.hpp:866-1002): single-layer SOT junction, field x current
    grid, 5 ns, final state only.  Throughput in accepted+rejected step ATTEMPTS is not observable from the
    host; reported as simulated time per wall second and as equivalent fixed steps of the initial dt."""
    n = int(1024 * math.sqrt(scale))
    B, T, dt = n * n, 5e-9, 1e-12
    j = helpers.sot_layer(m, dynamic=False)
    b = m.JunctionBatch(j, B)
    cur, fld = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-200e3, 200e3, n))
    b.set_override("free", "current", "constantValue", 1e11 * cur.ravel())
    b.set_external_field("free", np.stack([fld.ravel(), np.zeros(B), np.zeros(B)], axis=1))
    secs = timed(lambda: b.run_device(T, dt, T, solverMode=m.SolverMode.DormandPrince, log=False), reps=2)
    emit("Dormand-Prince 1-layer SOT, current x field grid, final state only", B * (T / dt), secs, 7 * (3 + 9 + 27 + 34) + 90, 0,
         {"instances": B, "sim_time_s": T, "initial_dt": dt, "simulated_seconds_per_wall_second": B * T / secs,
          "variant": _abi.last_variant()})


def _lattice(n_side, pitch=100e-9):
    js, pos = [], []
    rng = np.random.default_rng(0)
    for k in range(n_side * n_side):
        mag = rng.normal(0, 0.2, 3) + np.array([0.0, 0.0, 1.0])
        l = m.Layer("free", m.CVector(*mag), m.CVector(0, 0, 1), 1.2, 2e-9, math.pi * (20e-9) ** 2, [m.CVector(0, 0, 0)] * 3, 0.01)
        j = m.Junction([l])
        j.setLayerAnisotropyDriver("free", m.constantDriver(0.3e6))
        js.append(j)
        pos.append(m.CVector(pitch * (k % n_side), pitch * (k // n_side), 0.0))
    return m.reservoir.GroupInteraction(pos, js, "free")


def group_device(g, n_inst, steps, log=False):
    """device-resident run of a GroupInteraction batch through cmtjb_group_run_device"""
    gs, jns, layers, nl, xyz = g._pack()
    a = _coupled.run_args(steps * 1e-13 * (1 + 1e-9), 1e-13, steps * 1e-13 * (1 + 1e-9), n_inst, log=log, seed=1)
    mstate = torch.from_numpy(_coupled.initial_state(g.junctionList, n_inst)).cuda()
    status = torch.zeros(n_inst * g.noElements, dtype=torch.int32, device="cuda")
    f = lib.cmtjb_group_run_device
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]

    def run(keep=(xyz,)):  # the coordinate array must outlive the packed record that points at it
        rc = f(C.byref(gs), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), None, 0, C.byref(a), C.c_void_p(mstate.data_ptr()),
               None, C.c_void_p(status.data_ptr()), 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _abi.cuda_backend().last_error()

    return run


def s5g(scale):
    """Reservoir: 64x64 dipole-coupled macrospins (GroupInteraction semantics), RK4."""
    side = int(64 * math.sqrt(scale))
    N, steps = side * side, 2000
    t0 = time.perf_counter()
    g = _lattice(side)
    run1 = group_device(g, 1, 1)
    host_secs = timed(run1, reps=3)  # one step: validation, model upload and launch latency, no integration to speak of
    run = group_device(g, 1, steps)
    secs = timed(run, reps=2)
    pairs = N * (N - 1) * steps
    emit(f"S5 GroupInteraction {side}x{side} lattice RK4 (cooperative grid kernel)", N * steps, secs, 225.0, 0,
         {"elements": N, "steps": steps, "pair_interactions_per_s": pairs / secs, "pair_flop_frac_of_peak": pairs * 25.0 / secs / PEAK,
          "us_per_step": secs / steps * 1e6, "us_per_step_kernel_only": (secs - host_secs) / steps * 1e6, "call_overhead_s": host_secs,
          "host_setup_s": time.perf_counter() - t0})
    g2 = _lattice(8)
    n_inst = int(2048 * scale)
    run2 = group_device(g2, n_inst, steps)
    secs = timed(run2, reps=2)
    emit("S5 GroupInteraction batch: 8x8 arrays, one block each", 64 * n_inst * steps, secs, 225.0, 0,
         {"arrays": n_inst, "elements": 64, "steps": steps, "pair_interactions_per_s": n_inst * 64 * 63 * steps / secs})


def s5s(scale):
    """Series stack of 64 STT MTJs (single_layer_mtj_fictious-like), coupling 0.1, sine current; a batch
    of stacks with per-stack current amplitude."""
    n_inst, steps, N = int(4096 * scale), 2000, 64
    js = [helpers.stt_layer_mtj(m) for _ in range(N)]
    s = m.stack.SeriesStack(js)
    s.setCouplingStrength(0.1)
    s.setCoupledCurrentDriver(m.ScalarDriver.getSineDriver(1e10, 5e9, 6e9, 0.0))
    st, jns, layers, nl, coup = s._pack()
    a = _coupled.run_args(steps * 1e-13 * (1 + 1e-9), 1e-13, steps * 1e-13 * (1 + 1e-9), n_inst, log=False, seed=1)
    mstate = torch.from_numpy(_coupled.initial_state(s.junctionList, n_inst)).cuda()
    status = torch.zeros(n_inst * N, dtype=torch.int32, device="cuda")
    amps = torch.linspace(1e9, 9e9, n_inst, dtype=torch.float64, device="cuda")
    ov = (_abi.Override * 1)()
    ov[0].layer, ov[0].slot, ov[0].field, ov[0].values = _abi.LAYER_SYSTEM, 0, 1, amps.data_ptr()
    f = lib.cmtjb_stack_run_device
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]

    def run():
        rc = f(C.byref(st), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.cast(ov, C.c_void_p), 1, C.byref(a),
               C.c_void_p(mstate.data_ptr()), None, None, C.c_void_p(status.data_ptr()), 0,
               C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _abi.cuda_backend().last_error()

    secs = timed(run, reps=2)
    emit("S5 SeriesStack of 64 MTJs (KCL), batch of stacks", n_inst * N * steps, secs, 4 * (3 + 9 + 27 + 41) + 45, 0,
         {"stacks": n_inst, "junctions": N, "steps": steps})


def procs(scale):
    """The sweep procedures end to end through their reference-shaped API: host field
    list in, host spectra out.  PIMM: the BASELINE configs[1] shape at reduced length (2-layer MTJ, 1e5 field
    points x 5 ns, FFT of the projected m_z); VSD: 256 fields x 64 frequencies x 10 ns, spin-diode map."""
    P = m.utils.procedures
    rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
          P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
    nH = int(100000 * scale)
    H = np.zeros((nH, 3))
    H[:, 0] = np.linspace(-500e3, 500e3, nH)
    steps = 50000
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        spec, freqs, out = P.PIMM_procedure(helpers.two_layer_mtj(m), H, 1e-13, rp, simulation_duration=steps * 1e-13 * (1 + 1e-9),
                                            disturbance=1e-3, disable_tqdm=True, carry_state=False)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
    emit("PIMM_procedure end to end: 2-layer MTJ, field sweep, 5 ns, spectrum + Rx/Ry on the host", 2 * nH * steps, secs, 337.0, 0,
         {"field_points": nH, "steps": steps, "spectrum_shape": list(np.asarray(spec).shape)})
    nH2, nF, steps2 = int(256 * math.sqrt(scale)), int(64 * math.sqrt(scale)), 100000
    Hv = np.zeros((nH2, 3))
    Hv[:, 0] = np.linspace(50e3, 400e3, nH2)
    Hv[:, 2] = 20e3
    fs = np.linspace(2e9, 20e9, nF)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vm = P.VSD_procedure(helpers.two_layer_mtj(m), Hv, fs, 1e-13, rp, simulation_duration=steps2 * 1e-13 * (1 + 1e-9), Rtype="Rz",
                             disable_tqdm=True, carry_state=False)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
    emit("VSD_procedure end to end: 2-layer MTJ, field x frequency map, 10 ns, mixing + low-pass fused", 2 * nH2 * nF * steps2, secs,
         337.0 + 4 * 10, 0, {"fields": nH2, "frequencies": nF, "steps": steps2, "map_shape": list(vm.shape)})


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.replace(".", "").isdigit()] or ["s2", "s3", "s4", "s5g", "s5s", "dp", "procs"]
    scale = [float(a) for a in sys.argv[1:] if a.replace(".", "").isdigit()]
    scale = scale[0] if scale else 1.0
    print(json.dumps({"fp64_peak_tflops_measured": PEAK / 1e12, "hbm_gbs": HBM}), flush=True)
    for w in which:
        try:
            globals()[w](scale)
        except Exception as e:  # keep going: one config must not hide the others
            print(json.dumps({"config": w, "error": repr(e)[:300]}), flush=True)
