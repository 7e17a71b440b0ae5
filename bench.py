#!/usr/bin/env python
"""bench.py -- headline benchmark of cmtj_b200 (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): PIMM spectrum sweep -- a 2-layer MTJ
(fixture two_layer_mtj of the reference's tests/conftest.py:104-140), 10^5 external-field points
per GPU, 20 ns at dt = 1e-13 s (2x10^5 RK4 steps), step Oersted pulse on z, per-instance Hext and
perturbed initial magnetisation, log of every layer's m and R written every 100 steps.
One bench "step" = one full pass over the batch = 4x10^10 layer-timesteps per GPU.

  value : layer-timesteps/s with inputs resident in HBM (device entry, CUDA events)
  e2e   : the same metric through the host entry (JunctionBatch.run -> cmtjb_run_host) with
          pinned HOST buffers: per-instance inputs H2D and the full log D2H inside the timed region;
          e2e.pipeline says where that time went (cmtjb_host_stats)
  e2e_pimm : the same sweep through the procedure whose DELIVERABLE is the spectrum (PIMM_batch: host
          field list in, host spectra + static resistances out; 1.3 GB D2H instead of 11.2 GB)
  secondary : the other BASELINE configurations at their named sizes (workloads.py), each event-timed
          with its own clock record (N = 1 only)
  --impl reference : the reference's own CPU implementation (oracle/_ref, unmodified
          python/cmtj.cpp build) on all host cores, bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "layer-timesteps/s (FP64)"
FLOP_PER_LAYER_STEP = 337.0  # S1 canonical algorithmic FP64 flop per layer-timestep
N_LAYERS = 2
DT = 1e-13
WRITE_EVERY = 100


def workload(args):
    steps = int(round(args.sim_time / DT))
    return dict(
        workload="PIMM S1: 2-layer MTJ field sweep, RK4, dt=1e-13 s",
        instances_per_gpu=args.instances,
        sim_time_s=args.sim_time,
        steps_per_instance=steps,
        write_every=WRITE_EVERY,
        log="m of both layers + R every 100 steps ([7][%d][B] float64)" % (steps // WRITE_EVERY),
        l2="log output per step (%.1f GB) exceeds the 126 MB L2; no explicit flush"
        % (7 * (steps // WRITE_EVERY) * args.instances * 8 / 1e9),
        parallelism="instances sharded contiguously, one process per GPU, no collective",
    )


def make_inputs(B, seed):
    rng = np.random.default_rng(seed)
    H = np.zeros((B, 3))
    # FieldScan.amplitude_scan(-500e3, 500e3, B, theta=90, phi=0): along x
    H[:, 0] = np.linspace(-500e3, 500e3, B)
    # every point starts saturated along its own field (+-x), 1e-3 off axis: the state a PIMM measurement pulses.
    # (Round 1 started every point 1e-3 from the HARD axis z, an unstable equilibrium: there the reference's own
    # arithmetic amplifies a 1e-12 change of m0 to order one within the run, and no two implementations can agree
    # over 2e5 steps -- tests/test_gpu_fullsize.py keeps that start as its ill-conditioned case.)
    along = np.sign(H[:, :1]) * np.array([[1.0, 0.0, 0.0]])
    m_free = rng.normal(0, 1e-3, (B, 3)) + along
    m_bot = rng.normal(0, 1e-3, (B, 3)) + along
    return H, m_free, m_bot


def build_junction(mod):
    import helpers

    j = helpers.two_layer_mtj(mod, mag=(0.0, 0.0, 1.0))
    oe = mod.ScalarDriver.getStepDriver(0, 50, 0, 3 * DT)  # PIMM_procedure: Hoe_excitation=50, 3 steps
    j.setLayerOerstedFieldDriver("all", mod.AxialDriver(mod.NullDriver(), mod.NullDriver(), oe))
    return j


# ---------------------------------------------------------------------------------------------
# reference arm: the unmodified reference extension on the host cores
# ---------------------------------------------------------------------------------------------
_inputs_cache = {}


def _ref_worker(task):
    lo, hi, B, seed, sim_time = task
    import helpers

    ref = helpers.reference_module()
    if (B, seed) not in _inputs_cache:
        _inputs_cache[(B, seed)] = make_inputs(B, seed)
    H, m_free, m_bot = _inputs_cache[(B, seed)]
    done = 0
    for i in range(lo, hi):
        j = build_junction(ref)
        j.setLayerExternalFieldDriver("all", ref.AxialDriver(*[float(v) for v in H[i]]))
        j.setLayerMagnetisation("free", ref.CVector(*m_free[i]))
        j.setLayerMagnetisation("bottom", ref.CVector(*m_bot[i]))
        j.runSimulation(sim_time, DT, WRITE_EVERY * DT * (1 + 1e-9))
        log = j.getLog()
        done += len(log["time"])
    return done


def run_reference_sample(B, seed, sim_time, n_inst, pool, cores):
    """Time n_inst instances (spread over the sweep) through the reference on `cores` processes."""
    idx = np.linspace(0, B - 1, n_inst).astype(int)
    # contiguous chunks of the selected instances per worker
    chunks = np.array_split(np.arange(n_inst), cores)
    tasks = []
    for ch in chunks:
        for k in ch:
            tasks.append((int(idx[k]), int(idx[k]) + 1, B, seed, sim_time))
    t0 = time.perf_counter()
    pool.map(_ref_worker, tasks, chunksize=max(1, len(tasks) // (cores * 4)))
    dt = time.perf_counter() - t0
    steps = int(round(sim_time / DT))
    return n_inst * steps * N_LAYERS / dt, dt


def reference_arm(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp

    import helpers

    if helpers.reference_module() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_cmtj*.so not built"}))
        return
    cores = os.cpu_count() or 1
    steps = int(round(args.sim_time / DT))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # size the per-step sample for about 8 s of wall time: ~1.1e6 layer-steps/s/core
        n_inst = max(cores, int(8.0 * 1.0e6 * cores / (steps * N_LAYERS)))
        n_inst = (n_inst // cores) * cores
        for _ in range(args.warmup if args.warmup < 2 else 1):
            run_reference_sample(args.instances, 0, args.sim_time, cores, pool, cores)
        vals, times = [], []
        for _ in range(args.steps):
            v, dt = run_reference_sample(args.instances, 0, args.sim_time, n_inst, pool, cores)
            vals.append(v)
            times.append(dt)
    value = float(n_inst * steps * N_LAYERS * len(times) / sum(times))
    sample = f"{n_inst} of the {args.instances} sweep instances per step (evenly spaced), full {steps} RK4 steps each"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "layer-timesteps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload(args),
        "cpu_baseline": {
            "value": value,
            "unit": "layer-timesteps/s",
            "cores": cores,
            "kind": "reference",
            "sample": sample,
        },
        "e2e": {"value": value, "unit": "layer-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 os.environ.get("BENCH_SAMPLER_MS", "100")],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(pw)) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------------
# the spectrum deliverable and the other configurations
# ---------------------------------------------------------------------------------------------
def pimm_leg(args, cm, torch, dist, dev, local_rank, rank, world, barrier, H, steps, lib):
    """configs[1] as its user runs it: PIMM_batch(junction, Hvecs, ...) -> (spectrum [B, n_f], freqs, {Rx, Ry, m_avg}) on
    the host.  The field list goes in as a host array, spectra and static resistances come back as host arrays; the
    2e5-sample projected signals never leave the GPU."""
    import helpers

    P = cm.utils.procedures
    rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
          P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
    B = H.shape[0]
    T = steps * DT * (1 + 1e-9)
    np.random.seed(1234 + rank)  # the procedure draws its disturbances from numpy's global generator, like the reference

    def step():
        j = helpers.two_layer_mtj(cm, mag=(0.0, 0.0, 1.0))
        return P.PIMM_batch(j, H, DT, rp, simulation_duration=T, disturbance=1e-3, disable_tqdm=True, device=local_rank)

    # warm-up like the full-log leg: W calls at least, up to W + 6 while a call is still > 10 % faster than the one before
    # (the first calls page-lock the host result buffers that torch's pinned allocator then reuses, and allocate the
    # device moment storage that its caching allocator then reuses)
    sampler = ClockSampler(local_rank)  # started before the warm-up, see the full-log leg
    sampler.start()
    n_warm, prev = 0, None
    while True:
        t_w = time.perf_counter()
        spectrum, freqs, other = step()
        t_w = time.perf_counter() - t_w
        n_warm += 1
        improving = prev is not None and t_w < 0.9 * prev
        prev = t_w
        if dist is not None:
            flag = torch.tensor([1.0 if improving else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            improving = bool(flag.item() > 0)
        if n_warm >= max(args.warmup, 3) and (not improving or n_warm >= max(args.warmup, 3) + 6):
            break
    barrier()
    t0 = time.perf_counter()
    each = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        spectrum, freqs, other = step()
        each.append(1e3 * (time.perf_counter() - t_s))
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    clocks = sampler.stop(t0, time.perf_counter())
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d2h = int(np.asarray(spectrum).nbytes + sum(np.asarray(other[k]).nbytes for k in ("Rx", "Ry", "m_avg")))
    return {
        "value": B * steps * N_LAYERS * world * args.steps / float(t.item()),
        "unit": "layer-timesteps/s",
        "ms_per_step": 1e3 * float(t.item()) / args.steps,
        "ms_each": each,
        "h2d_bytes_per_step": int((3 + 6) * B * 8),
        "d2h_bytes_per_step": d2h,
        "warmup_steps": n_warm,
        "api": "cmtj_b200.utils.PIMM_batch (= PIMM_procedure(carry_state=False)): projected signal reduced chunk by chunk to Chebyshev block moments on the device, spectra + static resistances to pinned host arrays",
        "spectrum_shape": list(np.asarray(spectrum).shape),
        "checksum_spectrum": float(np.asarray(spectrum)[::997].sum()),
        "clocks": clocks,
    }


def _timed(torch, fn, local_rank, min_warm_s=0.5, reps=3):
    """warm up for >= 0.5 s of wall time, then `reps` event-timed calls with the clocks sampled over them"""
    t0 = time.perf_counter()
    while True:
        fn()
        torch.cuda.synchronize()
        if time.perf_counter() - t0 > min_warm_s:
            break
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.12)
    w0 = time.perf_counter()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    clocks = sampler.stop(w0, time.perf_counter())
    return float(np.median(ts)), clocks


def secondary_block(cm, torch, lib, local_rank):
    """BASELINE configs 1, 3, 4, 5 at their named sizes (workloads.py), device-resident, CUDA events; one entry each."""
    import workloads as wl
    from cmtj_b200 import _abi, _coupled

    hbm = 6650.0
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {}

    def entry(name, units, unit, secs, clocks, **extra):
        out[name] = dict({"value": units / secs, "unit": unit, "seconds": secs, "clocks": clocks}, **extra)

    def guarded(name, fn):
        try:
            fn()
        except Exception as e:  # one configuration must not take the line down
            out[name] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # S2: VSD field scan with the log written EVERY step (HBM-write bound by construction): 40 ns = 4e5 steps; as many
    # field points as a ~100 GB log allows
    def s2():
        steps = 400000
        B = 8192
        j = wl.s2_junction(cm)
        b = cm.JunctionBatch(j, B, device=local_rank)
        b.set_external_field("free", wl.s2_fields(B))
        T, wf = wl.exact(steps), wl.exact(1)
        buf = b.run_device(T, DT, wf)["_log"]
        secs, clocks = _timed(torch, lambda: b.run_device(T, DT, wf, out=buf), local_rank, reps=2)
        nbytes = buf.numel() * 8
        entry("S2 vsd-basic-field-scan: 1-layer, sine Oersted drive, 4e5 RK4 steps, log every step", B * steps, "layer-timesteps/s", secs, clocks,
              instances=B, steps=steps, log_bytes=nbytes, hbm_write_gbs=nbytes / secs / 1e9, hbm_frac=nbytes / secs / 1e9 / hbm,
              hbm_peak_gbs=hbm, variant=_abi.last_variant(),
              note="8192 field points (a 105 GB every-step log) are 256 warps on 592 SM sub-partitions: the run is latency-bound, not HBM-bound")

    guarded("S2", s2)

    # S3: p-bit, 1e6 thermal trials x 5000 steps, final state only
    def s3():
        B, steps = wl.S3_INSTANCES, wl.S3_STEPS
        b = cm.JunctionBatch(wl.s3_junction(cm), B, device=local_rank)
        b.set_external_field("free", wl.s3_fields(B))
        T = wl.exact(steps)
        for mode in ("Heun", "EulerHeun"):
            sm = getattr(cm.SolverMode, mode)
            secs, clocks = _timed(torch, lambda: b.run_device(T, DT, T, solverMode=sm, log=False, seed=1), local_rank)
            entry(f"S3 p-bit: 1-layer STT at 300 K, 1e6 trials x 5000 {mode} steps", B * steps, "layer-timesteps/s", secs, clocks,
                  instances=B, steps=steps, variant=_abi.last_variant())

    guarded("S3", s3)

    # S4: CIMS stability diagram, 2048 x 2048 grid, 3-layer stack, 20 000 Heun steps
    def s4():
        n, steps = wl.S4_GRID, wl.S4_STEPS
        b = wl.s4_batch(cm, n, device=local_rank)
        T = wl.exact(steps, wl.S4_DT)
        secs, clocks = _timed(torch, lambda: b.run_device(T, wl.S4_DT, T, solverMode=cm.SolverMode.Heun, log=False), local_rank, reps=2)
        entry("S4 CIMS stability diagram: 3-layer stack, 2048 x 2048 current x field grid, 20000 Heun steps", 3 * n * n * steps,
              "layer-timesteps/s", secs, clocks, instances=n * n, steps=steps, variant=_abi.last_variant())

    guarded("S4", s4)

    # S5: 64 x 64 dipole-coupled macrospins; 4096 series stacks of 64 MTJs
    def s5g():
        g = wl.s5_lattice(cm, wl.S5G_SIDE)
        N, steps = wl.S5G_SIDE**2, wl.S5G_STEPS
        gs, jns, layers, nl, xyz = g._pack()
        a = _coupled.run_args(wl.exact(steps), DT, wl.exact(steps), 1, log=False, seed=1)
        mstate = torch.from_numpy(_coupled.initial_state(g.junctionList, 1)).cuda()
        status = torch.zeros(N, dtype=torch.int32, device="cuda")
        f = lib.cmtjb_group_run_device
        f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]

        def run(keep=(xyz,)):
            rc = f(C.byref(gs), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), None, 0, C.byref(a), C.c_void_p(mstate.data_ptr()),
                   None, C.c_void_p(status.data_ptr()), local_rank, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, _abi.cuda_backend().last_error()

        secs, clocks = _timed(torch, run, local_rank)
        pairs = N * (N - 1) * steps
        entry(f"S5 reservoir: 64 x 64 dipole-coupled macrospins, {steps} RK4 steps (cooperative grid kernel)", N * steps, "layer-timesteps/s", secs,
              clocks, elements=N, steps=steps, us_per_step=secs / steps * 1e6, pair_interactions_per_s=pairs / secs,
              note="one cooperative launch for the whole run; seconds include the host side of the call (validation and upload of "
                   "4096 junction models, ~17 ms)")

    guarded("S5g", s5g)

    def s5s():
        n_inst, steps, N = wl.S5S_STACKS, wl.S5S_STEPS, wl.S5S_JUNCTIONS
        s = wl.s5_stack(cm)
        st, jns, layers, nl, coup = s._pack()
        a = _coupled.run_args(wl.exact(steps), DT, wl.exact(steps), n_inst, log=False, seed=1)
        mstate = torch.from_numpy(_coupled.initial_state(s.junctionList, n_inst)).cuda()
        status = torch.zeros(n_inst * N, dtype=torch.int32, device="cuda")
        amps = torch.from_numpy(wl.s5_stack_amplitudes(n_inst)).cuda()
        ov = (_abi.Override * 1)()
        ov[0].layer, ov[0].slot, ov[0].field, ov[0].values = _abi.LAYER_SYSTEM, 0, 1, amps.data_ptr()
        f = lib.cmtjb_stack_run_device
        f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]

        def run(keep=(coup,)):
            rc = f(C.byref(st), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), C.cast(ov, C.c_void_p), 1, C.byref(a),
                   C.c_void_p(mstate.data_ptr()), None, None, C.c_void_p(status.data_ptr()), local_rank,
                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, _abi.cuda_backend().last_error()

        secs, clocks = _timed(torch, run, local_rank)
        entry("S5 series stacks: 4096 stacks of 64 STT MTJs (KCL), 2000 RK4 steps", n_inst * N * steps, "junction-timesteps/s", secs, clocks,
              stacks=n_inst, junctions=N, steps=steps)

    guarded("S5s", s5s)

    # the spin-diode procedure end to end (configs[0]'s kind of sweep, as one batch): 256 fields x 64 frequencies x 10 ns
    def vsd():
        import helpers

        P = cm.utils.procedures
        rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
              P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
        nH, nF, steps = 256, 64, 100000
        Hv = np.zeros((nH, 3))
        Hv[:, 0] = np.linspace(50e3, 400e3, nH)
        Hv[:, 2] = 20e3
        fs = np.linspace(2e9, 20e9, nF)
        shape = []

        def run():
            vm = P.VSD_batch(helpers.two_layer_mtj(cm), Hv, fs, DT, rp, simulation_duration=wl.exact(steps), Rtype="Rz", disable_tqdm=True,
                             device=local_rank)
            shape[:] = list(np.asarray(vm).shape)

        secs, clocks = _timed(torch, run, local_rank)
        entry("VSD_batch end to end: 2-layer MTJ, 256 fields x 64 frequencies x 10 ns, spin-diode map on the host (mixing, low-pass "
              "and mean inside the time loop)", N_LAYERS * nH * nF * steps, "layer-timesteps/s", secs, clocks, instances=nH * nF, steps=steps,
              map_shape=shape, variant=_abi.last_variant())

    guarded("VSD", vsd)
    return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cmtj_b200", choices=["cmtj_b200", "reference"])
    ap.add_argument("--instances", type=int, default=100000, help="sweep instances per GPU")
    ap.add_argument("--sim-time", type=float, default=20e-9)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pimm", action="store_true", help="skip the spectrum-deliverable end-to-end leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch

    import cmtj_b200
    from cmtj_b200 import _abi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION / WARN
        # (this image exports NCCL_DEBUG).  Only those two levels are silenced; INFO / TRACE stay as asked for.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        import torch.distributed as dist

        # ... and whatever native code still writes to fd 1 while the communicator comes up goes to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    lib = _abi.load_library()
    lib.cmtjb_launch_count.restype = C.c_longlong

    B = args.instances
    steps = int(round(args.sim_time / DT))
    T = args.sim_time
    wf = WRITE_EVERY * DT * (1 + 1e-9)  # (unsigned)(wf/dt) == 100 exactly (SURVEY 8a note 2)
    units_per_step = B * steps * N_LAYERS  # layer-timesteps per GPU per bench step

    j = build_junction(cmtj_b200)
    H, m_free, m_bot = make_inputs(B, seed=rank)

    def fresh_batch():
        b = cmtj_b200.JunctionBatch(j, B, device=local_rank, instance_offset=rank * B)
        b.set_external_field("all", H)
        b.set_magnetisation("free", m_free)
        b.set_magnetisation("bottom", m_bot)
        return b

    # ---- FP64 roofline denominator: measured live, DFMA microbenchmark on every SM -----------------
    flops, ms = C.c_double(), C.c_double()
    lib.cmtjb_fp64_peak(local_rank, C.c_int64(100000), C.byref(flops), C.byref(ms))
    fp64_peak = flops.value
    lib.cmtjb_fp64_peak_sm_mhz.restype = C.c_double
    fp64_peak_mhz = float(lib.cmtjb_fp64_peak_sm_mhz())
    # what the pipe sustains by operand shape (cmtjb_fp64_pipe_probe): independent DFMA chains on fixed
    # register triplets at 8 resident warps per sub-partition (the pipe's practical peak) and a loop of the LLG right-hand
    # side's shape -- three different live register pairs per DFMA -- at the S1 kernel's 4 warps (its ceiling)
    pipe_probe = {}
    pf = lib.cmtjb_fp64_pipe_probe
    pf.restype = C.c_int
    pf.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name, kind, warps, iters in (("dfma_3reg_8warps", 1, 8, 100000), ("rhs_shape_4warps", 3, 4, 200000), ("rhs_shape_two_vectors_4warps", 4, 4, 200000)):
        fr, mh = C.c_double(), C.c_double()
        if pf(local_rank, kind, warps, iters, C.byref(fr), C.byref(mh)) == 0:
            pipe_probe[name] = round(fr.value, 4)

    # ---- device-resident leg ("value") --------------------------------------------------------------------
    b = fresh_batch()
    res = b.run_device(T, DT, wf)  # allocates the log once; also the first warm-up
    logbuf = res["_log"]
    n_samples = logbuf.shape[1]
    m0_dev = b._m_dev.clone()
    variant = b.last_plan.variant.decode()

    def one_step():
        b._m_dev.copy_(m0_dev)  # every step integrates the same 20 ns from the same initial state
        b.run_device(T, DT, wf, out=logbuf)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup - 1, 0)):
        one_step()
    barrier()
    stream = torch.cuda.current_stream(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.cmtjb_launch_count()
    t_wall0 = time.perf_counter()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record(stream)
    kern_ms = []
    for e0, e1 in evs:
        b._m_dev.copy_(m0_dev)
        e0.record(stream)
        b.run_device(T, DT, wf, out=logbuf)
        e1.record(stream)
    e_all1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.cmtjb_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    total_ms = e_all0.elapsed_time(e_all1)
    kern_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = units_per_step * world * args.steps / (total_ms_max * 1e-3)

    # ---- end-to-end leg: host entry, pinned host buffers, H2D + D2H inside the timed region ----------
    e2e = None
    if not args.no_e2e:
        del res
        # pinned pages land on the NUMA node of the allocating thread: allocate from the GPU-local CPUs
        from cmtj_b200 import _numa

        prev_aff = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        bound = _numa.bind_host_to_gpu(local_rank)
        host_log_t = torch.empty((7, n_samples, B), dtype=torch.float64, pin_memory=True)
        host_log_t.zero_()  # first touch while bound
        host_log = host_log_t.numpy()
        if bound is not None:
            os.sched_setaffinity(0, prev_aff)
        # raw D2H rate of this box's PCIe path into that buffer (diagnostic: e2e is copy-bound when
        # d2h_bytes / this rate exceeds the kernel time)
        probe = logbuf[0]
        ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_log_t[0].copy_(probe, non_blocking=True)
        torch.cuda.synchronize(dev)
        ec0.record(stream)
        host_log_t[0].copy_(probe, non_blocking=True)
        ec1.record(stream)
        torch.cuda.synchronize(dev)
        d2h_gbs = probe.numel() * 8 / (ec0.elapsed_time(ec1) * 1e-3) / 1e9
        bh = fresh_batch()
        h2d = (3 + 6) * B * 8
        d2h = host_log.nbytes + 6 * B * 8 + B * 4
        m_init = bh.m_state.copy()

        def e2e_step():
            bh.m_state[...] = m_init
            return bh.run(T, DT, wf, out=host_log)

        # warm-up: the pinned allocation above leaves the GPU idle for seconds; W untimed steps bring the
        # clocks back up before the timed region (sampled below like the device leg).  The clock sampler starts BEFORE the
        # warm-up: nvidia-smi's start-up (NVML initialisation over every GPU of the box) holds driver locks for some hundred
        # milliseconds, which used to land on the first timed step (its copies stalled: 750-1090 ms against 380 ms for the
        # steps after it); its rows are filtered to the timed region by time stamp.
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        n_warm, prev = 0, None
        while True:
            t_w = time.perf_counter()
            e2e_step()
            t_w = time.perf_counter() - t_w
            n_warm += 1
            # W steps at least; with several ranks sharing one host the first passes over the pinned buffers are
            # much slower than the steady state (1343 -> 912 -> 628 ms observed at N = 8): keep warming, up to
            # W + 6 steps, while a step is still more than 10 % faster than the one before (all ranks agree)
            improving = prev is not None and t_w < 0.9 * prev
            prev = t_w
            if dist is not None:
                flag = torch.tensor([1.0 if improving else 0.0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                improving = bool(flag.item() > 0)
            if n_warm >= max(args.warmup, 3) and (not improving or n_warm >= max(args.warmup, 3) + 6):
                break
        barrier()
        t0 = time.perf_counter()
        each = []
        stats = np.zeros((args.steps, 9))
        lib.cmtjb_host_stats.argtypes = [C.c_void_p, C.c_int32]
        for k_step in range(args.steps):
            t_s = time.perf_counter()
            out = e2e_step()
            each.append(1e3 * (time.perf_counter() - t_s))
            lib.cmtjb_host_stats(stats[k_step].ctypes.data, 9)
        torch.cuda.synchronize(dev)
        dt_e2e = time.perf_counter() - t0
        clocks_e2e = sampler2.stop(t0, time.perf_counter())
        checksum = float(out["free_mz"][-1].sum())  # the result is read on the host
        t = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {
            "value": units_per_step * world * args.steps / float(t.item()),
            "unit": "layer-timesteps/s",
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h),
            "ms_per_step": 1e3 * float(t.item()) / args.steps,
            "api": "JunctionBatch.run -> cmtjb_run_host (pinned host log buffer, time-chunked D2H overlap)",
            "checksum_free_mz_last": checksum,
            "ms_each": each,
            "warmup_steps": n_warm,
            "clocks": clocks_e2e,
            "pcie_d2h_gbs_measured": d2h_gbs,
            # where a step's time goes (cmtjb_host_stats, mean over the timed steps, this rank): the integration
            # kernels, gaps on the compute stream while a ring buffer was still draining to the host, the copy engine's
            # busy time, and the part of the last chunk's copy nothing could hide
            "pipeline": {
                "total_ms": float(stats[:, 0].mean()),
                "integrate_ms": float(stats[:, 1].mean()),
                "ring_wait_ms": float(stats[:, 2].mean()),
                "copy_busy_ms": float(stats[:, 3].mean()),
                "tail_copy_ms": float(stats[:, 4].mean()),
                "device_pipeline_ms": float(stats[:, 5].mean()),
                "d2h_gbs_sustained": float(stats[:, 6].mean() / max(stats[:, 3].mean(), 1e-9) / 1e6),
                "time_chunks": int(stats[0, 7]),
                "ring_buffers": int(stats[0, 8]),
            },
            "host_numa_binding": f"{len(bound)} GPU-local CPUs during pinned allocation" if bound else "none (single node or unknown)",
        }
        lib.cmtjb_release_cache()

    # ---- end-to-end leg 2: the spectrum deliverable (PIMM_batch) ------------------------------------------------------
    e2e_pimm = None
    if not args.no_e2e and not args.no_pimm:
        del logbuf, m0_dev, b
        torch.cuda.empty_cache()
        e2e_pimm = pimm_leg(args, cmtj_b200, torch, dist, dev, local_rank, rank, world, barrier, H, steps, lib)
        torch.cuda.empty_cache()

    # ---- the other BASELINE configurations (N = 1 only) ---------------------------------------------------------------
    secondary = None
    if world == 1 and not args.no_secondary:
        try:
            del logbuf, m0_dev, b
        except NameError:
            pass
        torch.cuda.empty_cache()
        secondary = secondary_block(cmtj_b200, torch, lib, local_rank)
        torch.cuda.empty_cache()

    # ---- CPU baseline + parity gate (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp

        import helpers

        # parity gate: 16 instances spread over the sweep, ALL 200 000 steps, every logged sample, CUDA vs the oracle
        # restatement; three more oracle runs perturb one component of m0 by 1e-12 each: how far the REFERENCE arithmetic
        # itself moves (tests/parity_report.py: the bound is max(1e-9, 10 x that response) until the response reaches 1e-3)
        from parity_report import compare_with_sensitivity

        sel = np.linspace(0, B - 1, 16).astype(int)
        logs = []
        for be, axis in ((None, None), (helpers.oracle_backend(), None), (helpers.oracle_backend(), 0), (helpers.oracle_backend(), 1),
                         (helpers.oracle_backend(), 2)):
            with helpers.facade_on(be):
                bb = cmtj_b200.JunctionBatch(j, len(sel))
                bb.set_external_field("all", H[sel])
                mf = m_free[sel].copy()
                if axis is not None:
                    mf[:, axis] = mf[:, axis] * (1 + 1e-12) + 1e-12
                bb.set_magnetisation("free", mf)
                bb.set_magnetisation("bottom", m_bot[sel])
                logs.append(bb.run(T, DT, wf))
        keys = [k for k in logs[1] if k != "time"]
        try:
            parity = dict(compare_with_sensitivity("bench.py parity gate: 16 of the sweep's instances, all %d RK4 steps" % steps, keys,
                                                   logs[0], logs[1], logs[2:]))
            parity["gate"] = "passed"
        except AssertionError as e:  # reported in the line (and on stderr), so that the measurement is not lost with it
            parity = {"gate": "FAILED", "detail": "CUDA vs oracle outside max(1e-9, 10 x reference sensitivity): %r" % (e.args,)}
            print("bench.py parity gate FAILED: %s" % parity["detail"], file=sys.stderr)
        parity.pop("test", None)
        parity["note"] = ("max_rel_err_of_the_others: the instances that meet 1e-9 outright; the rest sit where a 1e-12 change of m0 moves the "
                          "REFERENCE itself by more than 1e-10 (reference_sensitivity_per_instance): asserted against 10 x that response")
        if helpers.reference_module() is not None:
            cores = os.cpu_count() or 1
            ctx = mp.get_context("fork")
            n_inst = max(cores, int(12.0 * 1.0e6 * cores / (steps * N_LAYERS)))
            n_inst = (n_inst // cores) * cores
            with ctx.Pool(cores) as pool:
                run_reference_sample(B, 0, T, cores, pool, cores)
                v, dt = run_reference_sample(B, 0, T, n_inst, pool, cores)
            cpu = {
                "value": v,
                "unit": "layer-timesteps/s",
                "cores": cores,
                "kind": "reference",
                "sample": f"{n_inst} of the {B} sweep instances (evenly spaced), full {steps} RK4 steps each, "
                f"multiprocessing.Pool({cores}), {dt:.1f} s",
            }
        else:
            cpu = {"value": None, "unit": "layer-timesteps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    if rank == 0:
        kern_s = float(np.mean(kern_ms)) * 1e-3
        rate = units_per_step / kern_s
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        traffic, executed, frac = None, None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        # the FP64 pipe issues 64 lanes per SM and clock: the roofline of this kernel, at the clock sampled while it ran
        nominal = 2.0 * 64.0 * n_sm * sm_mhz * 1e6
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            if tj.get("fp64_thread_inst_per_layer_step"):
                # physical utilisation of the pipe: executed FP64 thread instructions (instruction mix of the committed ncu
                # capture of this kernel) x the rate measured here, over the pipe's issue slots
                frac = tj["fp64_thread_inst_per_layer_step"] * rate / (nominal / 2.0)
                executed = {
                    "fp64_thread_inst_per_layer_step": tj["fp64_thread_inst_per_layer_step"],
                    "flop_per_layer_step": tj.get("executed_flop_per_layer_step"),
                    "tflops": (tj.get("executed_flop_per_layer_step") or 0.0) * rate / 1e12,
                    "source": tj.get("source"),
                }
        log_bytes = 7 * n_samples * B * 8
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "layer-timesteps/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": dict(workload(args), kernel_variant=variant, global_instances=B * world),
            "clocks": clocks,
            "e2e": e2e,
            "e2e_pimm": e2e_pimm,
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64",
                # achieved / peak: FP64 thread instructions the kernel executes per second over the pipe's issue slots
                # (instructions, not flop: a DMUL or DADD occupies the pipe like a DFMA)
                "achieved": (executed["fp64_thread_inst_per_layer_step"] * rate / 1e12) if executed else None,
                "peak": nominal / 2.0 / 1e12,
                "unit": "T FP64 thread-instructions/s",
                "frac": frac,
                "traffic": traffic,
                "peak_source": "64 FP64 lanes x %d SMs x the SM clock sampled during the timed region (%.0f MHz); the dependent-DFMA "
                "microbenchmark cmtjb_fp64_peak reaches %.1f TFLOP/s at an in-kernel clock of %.0f MHz = %.3f of 2 x 64 x SMs x that clock"
                % (n_sm, sm_mhz, fp64_peak / 1e12, fp64_peak_mhz, fp64_peak / (2.0 * 64.0 * n_sm * max(fp64_peak_mhz, 1.0) * 1e6)),
                "kernel_ms": kern_s * 1e3,
                "executed": executed,
                "algorithmic": {
                    "flop_per_layer_step": FLOP_PER_LAYER_STEP,
                    "tflops": units_per_step * FLOP_PER_LAYER_STEP / kern_s / 1e12,
                    "note": "SURVEY 8(d) count of the un-hoisted reference arithmetic (full demag mat-vec, divides); the kernel "
                    "executes fewer flop for the same result, so this figure may exceed the pipe's peak and is not a utilisation",
                },
                "fp64_peak_microbenchmark": {"tflops": fp64_peak / 1e12, "sm_mhz_in_kernel": fp64_peak_mhz,
                                             "nominal_tflops_at_sampled_clock": nominal / 1e12},
                "pipe_probe_issue_slot_frac": pipe_probe,
                "frac_of_rhs_shape_ceiling": (frac / max(pipe_probe.get("rhs_shape_4warps", 0), pipe_probe.get("rhs_shape_two_vectors_4warps", 0)))
                if frac and pipe_probe.get("rhs_shape_4warps") else None,
                "hbm": {
                    "achieved_gbs": log_bytes / kern_s / 1e9,
                    "peak_gbs": hbm_peak,
                    "frac": log_bytes / kern_s / 1e9 / hbm_peak,
                    "note": "log writes only; this run is FP64-pipe bound, not HBM bound",
                },
            },
            "cpu_baseline": cpu,
            "parity_gate": parity,
            "secondary": secondary,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
