#!/usr/bin/env python
"""bench.py -- headline benchmark of cmtj_b200 (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): PIMM spectrum sweep -- a 2-layer MTJ
(fixture two_layer_mtj of the reference's tests/conftest.py:104-140), 10^5 external-field points
per GPU, 20 ns at dt = 1e-13 s (2x10^5 RK4 steps), step Oersted pulse on z, per-instance Hext and
perturbed initial magnetisation, log of every layer's m and R written every 100 steps.
One bench "step" = one full pass over the batch = 4x10^10 layer-timesteps per GPU.

  value : layer-timesteps/s with inputs resident in HBM (device entry, CUDA events)
  e2e   : the same metric through the host entry (JunctionBatch.run -> cmtjb_run_host) with
          pinned HOST buffers: per-instance inputs H2D and the full log D2H inside the timed region
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
    m_free = rng.normal(0, 1e-3, (B, 3)) + np.array([0.0, 0.0, 1.0])
    m_bot = rng.normal(0, 1e-3, (B, 3)) + np.array([0.0, 0.0, 1.0])
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
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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

    for _ in range(max(args.warmup - 1, 0)):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
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
        # clocks back up before the timed region (sampled below like the device leg)
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
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        barrier()
        t0 = time.perf_counter()
        each = []
        for _ in range(args.steps):
            t_s = time.perf_counter()
            out = e2e_step()
            each.append(1e3 * (time.perf_counter() - t_s))
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
            "host_numa_binding": f"{len(bound)} GPU-local CPUs during pinned allocation" if bound else "none (single node or unknown)",
        }
        lib.cmtjb_release_cache()

    # ---- CPU baseline + parity gate (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp

        import helpers

        # parity gate: 64 random instances, first 2000 steps, CUDA vs the oracle restatement
        sel = np.sort(np.random.default_rng(123).choice(B, size=64, replace=False))
        logs = []
        for be in (None, helpers.oracle_backend()):
            with helpers.facade_on(be):
                bb = cmtj_b200.JunctionBatch(j, 64)
                bb.set_external_field("all", H[sel])
                bb.set_magnetisation("free", m_free[sel])
                bb.set_magnetisation("bottom", m_bot[sel])
                logs.append(bb.run(2000 * DT * (1 + 1e-9), DT, 10 * DT * (1 + 1e-9)))
        worst = 0.0
        for k in logs[1]:
            if k != "time":
                worst = max(worst, helpers.max_rel_err(logs[0][k], logs[1][k]))
        parity = {"instances": 64, "steps": 2000, "max_rel_err_vs_oracle": worst, "bound": 1e-9}
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
        achieved = units_per_step * FLOP_PER_LAYER_STEP / kern_s
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        traffic = None
        executed = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            if tj.get("fp64_thread_inst_per_layer_step") and clocks.get("sm_mhz"):
                # physical utilisation of the FP64 pipe: executed FP64 thread instructions (instruction mix of
                # the committed ncu capture) over 64 lanes x SMs x the SM clock sampled during the timed region
                n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                slots = 64.0 * n_sm * clocks["sm_mhz"] * 1e6
                rate = units_per_step / kern_s
                executed = {
                    "fp64_thread_inst_per_layer_step": tj["fp64_thread_inst_per_layer_step"],
                    "flop_per_layer_step": tj.get("executed_flop_per_layer_step"),
                    "issue_slot_frac": tj["fp64_thread_inst_per_layer_step"] * rate / slots,
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
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64",
                "achieved": achieved / 1e12,
                "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "traffic": traffic,
                "peak_source": "measured live by cmtjb_fp64_peak (dependent-DFMA microbenchmark on all SMs); "
                "MEASURED_PEAKS.json carries no FP64 figure",
                "algorithmic_flop_per_layer_step": FLOP_PER_LAYER_STEP,
                "kernel_ms": kern_s * 1e3,
                "executed": executed,
                "hbm": {
                    "achieved_gbs": log_bytes / kern_s / 1e9,
                    "peak_gbs": hbm_peak,
                    "frac": log_bytes / kern_s / 1e9 / hbm_peak,
                    "note": "log writes only; this run is FP64-pipe bound, not HBM bound",
                },
            },
            "cpu_baseline": cpu,
            "parity_gate": parity,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
