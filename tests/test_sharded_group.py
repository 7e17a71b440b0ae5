"""One dipole-coupled array split over several ranks (``GroupInteraction.run_sharded`` ->
``cmtjb_group_run_sharded``, core/reservoir.hpp:129-152 with the element loop cut into ranges).

CPU (``-m "not gpu"``): the oracle's host-memory stand-in for the library runs the protocol, (a) with the ranks as
threads of this process and a barrier-based exchange, (b) as a world_size-2 ``gloo`` job with the default collective
exchange.  Both must reproduce the unsharded oracle run bit for bit (same summation order).

GPU (``-m gpu``): the CUDA path at world = 1 against the oracle (<= 1e-9), and at world = 2 / 3 / 4 -- ranks as threads
sharing the one GPU, exchange by device copies -- bit-identical to world = 1 (the pair sums are sliced by global index,
so the summation order does not depend on the number of ranks)."""
import math
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import cmtj_b200
import helpers

m = cmtj_b200
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lattice_group(n, temperature=0.0, pitch=100e-9, seed=11, two_layers=False):
    """n macrospins on a near-square lattice, random tilts around +z (a regular, non-chaotic array)."""
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n)))
    js, pos = [], []
    for k in range(n):
        mag = rng.normal(0, 0.2, 3) + np.array([0.0, 0.0, 1.0])
        layers = [m.Layer("free", m.CVector(*mag), m.CVector(0, 0, 1), 1.2, 2e-9, math.pi * (20e-9) ** 2, [m.CVector(0, 0, 0)] * 3, 0.01)]
        if two_layers:
            layers.append(m.Layer("bottom", m.CVector(0.1, 0, 1), m.CVector(0, 0, 1), 1.0, 3e-9, math.pi * (20e-9) ** 2, [m.CVector(0, 0, 0)] * 3, 0.02))
        j = m.Junction(layers)
        j.setLayerAnisotropyDriver("all", m.ScalarDriver.getConstantDriver(0.3e6))
        if two_layers:
            j.setIECDriver("free", "bottom", m.ScalarDriver.getConstantDriver(1e-5))
        if temperature:
            j.setLayerTemperatureDriver("all", m.ScalarDriver.getConstantDriver(temperature))
        js.append(j)
        pos.append(m.CVector(pitch * (k % side), pitch * (k // side), 0.0))
    return m.reservoir.GroupInteraction(pos, js, "free")


def logs_of(g):
    return [helpers.log_arrays(g.getLog(k)) for k in range(g.noElements)]


def threaded_ranks(world, make, run_kwargs, exchange_factory, T=0.05e-9):
    """`world` ranks as threads; returns every rank's GroupInteraction (own junctions written back)."""
    groups = [make() for _ in range(world)]
    errs = []
    shared = {"bufs": [None] * world}
    barrier = threading.Barrier(world)

    def work(r):
        try:
            groups[r].run_sharded(T, 1e-13, 1e-12, rank=r, world=world, exchange=exchange_factory(r, world, shared, barrier),
                                  gather=False, seed=7, **run_kwargs)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return groups


def host_exchange_factory(rank, world, shared, barrier):
    import ctypes as C

    def exchange(buf, count, step, stream):
        shared["bufs"][rank] = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(world * count,))
        barrier.wait()
        for r in range(world):
            if r != rank:
                shared["bufs"][rank][r * count : (r + 1) * count] = shared["bufs"][r][r * count : (r + 1) * count]
        barrier.wait()
        return 0

    return exchange


def merged_logs(groups, world):
    n = groups[0].noElements
    per = -(-n // world)
    return [helpers.log_arrays(groups[min(k // per, world - 1)].getLog(k)) for k in range(n)]


@pytest.mark.parametrize("world,n,inter,temperature,two", [
    (2, 9, None, 0.0, False),
    (3, 10, "computeDipoleInteractionNoumra", 0.0, True),  # ragged: 4 + 4 + 2
    (4, 5, None, 0.0, False),     # per = 2: the last rank owns nothing
    (2, 8, None, 300.0, False),   # thermal: noise keyed by the global element index
])
def test_threaded_ranks_equal_the_unsharded_oracle(on_oracle, world, n, inter, temperature, two):
    def make():
        g = lattice_group(n, temperature, two_layers=two)
        if inter:
            g.setInteractionFunction(getattr(m.reservoir, inter))
        return g

    whole = make()
    m.seed(3)
    import cmtj_b200._coupled as cc

    # the unsharded run with the same Philox / mt19937 key as the sharded one
    be = helpers.oracle_backend()
    gq, jns, layers, nl, xyz = whole._pack()
    a = cc.run_args(0.05e-9, 1e-13, 1e-12, 1, seed=7)
    plan = be.group_plan(gq, jns, layers, a)
    st = cc.initial_state(whole.junctionList, 1)
    jlog = np.zeros((int(plan.n_junction_channels), int(plan.n_samples), n))
    be.check(be.group_run_host(gq, jns, layers, None, 0, a, st, jlog, np.zeros(n, dtype=np.int32)))
    groups = threaded_ranks(world, make, {"_host_memory": True}, host_exchange_factory)
    got = merged_logs(groups, world)
    names = whole.junctionList[0]._channel_names()
    for k in range(n):
        for ch, name in enumerate(names):
            assert np.array_equal(got[k][name], jlog[ch, :, k]), (k, name)


def test_two_rank_gloo_job_equals_the_unsharded_run(tmp_path, on_oracle):
    out = tmp_path / "group.npz"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 31500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_sharded_group_worker.py"), str(out)]
    subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)
    got = np.load(out)
    g = lattice_group(9)
    g.runSimulation(0.05e-9, 1e-13, 1e-12)
    for k, log in enumerate(logs_of(g)):
        for name, v in log.items():
            assert np.array_equal(got[f"r0/{k}/{name}"], v), (k, name)
            assert np.array_equal(got[f"r1/{k}/{name}"], v), (k, name)  # gather=True: every rank holds every log


# ------------------------------------------------------------------------------------------------------------
def device_exchange_factory(rank, world, shared, barrier):
    import torch

    from cmtj_b200.reservoir import _DeviceSpan

    def exchange(buf, count, step, stream):
        torch.cuda.ExternalStream(stream).synchronize() if stream else torch.cuda.synchronize()
        shared["bufs"][rank] = torch.as_tensor(_DeviceSpan(buf, world * count), device="cuda:0")
        barrier.wait()
        for r in range(world):
            if r != rank:
                shared["bufs"][rank][r * count : (r + 1) * count].copy_(shared["bufs"][r][r * count : (r + 1) * count])
        torch.cuda.synchronize()
        barrier.wait()
        return 0

    return exchange


@pytest.mark.gpu
@pytest.mark.parametrize("n,inter,two", [(400, None, False), (150, "computeDipoleInteractionNoumra", True)])
def test_sharded_cuda_array_matches_the_oracle_and_any_world_size(n, inter, two):
    from parity_report import record

    def make():
        g = lattice_group(n, two_layers=two)
        if inter:
            g.setInteractionFunction(getattr(m.reservoir, inter))
        return g

    T = 0.03e-9
    with helpers.facade_on(helpers.oracle_backend()):
        want = make()
        want.runSimulation(T, 1e-13, 1e-12)
        want = logs_of(want)
    one = make()
    one.run_sharded(T, 1e-13, 1e-12, rank=0, world=1, seed=7)
    one = logs_of(one)
    worst = 0.0
    for k in range(n):
        for name, v in want[k].items():
            if name != "time":
                worst = max(worst, helpers.max_rel_err(one[k][name], v))
    record("sharded group, world 1 vs oracle", max_rel_err=worst, bound_1e9=1e-9, max_bound_applied=1e-9, widened=False)
    assert worst <= 1e-9, worst
    for world in (2, 3, 4):
        got = merged_logs(threaded_ranks(world, make, {}, device_exchange_factory, T=T), world)
        for k in range(n):
            for name, v in one[k].items():
                assert np.array_equal(got[k][name], v), (world, k, name)


@pytest.mark.gpu
def test_sharded_cuda_array_thermal_is_keyed_by_global_element():
    def make():
        return lattice_group(96, temperature=300.0)

    one = make()
    one.run_sharded(0.02e-9, 1e-13, 1e-12, rank=0, world=1, seed=7)
    one = logs_of(one)
    got = merged_logs(threaded_ranks(3, make, {}, device_exchange_factory, T=0.02e-9), 3)
    for k in range(96):
        for name, v in one[k].items():
            assert np.array_equal(got[k][name], v), (k, name)
    assert np.std(one[5]["free_mx"]) > 0


# ------------------------------------------------------------------------------------------------------------
def thread_peer_areas_factory(shared, barrier):
    """ranks as threads of one process: the 'peers' are each other's plain device pointers (no IPC needed)"""
    import ctypes as C

    from cmtj_b200.reservoir import _PeerAreas

    def make(be, dist, rank, world, device, n):
        lib = be.lib
        lib.cmtjb_group_peer_area_bytes.restype = C.c_int64
        own = C.c_void_p()
        be.check(lib.cmtjb_peer_alloc(C.c_int(device), C.c_int64(lib.cmtjb_group_peer_area_bytes(C.c_int32(n), C.c_int32(world))), C.byref(own)))
        shared["bufs"][rank] = own.value
        barrier.wait()
        return _PeerAreas(be, device, own.value, list(shared["bufs"]), [], barrier.wait)

    return make


def threaded_peer_ranks(world, make, T):
    groups = [make() for _ in range(world)]
    errs = []
    shared = {"bufs": [None] * world}
    barrier = threading.Barrier(world)
    factory = thread_peer_areas_factory(shared, barrier)

    def work(r):
        try:
            groups[r].run_sharded(T, 1e-13, 1e-12, rank=r, world=world, transport="peer", gather=False, seed=7, _peer_areas=factory)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return groups


@pytest.mark.gpu
@pytest.mark.parametrize("n,world,inter,two,temperature", [
    (400, 2, None, False, 0.0),
    (150, 3, "computeDipoleInteractionNoumra", True, 0.0),
    (5, 4, None, False, 0.0),       # the last rank owns nothing and still takes part in the hand-shake
    (96, 3, None, False, 300.0),    # thermal noise keyed by the global element index
])
def test_peer_memory_exchange_is_bit_identical_to_one_rank(monkeypatch, n, world, inter, two, temperature):
    """cmtjb_group_run_sharded_peer: the step kernels push their slice into every rank's exchange area and hand-shake
    through arrival counters (ranks = threads sharing this GPU, each on its own stream); same bits as world = 1."""
    monkeypatch.setenv("CMTJB_PEER_TIMEOUT_MS", "8000")  # a scheduling dead-end on the shared GPU fails instead of hanging

    def make():
        g = lattice_group(n, temperature, two_layers=two)
        if inter:
            g.setInteractionFunction(getattr(m.reservoir, inter))
        return g

    T = 0.03e-9
    one = make()
    one.run_sharded(T, 1e-13, 1e-12, rank=0, world=1, seed=7)
    one = logs_of(one)
    got = merged_logs(threaded_peer_ranks(world, make, T), world)
    for k in range(n):
        for name, v in one[k].items():
            assert np.array_equal(got[k][name], v), (world, k, name)
    # world = 1 through the peer path as well (an area of one's own)
    alone = threaded_peer_ranks(1, make, T)
    for k in range(n):
        for name, v in one[k].items():
            assert np.array_equal(logs_of(alone[0])[k][name], v), (k, name)
