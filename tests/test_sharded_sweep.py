"""N > 1 path on CPU: a world_size-2 `gloo` job (torch.distributed.run) runs a sharded sweep and
gathers on rank 0; the result must equal the same sweep as ONE batch, bit for bit -- including the
thermal case, because noise streams are keyed by the GLOBAL instance index."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cmtj_b200
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100000, 2**20 + 3):
        for world in (1, 2, 3, 8):
            r = [cmtj_b200.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("mode", ["rk4", "thermal"])
def test_two_rank_gloo_sweep_equals_single_batch(tmp_path, on_oracle, mode):
    out = tmp_path / "gathered.npz"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_sharded_worker.py"), str(out), mode]
    subprocess.run(cmd, check=True, env=env, timeout=600, cwd=ROOT)
    got = np.load(out)

    B = 11
    rng = np.random.default_rng(0)
    H = np.stack([np.linspace(-300e3, 300e3, B), np.full(B, 1e4), np.zeros(B)], axis=1)
    m0 = rng.normal(0, 0.1, (B, 3)) + np.array([0, 0, 1.0])
    if mode == "thermal":
        j = helpers.stt_layer_mtj(cmtj_b200, temperature=300.0)
        solver = cmtj_b200.SolverMode.Heun
    else:
        j = helpers.two_layer_mtj(cmtj_b200)
        solver = cmtj_b200.SolverMode.RK4
    b = cmtj_b200.JunctionBatch(j, B)
    b.set_external_field("all", H)
    b.set_magnetisation("free", m0)
    want = b.run(0.1e-9, 1e-13, 1e-12, solverMode=solver, seed=5)
    for k, v in want.items():
        assert np.array_equal(got[k], v), k
    assert np.array_equal(got["m_final"], b.get_magnetisation("free"))
