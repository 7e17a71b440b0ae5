"""Worker of tests/test_sharded_sweep.py: one rank of a world_size-2 gloo job on CPU.  The oracle
stands in for the device library (no GPU here); what is under test is the host-side sharding,
instance_offset keying and the rank-0 gather of cmtj_b200/sweep.py."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmtj_b200  # noqa: E402
import helpers  # noqa: E402


def main():
    out_path, mode = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B = 11  # not divisible by the world size
    rng = np.random.default_rng(0)
    H = np.stack([np.linspace(-300e3, 300e3, B), np.full(B, 1e4), np.zeros(B)], axis=1)
    m0 = rng.normal(0, 0.1, (B, 3)) + np.array([0, 0, 1.0])
    with helpers.facade_on(helpers.oracle_backend()):
        if mode == "thermal":
            j = helpers.stt_layer_mtj(cmtj_b200, temperature=300.0)
        else:
            j = helpers.two_layer_mtj(cmtj_b200)
        sw = cmtj_b200.ShardedSweep(j, B)
        assert (sw.rank, sw.world) == (rank, world)
        sw.set_external_field("all", H)
        sw.set_magnetisation("free", m0)
        log = sw.run(0.1e-9, 1e-13, 1e-12, solverMode=cmtj_b200.SolverMode.Heun if mode == "thermal" else cmtj_b200.SolverMode.RK4, seed=5)
        mf = sw.gather_magnetisation("free")
    if rank == 0:
        np.savez(out_path, m_final=mf, **{k: v for k, v in log.items()})
    else:
        assert log is None and mf is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
