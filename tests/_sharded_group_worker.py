"""Worker of tests/test_sharded_group.py: one rank of a world_size-2 gloo job on CPU.  The oracle's host-memory
stand-in plays the device library; what is under test is GroupInteraction.run_sharded's partition, its default
torch.distributed exchange callback (one in-place all-gather per step) and the final gather of the logs."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmtj_b200  # noqa: E402, F401
import helpers  # noqa: E402
from test_sharded_group import lattice_group, logs_of  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    with helpers.facade_on(helpers.oracle_backend()):
        g = lattice_group(9)  # 5 + 4
        res = g.run_sharded(0.05e-9, 1e-13, 1e-12, _host_memory=True)
        assert (res["lo"], res["hi"]) == ((0, 5) if rank == 0 else (5, 9))
    flat = {f"r{rank}/{k}/{name}": v for k, log in enumerate(logs_of(g)) for name, v in log.items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, flat)
    if rank == 0:
        merged = {}
        for d in gathered:
            merged.update(d)
        np.savez(out_path, **merged)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
