"""A dipole-coupled array split over the GPUs of one box with the NCCL all-gather exchange
(GroupInteraction.run_sharded -> cmtjb_group_run_sharded).  Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/group_sharded_run.py [n_elements] [steps]
Rank 0 also integrates the whole array alone (world = 1) and checks that the gathered logs are bit-identical, then
prints one JSON line with the per-step time of both."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, torch.distributed as dist
import cmtj_b200  # noqa: F401
from test_sharded_group import lattice_group, logs_of

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
T = steps * 1e-13 * (1 + 1e-9)
out = {}
for rep in range(2):  # the first pass pays NCCL's lazy channel set-up
    g = lattice_group(n)
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    g.run_sharded(T, 1e-13, 1e-12, seed=7, gather=True)
    torch.cuda.synchronize(); dist.barrier(); out["sharded_s"] = time.perf_counter() - t0
if rank == 0:
    for rep in range(2):
        one = lattice_group(n)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        one.run_sharded(T, 1e-13, 1e-12, rank=0, world=1, seed=7)
        torch.cuda.synchronize(); out["single_s"] = time.perf_counter() - t0
    a, b = logs_of(g), logs_of(one)
    same = all(np.array_equal(a[k][name], b[k][name]) for k in range(n) for name in b[k])
    coop = lattice_group(n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    coop.runSimulation(T, 1e-13, 1e-12)
    torch.cuda.synchronize(); out["cooperative_single_gpu_s"] = time.perf_counter() - t0
    c = logs_of(coop)
    scale = max(np.max(np.abs(c[k][nm])) for k in range(n) for nm in c[k] if nm != "time")
    err = max(np.max(np.abs(a[k][nm] - c[k][nm])) for k in range(n) for nm in c[k] if nm != "time") / scale
    print(json.dumps({"what": "GroupInteraction.run_sharded, NCCL all-gather of the top-layer m per step", "elements": n, "steps": steps,
                      "world": world, "bit_identical_to_world_1": bool(same), "max_rel_diff_vs_cooperative_kernel": float(err),
                      "us_per_step_sharded": 1e6 * out["sharded_s"] / steps, "us_per_step_world_1": 1e6 * out["single_s"] / steps,
                      "us_per_step_cooperative_kernel_incl_host": 1e6 * out["cooperative_single_gpu_s"] / steps}), flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
