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
def T_of(k):
    return k * 1e-13 * (1 + 1e-9)


# per-step DEVICE time comes from the library (CMTJB_TRACE_SHARD: CUDA events around the step loop, printed per rank);
# host wall time of a call is dominated by packing and uploading n junction models
os.environ["CMTJB_TRACE_SHARD"] = "1"
T = T_of(steps)
lattice_group(n).run_sharded(T_of(10), 1e-13, 1e-12, seed=7, gather=False, transport="nccl")  # NCCL's lazy channel set-up
out = {"what": "GroupInteraction.run_sharded: one dipole array over the GPUs of a box, top-layer m exchanged every step", "elements": n,
       "steps": steps, "world": world}
logs = {}
for transport in ("nccl", "peer"):
    g = lattice_group(n)
    dist.barrier()
    g.run_sharded(T, 1e-13, 1e-12, seed=7, gather=True, transport=transport)
    logs[transport] = logs_of(g)
if rank == 0:
    a0, b0 = logs["nccl"], logs["peer"]
    out["peer_bit_identical_to_nccl"] = bool(all(np.array_equal(a0[k][name], b0[k][name]) for k in range(len(b0)) for name in b0[k]))
    one = lattice_group(n)
    one.run_sharded(T, 1e-13, 1e-12, rank=0, world=1, seed=7, transport="nccl")
    b = logs_of(one)
    out["bit_identical_to_world_1"] = bool(all(np.array_equal(b0[k][name], b[k][name]) for k in range(n) for name in b[k]))
    coop = lattice_group(n)
    torch.cuda.synchronize(); t0 = time.perf_counter(); coop.runSimulation(T_of(10), 1e-13, 1e-12); torch.cuda.synchronize(); ts = time.perf_counter() - t0
    coop = lattice_group(n)
    torch.cuda.synchronize(); t0 = time.perf_counter(); coop.runSimulation(T, 1e-13, 1e-12); torch.cuda.synchronize(); tl = time.perf_counter() - t0
    out["us_per_step_cooperative_kernel_one_gpu_wall_difference"] = 1e6 * (tl - ts) / (steps - 10)
    c = logs_of(coop)
    scale = max(np.max(np.abs(c[k][nm])) for k in range(n) for nm in c[k] if nm != "time")
    out["max_rel_diff_vs_cooperative_kernel"] = float(max(np.max(np.abs(b0[k][nm] - c[k][nm])) for k in range(n) for nm in c[k] if nm != "time") / scale)
    print(json.dumps(out), flush=True)
    assert out["peer_bit_identical_to_nccl"] and out["bit_identical_to_world_1"]
dist.barrier()
dist.destroy_process_group()
