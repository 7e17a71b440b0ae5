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


def timed(fn):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize(); dist.barrier()
    return time.perf_counter() - t0


T = T_of(steps)
short = max(10, steps // 5)
g = lattice_group(n)
g.run_sharded(T_of(10), 1e-13, 1e-12, seed=7, gather=False, transport="nccl")  # NCCL's lazy channel set-up
# marginal time per step: (long run - short run) / (steps - short); the fixed part is packing and uploading n junction models
out = {"what": "GroupInteraction.run_sharded: one dipole array over the GPUs of a box, top-layer m exchanged every step", "elements": n,
       "steps": steps, "world": world}
for transport in ("nccl", "peer"):
    t_short = timed(lambda: lattice_group(n).run_sharded(T_of(short), 1e-13, 1e-12, seed=7, gather=False, transport=transport))
    g = lattice_group(n)
    t_long = timed(lambda: g.run_sharded(T, 1e-13, 1e-12, seed=7, gather=True, transport=transport))
    out["us_per_step_" + transport] = 1e6 * (t_long - t_short) / (steps - short)
    if transport == "nccl":
        g_nccl = logs_of(g)
if rank == 0:
    a0, b0 = g_nccl, logs_of(g)
    out["peer_bit_identical_to_nccl"] = bool(all(np.array_equal(a0[k][name], b0[k][name]) for k in range(len(b0)) for name in b0[k]))
if rank == 0:
    one = lattice_group(n)
    t1s = time.perf_counter(); lattice_group(n).run_sharded(T_of(short), 1e-13, 1e-12, rank=0, world=1, seed=7); torch.cuda.synchronize(); t1s = time.perf_counter() - t1s
    t1l = time.perf_counter(); one.run_sharded(T, 1e-13, 1e-12, rank=0, world=1, seed=7); torch.cuda.synchronize(); t1l = time.perf_counter() - t1l
    out["us_per_step_world_1"] = 1e6 * (t1l - t1s) / (steps - short)
    a, b = logs_of(g), logs_of(one)
    same = all(np.array_equal(a[k][name], b[k][name]) for k in range(n) for name in b[k])
    coop = lattice_group(n)
    tcs = time.perf_counter(); lattice_group(n).runSimulation(T_of(short), 1e-13, 1e-12); torch.cuda.synchronize(); tcs = time.perf_counter() - tcs
    tcl = time.perf_counter(); coop.runSimulation(T, 1e-13, 1e-12); torch.cuda.synchronize(); tcl = time.perf_counter() - tcl
    out["us_per_step_cooperative_kernel_one_gpu"] = 1e6 * (tcl - tcs) / (steps - short)
    c = logs_of(coop)
    scale = max(np.max(np.abs(c[k][nm])) for k in range(n) for nm in c[k] if nm != "time")
    err = max(np.max(np.abs(a[k][nm] - c[k][nm])) for k in range(n) for nm in c[k] if nm != "time") / scale
    out["bit_identical_to_world_1"] = bool(same)
    out["max_rel_diff_vs_cooperative_kernel"] = float(err)
    print(json.dumps(out), flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
