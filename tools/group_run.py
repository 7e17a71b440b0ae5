"""One device-resident GroupInteraction run (S5g: side x side dipole lattice) per repetition, event-timed; used under ncu.
usage: group_run.py side steps [reps]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import cmtj_b200 as m, workloads as wl
from cmtj_b200 import _abi, _coupled
side, steps = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _abi.load_library()
g = wl.s5_lattice(m, side)
N = side * side
gs, jns, layers, nl, xyz = g._pack()
a = _coupled.run_args(wl.exact(steps), 1e-13, wl.exact(steps), 1, log=False, seed=1)
mstate = torch.from_numpy(_coupled.initial_state(g.junctionList, 1)).cuda()
status = torch.zeros(N, dtype=torch.int32, device="cuda")
f = lib.cmtjb_group_run_device
f.restype = C.c_int
f.argtypes = [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = f(C.byref(gs), C.cast(jns, C.c_void_p), C.cast(layers, C.c_void_p), None, 0, C.byref(a), C.c_void_p(mstate.data_ptr()), None,
           C.c_void_p(status.data_ptr()), 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _abi.cuda_backend().last_error()
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {N} elements x {steps} steps: {ms:.2f} ms between events ({ms * 1e3 / steps:.2f} us/step), host side of the call {t_host * 1e3:.2f} ms "
          f"-> kernel ~{(ms - t_host * 1e3) * 1e3 / steps:.2f} us/step, {N * (N - 1) * steps / (ms * 1e-3):.3e} pairs/s", flush=True)
