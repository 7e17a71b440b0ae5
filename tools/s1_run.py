"""Run the S1 sweep shape (2-layer MTJ, RK4, per-instance Hext) once or a few times on the
device path; used under ncu.  usage: s1_run.py B steps write_every [reps] [dyn]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import cmtj_b200, helpers

B, steps, we = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dyn = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
j = helpers.two_layer_mtj(cmtj_b200, mag=(0, 0, 1))
if dyn:
    oe = cmtj_b200.ScalarDriver.getStepDriver(0, 50, 0, 3e-13)
    j.setLayerOerstedFieldDriver("all", cmtj_b200.AxialDriver(cmtj_b200.NullDriver(), cmtj_b200.NullDriver(), oe))
rng = np.random.default_rng(0)
b = cmtj_b200.JunctionBatch(j, B)
H = np.zeros((B, 3)); H[:, 0] = np.linspace(-500e3, 500e3, B)
b.set_external_field("all", H)
b.set_magnetisation("free", rng.normal(0, 1e-3, (B, 3)) + [0, 0, 1.0])
b.set_magnetisation("bottom", rng.normal(0, 1e-3, (B, 3)) + [0, 0, 1.0])
T = steps * 1e-13 * (1 + 1e-9)
out = None
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = b.run_device(T, 1e-13, (we * 1e-13 * (1 + 1e-9)) if we > 0 else T, log=we > 0, out=out)
    e1.record(); torch.cuda.synchronize()
    out = res.get("_log")
    dt = e0.elapsed_time(e1) * 1e-3
    nbytes = 0 if out is None else out.numel() * 8
    print(f"rep {r} variant={b.last_plan.variant.decode()} B={B} steps={steps} we={we}: {dt*1e3:.2f} ms "
          f"{2*B*steps/dt:.3e} layer-steps/s log {nbytes/1e9:.2f} GB -> {nbytes/dt/1e9:.0f} GB/s", flush=True)
