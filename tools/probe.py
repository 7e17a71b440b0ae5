"""Quick GPU probe: FP64 peak + a first throughput number for the S1 sweep shape."""
import ctypes as C, time, sys, json
import numpy as np
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmtj_b200, helpers
from cmtj_b200 import _abi
import torch

lib = _abi.load_library()
fl, ms = C.c_double(), C.c_double()
for it in (20000, 200000):
    lib.cmtjb_fp64_peak(0, C.c_int64(it), C.byref(fl), C.byref(ms))
    print(f"fp64 peak probe iters={it}: {fl.value/1e12:.2f} TFLOP/s in {ms.value:.2f} ms", flush=True)

def s1_batch(B, dyn=True):
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
    return b

for dyn in (True, False):
    for B in (148 * 128 * 4, 100000, 148 * 128 * 8):
        for steps, wf in ((20000, 1e-11), (20000, 1e-13)):
            b = s1_batch(B, dyn)
            T = steps * 1e-13 * (1 + 1e-9)
            r = b.run_device(T, 1e-13, wf, log=(wf > 5e-13 or B <= 100000)); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = b.run_device(T, 1e-13, wf, log=(wf > 5e-13 or B <= 100000)); e1.record(); torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
            print(f"dyn={dyn} B={B} steps={steps} wf={wf:g} variant={b.last_plan.variant.decode()} "
                  f"{dt*1e3:.1f} ms  {2*B*steps/dt:.3e} layer-steps/s", flush=True)
            del r
