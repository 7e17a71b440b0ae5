"""Per-chunk timeline of cmtjb_run_host on the bench workload (CMTJB_TRACE_HOST=1). usage: e2e_trace.py [chunks]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
if len(sys.argv) > 1:
    os.environ["CMTJB_HOST_CHUNKS"] = sys.argv[1]
import numpy as np, torch
import cmtj_b200, helpers
B, T, DT = 100000, 20e-9, 1e-13
j = helpers.two_layer_mtj(cmtj_b200, mag=(0, 0, 1))
oe = cmtj_b200.ScalarDriver.getStepDriver(0, 50, 0, 3e-13)
j.setLayerOerstedFieldDriver("all", cmtj_b200.AxialDriver(cmtj_b200.NullDriver(), cmtj_b200.NullDriver(), oe))
b = cmtj_b200.JunctionBatch(j, B)
H = np.zeros((B, 3)); H[:, 0] = np.linspace(-500e3, 500e3, B)
b.set_external_field("all", H)
host = torch.empty((7, 2000, B), dtype=torch.float64, pin_memory=True); host.zero_()
m0 = b.m_state.copy()
for rep in range(3):
    b.m_state[...] = m0
    if rep == 2:
        os.environ["CMTJB_TRACE_HOST"] = "1"
    t0 = time.perf_counter()
    b.run(T, DT, 100 * DT * (1 + 1e-9), out=host.numpy())
    print(f"rep {rep}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
