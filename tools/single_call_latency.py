"""Latency of ONE Junction.runSimulation call (B = 1), the way the reference's curated examples drive the library
(a Python loop of runSimulation + getLog per sweep point), against the unmodified reference on one host core."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import cmtj_b200, helpers

ref = helpers.reference_module()
for steps, every in ((1000, 1), (100000, 100), (400000, 1)):
    row = {}
    for name, mod in (("cuda", cmtj_b200), ("reference_cpu", ref)):
        if mod is None:
            continue
        j = helpers.two_layer_mtj(mod)
        ts = []
        for rep in range(4):
            j.clearLog()
            t0 = time.perf_counter()
            j.runSimulation(steps * 1e-13 * (1 + 1e-9), 1e-13, every * 1e-13)
            t1 = time.perf_counter()
            log = j.getLog()
            t2 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1))
        row[name] = (min(t[0] for t in ts[1:]), min(t[1] for t in ts[1:]))
    print(f"{steps} steps, log every {every}: " + "; ".join(f"{k}: run {1e3*a:.2f} ms, getLog {1e3*b:.2f} ms" for k, (a, b) in row.items()), flush=True)
