"""Time cmtjb_rfft_abs_device on a synthetic time-major log.  usage: spec_run.py B n fmax [modes...]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from cmtj_b200.utils import procedures as P

B, n, fmax = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
modes = sys.argv[4:] or [""]
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty((n, B), dtype=torch.float64, device="cuda")
step = max(1, n // 16)
for lo in range(0, n, step):  # filled in slices: no second buffer of the log's size
    x[lo:lo + step].normal_(generator=g)
ref = None
for mode in modes:
    os.environ.pop("CMTJB_SPECTRUM", None); os.environ.pop("CMTJB_SPECTRUM_N2", None)
    if mode:
        os.environ["CMTJB_SPECTRUM"] = mode.split(":")[0]
        if ":" in mode: os.environ["CMTJB_SPECTRUM_N2"] = mode.split(":")[1]
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y, fr = P.spectrum_device(x, 1e-13, fmax); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    err = ""
    if ref is None: ref = y
    else: err = " max |diff| / max vs first mode %.2e" % float((y - ref).abs().max() / ref.abs().max())
    print("mode %-12s B %d n %d K %d: %.2f ms (%.1f GB/s of log)%s" % (mode or "default", B, n, y.shape[1], ms, n * B * 8 / ms / 1e6, err), flush=True)
