"""One PIMM_batch call at the BASELINE configs[1] shape (2-layer MTJ, nH field points x steps RK4 steps, spectra on
the host), wall-timed; CMTJB_TRACE_PROC=1 prints the phases.  usage: pimm_run.py nH steps [reps]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import cmtj_b200 as m, helpers
P = m.utils.procedures
rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
      P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
nH, steps = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H = np.zeros((nH, 3)); H[:, 0] = np.linspace(-500e3, 500e3, nH)
import gc
for r in range(reps):
    if r == 2 and os.environ.get("PIMM_RUN_GC_FREEZE"):
        gc.collect(); gc.freeze()  # what is alive now is never scanned again: full collections stay cheap
    g0 = sum(s["collections"] for s in gc.get_stats()[2:])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    spec, freqs, out = P.PIMM_batch(helpers.two_layer_mtj(m), H, 1e-13, rp, simulation_duration=steps * 1e-13 * (1 + 1e-9), disturbance=1e-3, disable_tqdm=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    g1 = sum(s["collections"] for s in gc.get_stats()[2:])
    print(f"rep {r}: PIMM_batch {nH} x {steps}: {dt*1e3:.1f} ms = {2*nH*steps/dt:.3e} layer-steps/s; spectrum {np.asarray(spec).shape}; full GCs {g1 - g0}", flush=True)
