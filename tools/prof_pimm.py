"""cProfile of one PIMM_procedure / VSD_procedure call (host-side time; device waits show up in .cpu())."""
import cProfile, pstats, sys, os, math, io
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import cmtj_b200 as m, helpers
P = m.utils.procedures
rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
      P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
nH, steps = 100000, 50000
H = np.zeros((nH, 3)); H[:, 0] = np.linspace(-500e3, 500e3, nH)
def pimm():
    return P.PIMM_procedure(helpers.two_layer_mtj(m), H, 1e-13, rp, simulation_duration=steps * 1e-13 * (1 + 1e-9), disturbance=1e-3, disable_tqdm=True, carry_state=False)
Hv = np.zeros((256, 3)); Hv[:, 0] = np.linspace(50e3, 400e3, 256); Hv[:, 2] = 20e3
fs = np.linspace(2e9, 20e9, 64)
def vsd():
    return P.VSD_procedure(helpers.two_layer_mtj(m), Hv, fs, 1e-13, rp, simulation_duration=1e5 * 1e-13 * (1 + 1e-9), Rtype="Rz", disable_tqdm=True, carry_state=False)
for name, fn in (("pimm", pimm), ("vsd", vsd)):
    fn(); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable(); fn(); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
    print("=====", name); print("\n".join(l[:150] for l in s.getvalue().splitlines()[:40]))
