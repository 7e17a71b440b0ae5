"""One VSD_procedure call (used under ncu). usage: vsd_run.py nH nF steps"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, time
import cmtj_b200 as m, helpers
P = m.utils.procedures
nH, nF, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rp = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
      P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5)]
Hv = np.zeros((nH, 3)); Hv[:, 0] = np.linspace(50e3, 400e3, nH); Hv[:, 2] = 20e3
fs = np.linspace(2e9, 20e9, nF)
for r in range(2):
    t0 = time.perf_counter()
    vm = P.VSD_procedure(helpers.two_layer_mtj(m), Hv, fs, 1e-13, rp, simulation_duration=steps * 1e-13 * (1 + 1e-9), Rtype="Rz", disable_tqdm=True, carry_state=False)
    torch.cuda.synchronize()
    print(f"VSD {nH}x{nF} x {steps} steps: {time.perf_counter()-t0:.3f} s  {2*nH*nF*steps/(time.perf_counter()-t0):.3e} layer-steps/s", flush=True)
