"""Reference trajectories in bulk: the UNMODIFIED reference extension (oracle/_ref) driven from a process
pool so that the stochastic tests can afford >= 1000 reference runs per assertion (the reference is not
thread-safe per object but every worker builds its own junctions; its RNG is mt19937(random_device) per
object, so workers never share a stream).

Workers are spawned, not forked: the parent of a GPU test holds a CUDA context.  Everything a worker
needs is importable from tests/ (this module and helpers.py); it never imports torch or the CUDA library.
"""
import math
import multiprocessing as mp
import os

import numpy as np

_POOL = None


def n_workers():
    return max(1, min(os.cpu_count() or 1, 32))


def pool():
    global _POOL
    if _POOL is None:
        _POOL = mp.get_context("spawn").Pool(n_workers())
    return _POOL


def close():
    global _POOL
    if _POOL is not None:
        _POOL.terminate()
        _POOL = None


def run(fn, n_total, *args):
    """fn(n, *args) -> array with leading axis n; split n_total over the workers and concatenate."""
    w = n_workers()
    per = [n_total // (4 * w)] * (4 * w)
    for k in range(n_total - sum(per)):
        per[k] += 1
    parts = pool().starmap(fn, [(n, *args) for n in per if n > 0])
    return np.concatenate(parts, axis=0)


def _ref():
    import helpers

    return helpers.reference_module()


# ---- model builders shared with tests/test_stochastic_gpu.py (mod = reference module or cmtj_b200) ------
def pbit(mod, Hx=0.0, T=300.0, mag=(1, 0, 0.02), current=-1e10):
    """p-bit / thermal-switching junction of curated-examples/p-bit.py:188-227 (BASELINE config 3)
    with a 7 nm radius free layer so that 300 K matters within 0.5 ns."""
    demag = [(0.00024164288391924, 0, 0), (0, 0.000160046006320031, 0), (0, 0, 0.999598310229469)]
    l = mod.Layer.createSTTLayer("free", mod.CVector(*mag), mod.CVector(1, 0, 0), 1.6, 2e-9, math.pi * (0.7e-9) ** 2 * 1e2,
                                 [mod.CVector(*r) for r in demag], 0.1, 1.0, 0.9, 1.0)
    l.setReferenceLayer(mod.CVector(1, 0, 0))
    j = mod.Junction([l], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(350e3))
    j.setLayerCurrentDriver("free", mod.ScalarDriver.getConstantDriver(current))
    j.setLayerExternalFieldDriver("free", mod.AxialDriver(Hx, 0.0, 0.0))
    if T is not None:
        j.setLayerTemperatureDriver("free", mod.ScalarDriver.getConstantDriver(T))
    return j


def fmr(mod):
    """Thermally excited precession around +x in a 200 kA/m field (no current)."""
    return pbit(mod, Hx=200e3, mag=(1, 0, 0.0), current=0.0)


def onef_junction(mod, alpha, std, scale, axis_name="all", mag=(1, 0, 0)):
    """single_layer_mtj_fictious of the reference's conftest, released along its easy axis (a stable
    equilibrium), with 1/f^alpha noise on the free layer."""
    import helpers

    j = helpers.single_layer_mtj(mod, mag=mag)
    j.getLayer("free").setAlphaNoise(alpha, std, scale, getattr(mod.Axis, axis_name))
    return j


# ---- workers (module level: picklable) ---------------------------------------------------------------------
def switching_trials(n, Hx, steps, mode):
    """final m_x of n reference trials at field Hx"""
    ref = _ref()
    out = np.zeros(n)
    for k in range(n):
        j = pbit(ref, Hx)
        j.runSimulation(steps * 1e-13, 1e-13, steps * 1e-13, solverMode=getattr(ref.SolverMode, mode))
        out[k] = j.getLayerMagnetisation("free").x
    return out


def diffusion_trials(n, steps):
    ref = _ref()
    out = np.zeros((n, 3))
    for k in range(n):
        j = pbit(ref, mag=(0, 0, 1))
        j.runSimulation(steps * 1e-13, 1e-13, steps * 1e-13)
        out[k] = j.getLayerMagnetisation("free").tolist()
    return out


def fmr_spectra(n, steps, n_bins):
    """per trajectory: Hann-windowed periodogram of m_y(t) in bins [0, n_bins) and Var(m_y); [n, n_bins + 1]"""
    ref = _ref()
    win = np.hanning(steps)
    out = np.zeros((n, n_bins + 1))
    for k in range(n):
        j = fmr(ref)
        j.runSimulation(steps * 1e-13, 1e-13, 1e-13, solverMode=ref.SolverMode.Heun)
        y = np.asarray(j.getLog()["free_my"], dtype=np.float64)[:steps]
        out[k, :n_bins] = np.abs(np.fft.rfft((y - y.mean()) * win)[:n_bins]) ** 2
        out[k, n_bins] = y.var()
    return out


def alpha_buffers(n_series, n, alpha, std):
    """n_series BufferedAlphaNoise buffers of n ticks each (scale 1); [n_series, n]"""
    ref = _ref()
    out = np.zeros((n_series, n))
    for k in range(n_series):
        g = ref.noise.BufferedAlphaNoise(n, alpha, std, 1.0)
        out[k] = [g.tick() for _ in range(n)]
    return out


def onef_trajectories_z(n, steps, alpha, std, scale, axis_name, every):
    """m_z(t) of n reference runs under 1/f^alpha noise only, released along +x; [n, steps // every]"""
    ref = _ref()
    out = np.zeros((n, steps // every))
    for k in range(n):
        j = onef_junction(ref, alpha, std, scale, axis_name)
        j.runSimulation(steps * 1e-13, 1e-13, every * 1e-13 * (1 + 1e-9), solverMode=ref.SolverMode.EulerHeun)
        out[k] = np.asarray(j.getLog()["free_mz"], dtype=np.float64)[: steps // every]
    return out
