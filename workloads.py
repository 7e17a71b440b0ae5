"""The BASELINE.json configurations as builders, shared by bench.py (headline + `secondary` block), the
full-size parity tests (tests/test_gpu_fullsize.py) and tools/bench_configs.py, so that what is measured and
what is parity-tested is the same model at the same size.  `mod` is the module that provides the cmtj API
(cmtj_b200, or the reference extension for the CPU arm).

  S1  PIMM spectrum      2-layer MTJ, 1e5 field points x 2e5 RK4 steps, dt = 1e-13, log every 100 (bench.py headline)
  S2  VSD field scan     1-layer, sine Oersted drive, log EVERY step (curated-examples/vsd-basic-field-scan.py)
  S3  p-bit              1-layer STT at 300 K, 1e6 trials x 5000 Heun / Euler-Heun steps (curated-examples/p-bit.py)
  S4  CIMS diagram       3-layer stack, 2048 x 2048 current x field grid, 20 000 Heun steps, SOT step drivers
  S5g reservoir          64 x 64 dipole-coupled macrospins, RK4 (core/reservoir.hpp)
  S5s series stack       64 STT MTJs in series (KCL), a batch of stacks (core/stack.hpp)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = 1e-13


def exact(steps, dt=DT):
    """a duration whose (unsigned)(T / dt) is exactly `steps` (SURVEY 8a note 2)"""
    return steps * dt * (1 + 1e-9)


# ---- S2 ---------------------------------------------------------------------------------------------------------
S2_INSTANCES, S2_STEPS = 262144, 5000


def s2_junction(mod):
    demag = [mod.CVector(0, 0, 0), mod.CVector(0, 0, 0), mod.CVector(0, 0, 1)]
    l = mod.Layer("free", mod.CVector(0, 0, 1), mod.CVector(0, 0, 1), 0.525, 1.45e-9, 0.0, demag, 0.03)
    l.setReferenceLayer(mod.CVector(0, 1, 0))
    j = mod.Junction([l], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(1.54e5))
    j.setLayerOerstedFieldDriver("free", mod.AxialDriver(mod.NullDriver(), mod.NullDriver(), mod.ScalarDriver.getSineDriver(0, -5000, 14e9, 0)))
    return j


def s2_fields(B):
    th = math.radians(91)
    Hm = np.linspace(350e3, 630e3, B)
    return np.stack([Hm * math.sin(th), np.zeros(B), Hm * math.cos(th)], axis=1)


# ---- S3 ---------------------------------------------------------------------------------------------------------
S3_INSTANCES, S3_STEPS = 1_000_000, 5000


def s3_junction(mod):
    import helpers

    return helpers.stt_layer_mtj(mod, temperature=300.0)


def s3_fields(B):
    H = np.zeros((B, 3))
    H[:, 0] = np.repeat([-500e3, -50e3, 0, 50e3, 500e3], B // 5 + 1)[:B]
    return H


# ---- S4 ---------------------------------------------------------------------------------------------------------
S4_GRID, S4_STEPS, S4_DT = 2048, 20000, 1e-12


def s4_junction(mod):
    import helpers

    j = helpers.tri_layer_mtj(mod)
    j.setLayerFieldLikeTorqueDriver("free", mod.ScalarDriver.getStepDriver(0, 1e3, 0, 10e-9))
    j.setLayerDampingLikeTorqueDriver("free", mod.ScalarDriver.getStepDriver(0, 16e3, 0, 10e-9))
    j.getLayer("free").setReferenceLayer(mod.CVector(0, 1, 0))
    return j


def s4_grid(n):
    """per-instance (field-like amplitude, damping-like amplitude, Hext) of the n x n current x field grid"""
    cur, fld = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-400e3, 400e3, n))
    B = n * n
    return 1e3 * cur.ravel() ** 2, 16e3 * cur.ravel(), np.stack([fld.ravel(), np.zeros(B), np.zeros(B)], axis=1)


def s4_batch(cm, n, select=None, device=0):
    """JunctionBatch of the grid (cm = cmtj_b200), or of the instances `select` of it"""
    fl, dl, H = s4_grid(n)
    if select is not None:
        fl, dl, H = fl[select], dl[select], H[select]
    b = cm.JunctionBatch(s4_junction(cm), len(fl), device=device)
    b.set_override("free", "fieldLike", "amplitude", fl)
    b.set_override("free", "dampingLike", "amplitude", dl)
    b.set_external_field("all", H)
    return b


# ---- S5 ---------------------------------------------------------------------------------------------------------
S5G_SIDE, S5G_STEPS = 64, 20000
S5S_STACKS, S5S_JUNCTIONS, S5S_STEPS = 4096, 64, 2000


def s5_lattice(mod, n_side, pitch=100e-9, seed=0):
    """n_side x n_side macrospins on a square lattice, dipole-coupled (GroupInteraction, reservoir.hpp:129-226)"""
    js, pos = [], []
    rng = np.random.default_rng(seed)
    for k in range(n_side * n_side):
        mag = rng.normal(0, 0.2, 3) + np.array([0.0, 0.0, 1.0])
        l = mod.Layer("free", mod.CVector(*mag), mod.CVector(0, 0, 1), 1.2, 2e-9, math.pi * (20e-9) ** 2, [mod.CVector(0, 0, 0)] * 3, 0.01)
        j = mod.Junction([l])
        j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(0.3e6))
        js.append(j)
        pos.append(mod.CVector(pitch * (k % n_side), pitch * (k // n_side), 0.0))
    return mod.reservoir.GroupInteraction(pos, js, "free")


def s5_stack(mod, n_junctions=S5S_JUNCTIONS):
    """series stack of STT MTJs (p-bit junction parameters), coupling 0.1, sine current (stack.hpp:317-423)"""
    import helpers

    js = [helpers.stt_layer_mtj(mod) for _ in range(n_junctions)]
    s = mod.stack.SeriesStack(js)
    s.setCouplingStrength(0.1)
    s.setCoupledCurrentDriver(mod.ScalarDriver.getSineDriver(1e10, 5e9, 6e9, 0.0))
    return s


def s5_stack_amplitudes(n_stacks):
    return np.linspace(1e9, 9e9, n_stacks)
