"""Shared test helpers: oracle loader, reference-module loader, and model builders that
construct the SAME system against either API (the reference `_cmtj` or `cmtj_b200`)."""
import contextlib
import ctypes
import glob
import importlib.util
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_oracle = None
_ref = False


def oracle_backend():
    global _oracle
    if _oracle is None:
        from cmtj_b200 import _abi

        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "llgs_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        lib = ctypes.CDLL(so)

        class OracleBackend(_abi.Backend):
            """The CPU oracle behind the facade's backend interface.  The product passes no 1/f^alpha buffer (its
            library generates one on the device); the oracle takes the noise as an input, so this adapter draws it
            from the numpy restatement of the reference's generator (oracle/noise_oracle.py)."""

            def run_host(self, jn, layers, overrides, n_overrides, args, m_state, log, status, n_samples_out, time_log, device=0):
                keep = None
                if not args.onef_buffer and any(layers[i].noiseScale != 0 for i in range(jn.n_layers)):
                    import sys

                    if ORACLE_DIR not in sys.path:
                        sys.path.insert(0, ORACLE_DIR)
                    import noise_oracle

                    plan = self.plan(jn, layers, args)
                    keep = noise_oracle.junction_buffer(layers, jn.n_layers, int(plan.total_iterations), int(args.n_instances))
                    args.onef_buffer = keep.ctypes.data
                try:
                    return super().run_host(jn, layers, overrides, n_overrides, args, m_state, log, status, n_samples_out, time_log, device)
                finally:
                    if keep is not None:
                        args.onef_buffer = None

        _oracle = OracleBackend(lib, "oracle_", has_device_arg=False)
    return _oracle


@contextlib.contextmanager
def facade_on(backend):
    """Run the cmtj_b200 facade against `backend` for the duration of the block: None = the
    product's own CUDA library (no patching), otherwise e.g. oracle_backend().  Test
    infrastructure only: this monkeypatches cmtj_b200._abi.backend; the product has no switch."""
    from cmtj_b200 import _abi

    if backend is None:
        yield _abi.cuda_backend()
        return
    prev = _abi.backend
    _abi.backend = lambda: backend
    try:
        yield backend
    finally:
        _abi.backend = prev


def reference_module():
    """Import oracle/_ref/_cmtj*.so (None when it has not been built)."""
    global _ref
    if _ref is False:
        cands = glob.glob(os.path.join(ORACLE_DIR, "_ref", "_cmtj*.so"))
        if not cands:
            _ref = None
        else:
            spec = importlib.util.spec_from_file_location("_cmtj", cands[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref = mod
    return _ref


# ---------------------------------------------------------------------------------------
# model builders (mod is either the reference module or cmtj_b200)
# ---------------------------------------------------------------------------------------
DEMAG = [(0, 0, 0), (0, 0, 0), (0, 0, 1)]
ALPHA = 0.005
SURFACE = 150e-9 * 150e-9 * math.pi


def _demag(mod, d=DEMAG):
    return [mod.CVector(*r) for r in d]


def two_layer_mtj(mod, mag=(0.1, 0.1, 1.0)):
    """tests/conftest.py:104-140 of the reference (fixture two_layer_mtj)."""
    Kdir = mod.CVector(1, 0, 0)
    l1 = mod.Layer("free", mod.CVector(*mag), Kdir, 1.65, 3e-9, SURFACE, _demag(mod), ALPHA)
    l2 = mod.Layer("bottom", mod.CVector(*mag), Kdir, 1.45, 4e-9, SURFACE, _demag(mod), ALPHA)
    j = mod.Junction([l1, l2], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(1.05e3))
    j.setLayerAnisotropyDriver("bottom", mod.ScalarDriver.getConstantDriver(50e3))
    j.setIECDriver("free", "bottom", mod.ScalarDriver.getConstantDriver(-1.78e-3))
    j.setQuadIECDriver("free", "bottom", mod.ScalarDriver.getConstantDriver(-1.69e-4))
    return j


def single_layer_mtj(mod, mag=(0.1, 0.1, 1.0)):
    """tests/conftest.py:81-101 (single_layer_mtj_fictious)."""
    l1 = mod.Layer("free", mod.CVector(*mag), mod.CVector(1, 0, 0), 1.65, 3e-9, SURFACE, _demag(mod), ALPHA)
    l1.setReferenceLayer(mod.CVector(1, 0, 0))
    j = mod.Junction([l1], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(1.05e3))
    return j


def tri_layer_mtj(mod, mag=(0.1, 0.1, 1.0)):
    """tests/conftest.py:184-237 (tri_layer_mtj)."""
    Kdir = mod.CVector(1, 0, 0)
    l1 = mod.Layer("free", mod.CVector(*mag), Kdir, 1.65, 3e-9, SURFACE, _demag(mod), ALPHA)
    l2 = mod.Layer("middle", mod.CVector(*mag), Kdir, 1.45, 4e-9, SURFACE, _demag(mod), ALPHA)
    l3 = mod.Layer("bottom", mod.CVector(*mag), mod.CVector(0, 0, 1.0), 0.7, 1e-9, SURFACE, _demag(mod), ALPHA)
    j = mod.Junction([l1, l2, l3])
    for l, k in zip(["free", "middle", "bottom"], [1.05e3, 50e3, 1e4]):
        j.setLayerAnisotropyDriver(l, mod.ScalarDriver.getConstantDriver(k))
    j.setIECDriver("free", "middle", mod.ScalarDriver.getConstantDriver(-1.78e-3))
    j.setQuadIECDriver("free", "middle", mod.ScalarDriver.getConstantDriver(-1.69e-4))
    j.setIECDriver("middle", "bottom", mod.ScalarDriver.getConstantDriver(3e-4))
    j.setQuadIECDriver("middle", "bottom", mod.ScalarDriver.getConstantDriver(2e-6))
    return j


def multi_layer(mod, n, strip=False, mag=(0.1, 0.1, 1.0)):
    """An n-layer stack (n up to 8): alternating materials and easy axes, IEC + biquadratic coupling between
    every pair of neighbours, one tilted-demag layer, a sine Oersted drive and a static SOT layer at the bottom
    (exercises field terms, neighbour hand-off and torque terms for every layer position)."""
    Ms = [1.65, 1.45, 0.7, 1.2, 1.0, 0.9, 1.3, 1.1]
    th = [3e-9, 4e-9, 1e-9, 2e-9, 1.5e-9, 2.5e-9, 1e-9, 3e-9]
    K = [1.05e3, 50e3, 1e4, 3e4, 2e4, 0.8e4, 4e4, 1e3]
    axes = [(1, 0, 0), (1, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]
    layers = []
    for i in range(n):
        d = DEMAG if i != 2 else [(0.05, 0.01, 0), (0.01, 0.1, 0), (0, 0, 0.85)]
        v = (mag[0] + 0.05 * i, mag[1] - 0.03 * i, mag[2])
        if i == n - 1:
            l = mod.Layer.createSOTLayer(f"L{i}", mod.CVector(*v), mod.CVector(*axes[i]), Ms[i], th[i], SURFACE, _demag(mod, d), 0.02,
                                         120.0, -80.0)
            l.setReferenceLayer(mod.CVector(0, 1, 0))
        else:
            l = mod.Layer(f"L{i}", mod.CVector(*v), mod.CVector(*axes[i]), Ms[i], th[i], SURFACE, _demag(mod, d), ALPHA * (1 + i))
        layers.append(l)
    if strip:  # the Rp/Rap constructor is bilayer-only (junction.hpp:1185-1190)
        r = [10.0 + i for i in range(n)]
        j = mod.Junction(layers, r, [2.0 * x for x in r], [0.3] * n, [0.2] * n, [0.1] * n, [-0.15] * n, [0.4] * n)
    else:
        j = mod.Junction(layers)
    for i in range(n):
        j.setLayerAnisotropyDriver(f"L{i}", mod.ScalarDriver.getConstantDriver(K[i]))
    for i in range(n - 1):
        j.setIECDriver(f"L{i}", f"L{i + 1}", mod.ScalarDriver.getConstantDriver((-1) ** i * 4e-4 * (1 + 0.3 * i)))
        j.setQuadIECDriver(f"L{i}", f"L{i + 1}", mod.ScalarDriver.getConstantDriver(-2e-5 * (1 + i)))
    j.setLayerCurrentDriver(f"L{n - 1}", mod.ScalarDriver.getConstantDriver(3.0))
    j.setLayerOerstedFieldDriver("all", mod.AxialDriver(mod.NullDriver(), mod.NullDriver(), mod.ScalarDriver.getSineDriver(0, 500, 9e9, 0.3)))
    j.setLayerExternalFieldDriver("all", mod.AxialDriver(30e3, -5e3, 10e3))
    return j


def stt_layer_mtj(mod, mag=None, temperature=None):
    """p-bit style STT junction (curated-examples/p-bit.py:188-227 parameters)."""
    if mag is None:
        mag = (math.sin(0.2) * math.cos(1.0), math.sin(0.2) * math.sin(1.0), math.cos(0.2))
    demag = [(0.00024164288391924, 0, 0), (0, 0.000160046006320031, 0), (0, 0, 0.999598310229469)]
    l = mod.Layer.createSTTLayer(
        "free", mod.CVector(*mag), mod.CVector(1, 0, 0), 1.6, 2e-9, math.pi * (0.7e-9) ** 2 * 1e4,
        _demag(mod, demag), 0.1, 1.0, 0.9, 1.0,
    )
    l.setReferenceLayer(mod.CVector(1, 0, 0))
    j = mod.Junction([l], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(350e3))
    j.setLayerCurrentDriver("free", mod.ScalarDriver.getConstantDriver(-1e10))
    j.setLayerExternalFieldDriver("free", mod.AxialDriver(50e3, 0.0, 0.0))
    if temperature is not None:
        j.setLayerTemperatureDriver("free", mod.ScalarDriver.getConstantDriver(temperature))
    return j


def stt_two_layer_mtj(mod, current=None, temperature=None, mag=(0.1, 0.1, 0.95)):
    """Free STT layer on a reference layer (IEC coupled): the usual two-layer MTJ under a current drive."""
    demag = [(0.001, 0, 0), (0, 0.002, 0), (0, 0, 0.997)]
    free = mod.Layer.createSTTLayer("free", mod.CVector(*mag), mod.CVector(0, 0, 1), 1.2, 1.5e-9, SURFACE, _demag(mod, demag), 0.02,
                                    1.0, 0.1, 0.7)
    free.setReferenceLayer(mod.Reference.bottom)
    ref = mod.Layer("bottom", mod.CVector(0.05, 0.0, 1.0), mod.CVector(0, 0, 1), 1.4, 3e-9, SURFACE, _demag(mod, demag), 0.02)
    j = mod.Junction([free, ref], 100, 200)
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(0.6e6))
    j.setLayerAnisotropyDriver("bottom", mod.ScalarDriver.getConstantDriver(1.5e6))
    j.setIECDriver("free", "bottom", mod.ScalarDriver.getConstantDriver(2e-5))
    j.setLayerCurrentDriver("free", current if current is not None else mod.ScalarDriver.getConstantDriver(-3e10))
    j.setLayerExternalFieldDriver("all", mod.AxialDriver(10e3, 0.0, -30e3))
    if temperature is not None:
        j.setLayerTemperatureDriver("all", mod.ScalarDriver.getConstantDriver(temperature))
    return j


def sot_layer(mod, dynamic=True, mag=(0.1, 0.1, 0.9)):
    """1-layer SOT junction with K2 and step torque drivers (CIMS-like)."""
    demag = [(0, 0, 0), (0, 0, 0), (0, 0, 1)]
    if dynamic:
        l = mod.Layer("free", mod.CVector(*mag), mod.CVector(0, 0, 1), 1.0, 1.2e-9, 1e-14, _demag(mod, demag), 0.03)
    else:
        l = mod.Layer.createSOTLayer(
            "free", mod.CVector(*mag), mod.CVector(0, 0, 1), 1.0, 1.2e-9, 1e-14, _demag(mod, demag), 0.03, 2e-9, -3e-9
        )
    l.setReferenceLayer(mod.CVector(0, 1, 0))
    l.setSecondaryReferenceLayer(mod.CVector(1, 0, 0))
    j = mod.Junction([l], [100.0], [1.0], [-0.24], [0.0], [-0.46], [0.0], [-2.7])
    j.setLayerAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(0.6e6))
    j.setLayerSecondOrderAnisotropyDriver("free", mod.ScalarDriver.getConstantDriver(2e4))
    j.setLayerExternalFieldDriver("free", mod.AxialDriver(20e3, 5e3, -10e3))
    if dynamic:
        j.setLayerFieldLikeTorqueDriver("free", mod.ScalarDriver.getStepDriver(0, 3e3, 0.2e-9, 1.5e-9))
        j.setLayerDampingLikeTorqueDriver("free", mod.ScalarDriver.getStepDriver(100.0, -9e3, 0.2e-9, 1.5e-9))
        j.setLayerSecondaryFieldLikeTorqueDriver("free", mod.ScalarDriver.getSineDriver(0, 1e3, 5e9, 0.3))
        j.setLayerSecondaryDampingLikeTorqueDriver("free", mod.ScalarDriver.getConstantDriver(500.0))
    else:
        j.setLayerCurrentDriver("free", mod.ScalarDriver.getSineDriver(1e11, 5e11, 3e9, 0.0))
    return j


def log_arrays(log):
    return {k: np.asarray(v, dtype=np.float64) for k, v in log.items()}


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
