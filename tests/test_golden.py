"""Committed golden vectors (tests/golden/*.npz, generated from the unmodified reference by
tests/golden/make_golden.py):
  * CPU (`-m "not gpu"`): the oracle restatement, driven through the facade, reproduces them
    BIT-FOR-BIT, and the reference's own test_sto golden (3-decimal) at full length;
  * GPU (`-m gpu`): the CUDA path reproduces them to <= 1e-9 relative (north_star)."""
import json
import os
import sys

import numpy as np
import pytest

import cmtj_b200
import helpers

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
import scenarios  # noqa: E402

NAMES = sorted(scenarios.SCENARIOS)


def _golden(name):
    with np.load(os.path.join(HERE, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_bit_exact(on_oracle, name):
    got, want = scenarios.SCENARIOS[name](cmtj_b200), _golden(name)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), (k, helpers.max_rel_err(got[k], want[k]))


def _full_length_sto(m):
    demag = [m.CVector(0, 0, 0), m.CVector(0, 0, 0), m.CVector(0, 0, 1)]
    l1 = m.Layer.createSTTLayer("free", m.CVector(0, 0, 1), m.CVector(0, 0, 1), 1.0, 1.4e-9, 7e-10 * 7e-10, demag, 0.3, 1.0, 1, 1.0)
    l1.setReferenceLayer(m.CVector(0, 1.0, 1.0))
    j = m.Junction([l1], 100, 200)
    j.setLayerAnisotropyDriver("free", m.constantDriver(350e3))
    j.setLayerCurrentDriver("free", m.constantDriver(-6e9))
    j.runSimulation(50e-9, 1e-13, 1e-13)
    log = j.getLog()
    return {k: np.asarray(log[s][-100:]) for k, s in (("mx", "free_mx"), ("my", "free_my"), ("mz", "free_mz"), ("R", "R"))}


def _check_ref_sto(got):
    want = json.load(open(os.path.join(HERE, "ref_test_sto.json")))
    for k in ("mx", "my", "mz", "R"):
        # the reference's own comparison (tests/test_curated_examples.py:37-39, :147-149)
        np.testing.assert_array_almost_equal(np.round(got[k], 3), np.asarray(want[k]), decimal=3)


def test_oracle_reproduces_reference_sto_golden(on_oracle):
    """The reference's direct trajectory pin: last 100 samples of a 5e5-step STT oscillator run."""
    _check_ref_sto(_full_length_sto(cmtj_b200))


# ---- GPU ---------------------------------------------------------------------------------------
# relative tolerance per scenario: 1e-9 (north_star) except Dormand-Prince, whose accepted-step
# sequence feeds back through pow(err, -0.2) (1e-8, same bar as tests/test_gpu_parity.py).
GPU_TOL = {"dormand_prince": 1e-8}


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_golden(name):
    got, want = scenarios.SCENARIOS[name](cmtj_b200), _golden(name)
    assert set(got) == set(want)
    tol = GPU_TOL.get(name, 1e-9)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert helpers.max_rel_err(got[k], want[k]) <= tol, (k, helpers.max_rel_err(got[k], want[k]))


@pytest.mark.gpu
def test_cuda_reproduces_reference_sto_golden():
    _check_ref_sto(_full_length_sto(cmtj_b200))
