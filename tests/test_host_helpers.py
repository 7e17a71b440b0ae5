"""Host-side helpers that need no GPU: the lazily materialised per-point lists of the batched procedures, the
thermal-normals setting, argument checks of the sharded-array entry point."""
import pickle

import numpy as np
import pytest

import cmtj_b200
from cmtj_b200.utils.procedures import _Rows


def test_rows_behave_like_the_reference_lists():
    a = np.arange(12.0).reshape(4, 3)
    r = _Rows(a)
    assert isinstance(r, list)
    assert np.asarray(r) is a  # no copy, no per-row objects, until somebody asks for list semantics
    assert len(r) == 4 and np.array_equal(r[1], a[1]) and np.array_equal(r[-1], a[3])
    assert [float(x.sum()) for x in r] == [3.0, 12.0, 21.0, 30.0]
    assert np.array_equal(np.array(r), a)
    r.append(np.zeros(3))  # the reference appends per field point (procedures.py:183-193)
    assert len(r) == 5 and np.asarray(r).shape == (5, 3)
    assert len(list(_Rows(a))) == 4 and len(_Rows(a) + [1]) == 5 and bool(_Rows(a))
    assert len(pickle.loads(pickle.dumps(_Rows(a)))) == 4
    assert np.asarray(_Rows(a), dtype=np.float32).dtype == np.float32
    assert len(_Rows(np.zeros((0, 3)))) == 0 and not _Rows(np.zeros((0, 3)))


def test_thermal_normals_setting():
    prev = cmtj_b200.thermal_normals()
    try:
        assert cmtj_b200.thermal_normals("float64") == "float64"
        assert cmtj_b200.core._thermal_normals == 1
        assert cmtj_b200.thermal_normals("float32") == "float32"
        with pytest.raises(ValueError):
            cmtj_b200.thermal_normals("float16")
    finally:
        cmtj_b200.thermal_normals(prev)


def test_run_args_carry_the_normals_precision(on_oracle):
    from cmtj_b200 import _coupled

    prev = cmtj_b200.thermal_normals()
    try:
        cmtj_b200.thermal_normals("float64")
        assert _coupled.run_args(1e-9, 1e-13, 1e-12, 1).thermal_normals == 1
        cmtj_b200.thermal_normals("float32")
        assert _coupled.run_args(1e-9, 1e-13, 1e-12, 1).thermal_normals == 0
    finally:
        cmtj_b200.thermal_normals(prev)


def test_sharded_run_rejects_unknown_transport(on_oracle):
    from test_sharded_group import lattice_group

    g = lattice_group(4)
    with pytest.raises(ValueError):
        g.run_sharded(1e-12, 1e-13, 1e-13, rank=0, world=1, transport="smoke-signals", _host_memory=True)
    with pytest.raises(ValueError):  # peer memory is device memory: no callback, no host stand-in
        g.run_sharded(1e-12, 1e-13, 1e-13, rank=0, world=1, transport="peer", _host_memory=True)
