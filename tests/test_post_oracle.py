"""Pins oracle/post_oracle.py (the CPU restatement of the reference's sweep-procedure reductions):

* the committed goldens tests/golden/post_*.npz ARE what the unmodified reference procedures
  produce (regenerated here in a subprocess when /root/reference is present);
* per-point runs on the C oracle (bit-exact to the reference's runSimulation) reduced by
  post_oracle reproduce those goldens BIT-FOR-BIT."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cmtj_b200 as cm
import helpers
import post_helpers as ph

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import post_scenarios as ps  # noqa: E402


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="/root/reference not present")
def test_committed_goldens_are_what_the_reference_produces(ref, tmp_path):
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_golden_post.py"), "--out", str(tmp_path)])
    for name in ps.SCENARIOS:
        with np.load(tmp_path / f"post_{name}.npz") as fresh:
            want = ph.golden(name)
            assert set(fresh.files) == set(want)
            for k in want:
                assert np.array_equal(fresh[k], want[k]), (name, k)


class _P:
    ResistanceParameters = cm.utils.ResistanceParameters


def test_post_oracle_reproduces_pimm_golden_bit_exact(on_oracle):
    want = ph.golden("pimm_fresh")
    kw = dict(ps.PIMM_KW)
    kw.pop("int_step")
    for i, H in enumerate(ps.H_PIMM):
        got = ph.pimm_point(helpers.two_layer_mtj(cm), H, ps.DT, ps.rparams(_P, 2), **kw)
        assert np.array_equal(got["spectrum"].ravel(), want["spectrum"][i])
        assert np.array_equal(got["freqs"], want["freqs"])
        assert np.array_equal(got["Rx"], want["Rx"][i]) and np.array_equal(got["Ry"], want["Ry"][i])
        assert np.array_equal(got["m_avg"], want["m_avg"][i])


def test_post_oracle_reproduces_full_output_golden_bit_exact(on_oracle):
    want = ph.golden("pimm_fresh_full")
    kw = dict(ps.PIMM_KW)
    kw.pop("int_step")
    for i, H in enumerate(ps.H_PIMM[1:4]):
        got = ph.pimm_point(helpers.tri_layer_mtj(cm), H, ps.DT, ps.rparams(_P, 3), Hoe_direction=cm.Axis.yaxis, wait_time=0.1e-9,
                            parallel=True, **kw)
        assert np.array_equal(got["spectrum"].ravel(), want["spectrum"][i])
        for li, lid in enumerate(["free", "middle", "bottom"]):
            assert np.array_equal(got["layers"][li].ravel(), want["layer_" + lid][i])
        assert np.array_equal(got["Rx"], want["Rx"][i]) and np.array_equal(got["Ry"], want["Ry"][i])
        assert np.array_equal(got["m_traj"][..., -5:], want["m_traj_tail"][i])


@pytest.mark.parametrize("name,build,n,kw", [
    ("vsd_rz_fresh", helpers.two_layer_mtj, 2, dict(Rtype="Rz")),
    ("vsd_rx_fresh", helpers.two_layer_mtj, 2, dict(Rtype="Rx", Hoe_direction=cm.Axis.xaxis)),
    ("vsd_ry_parallel_single", helpers.single_layer_mtj, 1, dict(Rtype="Ry", parallel=True)),
])
def test_post_oracle_reproduces_vsd_golden_bit_exact(on_oracle, name, build, n, kw):
    want = ph.golden(name)["spectrum"]
    for i, H in enumerate(ps.H_VSD):
        for k, f in enumerate(ps.F_VSD):
            got = ph.vsd_point(build(cm), H, f, ps.DT, ps.rparams(_P, n), Hoe_excitation=ps.VSD_KW["Hoe_excitation"],
                               simulation_duration=ps.VSD_KW["simulation_duration"], **kw)
            assert got == want[i, k], (i, k, got, want[i, k])


def test_procedure_argument_validation_matches_the_reference():
    """Raised before any compute call (procedures.py:86-87, :117-118, :245-250)."""
    P = cm.utils
    j = helpers.two_layer_mtj(cm)
    with pytest.raises(ValueError, match="wait_time must be smaller"):
        P.PIMM_procedure(j, ps.H_PIMM, 1e-13, ps.rparams(_P, 2), wait_time=6e-9)
    with pytest.raises(ValueError, match="number of layers in the junction must match"):
        P.PIMM_procedure(j, ps.H_PIMM, 1e-13, ps.rparams(_P, 1))
    with pytest.raises(ValueError, match="Rz can only be used for 2 layer junctions"):
        P.VSD_procedure(helpers.tri_layer_mtj(cm), ps.H_VSD, ps.F_VSD, 1e-13, ps.rparams(_P, 3), Rtype="Rz")
    with pytest.raises(ValueError, match="number of layers in the junction must match"):
        P.VSD_procedure(j, ps.H_VSD, ps.F_VSD, 1e-13, None, Rtype="Rz")
    with pytest.raises(ValueError, match="Rtype must be either"):
        P.VSD_procedure(j, ps.H_VSD, ps.F_VSD, 1e-13, ps.rparams(_P, 2), Rtype="Rq")


def test_sweep_product_order_matches_itertools():
    idx, vals = cm.utils.sweep_product([[1.0, 2.0], [10.0, 20.0, 30.0]])
    assert idx.tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2]]
    assert vals[:, 1].tolist() == [10.0, 20.0, 30.0, 10.0, 20.0, 30.0]


def test_field_scan_matches_reference_formulas():
    H, v = cm.utils.FieldScan.amplitude_scan(-500e3, 500e3, 7, 90, 0)
    assert H.shape == (7,) and v.shape == (7, 3)
    np.testing.assert_allclose(v[:, 0], H, rtol=1e-15)
    Hb, vb = cm.utils.FieldScan.amplitude_scan(0, 1, 3, 30, 60, back=True)
    assert Hb.tolist() == [0.0, 0.5, 1.0, 0.5, 0.0] and vb.shape == (5, 3)
    th, ph_, r = cm.utils.FieldScan.vector2angle(*cm.utils.FieldScan.angle2vector(40, 70, 2.0).tolist())
    np.testing.assert_allclose([th, ph_, r], [40, 70, 2.0], rtol=1e-12)


def test_carried_sweep_through_the_facade_reproduces_golden_bit_exact(on_oracle):
    """The reference's own loop (procedures.py:120-141: field, pulse, seeded disturbance drawn from
    numpy's GLOBAL generator between runs, state carried) written against the facade.  Pins two
    host-side behaviours: runSimulation must not consume numpy's global stream (the Philox keys come
    from a private generator), and the carried magnetisation round-trips through
    getLayerMagnetisation / setLayerMagnetisation unchanged."""
    want = ph.golden("pimm_carried")
    j = helpers.two_layer_mtj(cm)
    np.random.seed(11)
    ids = j.getLayerIds()
    oe = cm.AxialDriver(cm.NullDriver(), cm.NullDriver(), cm.ScalarDriver.getStepDriver(0, 500, 0, ps.DT * 3))
    params = [(p.Rxx0, p.Rxy0, p.Ramr, p.Rahe, p.Rsmr, p.w, p.l) for p in ps.rparams(_P, 2)]
    for k, H in enumerate(ps.H_PIMM[:4]):
        j.clearLog()
        j.setLayerExternalFieldDriver("all", cm.AxialDriver(*[cm.ScalarDriver.getConstantDriver(float(h)) for h in H]))
        j.setLayerOerstedFieldDriver("all", oe)
        for lid in ids:
            o = j.getLayerMagnetisation(lid)
            n = cm.CVector(o.x + np.random.normal(0, 1e-3), o.y + np.random.normal(0, 1e-3), o.z + np.random.normal(0, 1e-3))
            n.normalize()
            j.setLayerMagnetisation(lid, n)
        j.runSimulation(ps.PIMM_KW["simulation_duration"], ps.DT, ps.DT)
        layers = j.layers
        r = ph.post_oracle.pimm_point(j.getLog(), ids, [l.thickness for l in layers], [l.Ms for l in layers], "z", ps.DT, params,
                                      False, 0.0, ps.PIMM_KW["max_frequency"], ps.PIMM_KW["take_last_n"])
        assert np.array_equal(r["spectrum"].ravel(), want["spectrum"][k]), k
        assert np.array_equal(r["m_avg"], want["m_avg"][k]), k


def test_runs_do_not_consume_numpys_global_generator(on_oracle):
    j = helpers.single_layer_mtj(cm)
    j.setLayerTemperatureDriver("all", cm.constantDriver(300.0))
    np.random.seed(3)
    a = np.random.normal()
    np.random.seed(3)
    j.runSimulation(1e-11, 1e-13, 1e-13, solverMode=cm.SolverMode.Heun)
    assert np.random.normal() == a
