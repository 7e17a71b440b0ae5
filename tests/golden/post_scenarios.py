"""Sweep-procedure scenarios shared by make_golden_post.py (run against the UNMODIFIED reference:
oracle/_ref extension + /root/reference/cmtj/utils/procedures.py) and the parity tests of
cmtj_b200.utils.procedures.  `m` exposes the core API, `P` the procedures module
(PIMM_procedure, VSD_procedure, ResistanceParameters).  Every scenario returns {name: array}.

"fresh" scenarios give every sweep point its own junction (what the batched device procedure
computes); "carried" scenarios pass the whole sweep to one call (the reference's hysteresis
semantics; ours with the default carry_state=True; the fresh-start scenarios pass carry_state=False to ours)."""
import sys

import numpy as np

import helpers


def _parallel_fn(P):
    """calculate_resistance_parallel of the resistance module P's procedures were written against."""
    return sys.modules[P.calculate_resistance_series.__module__].calculate_resistance_parallel

DT = 1e-13
H_PIMM = np.array([[-300e3, 2e3, 8e3], [-120e3, 5e3, 1e3], [10e3, 1e3, 3e3], [150e3, -4e3, 2e3], [400e3, 1e3, -6e3]])
PIMM_KW = dict(int_step=DT, simulation_duration=0.5e-9, take_last_n=20, max_frequency=80e9, Hoe_excitation=500, Hoe_duration=3)
H_VSD = np.array([[40e3, 8e3, 150e3], [-90e3, 3e3, 260e3]])
F_VSD = np.array([4e9, 7e9, 12e9])
VSD_KW = dict(int_step=DT, simulation_duration=1e-9, Hoe_excitation=400)


def rparams(P, n):
    base = [P.ResistanceParameters(Rxx0=100, Rxy0=1, Rahe=1, Rsmr=-0.2, Ramr=0.5, w=3e-5, l=2e-5),
            P.ResistanceParameters(Rxx0=80, Rxy0=2, Rahe=-0.5, Rsmr=0.1, Ramr=-0.3, w=3e-5, l=2e-5),
            P.ResistanceParameters(Rxx0=120, Rxy0=0.5, Rahe=0.3, Rsmr=0.05, Ramr=0.2, w=1e-5, l=2e-5)]
    return base[:n]


def _pack_pimm(prefix, out, spectrum, freqs, other, layer_ids=None):
    out[prefix + "spectrum"] = np.asarray(spectrum, dtype=np.float64)
    out[prefix + "freqs"] = np.asarray(freqs, dtype=np.float64)
    for k in ("Rx", "Ry", "m_avg", "H"):
        out[prefix + k] = np.asarray(other[k], dtype=np.float64)
    if layer_ids:
        for lid in layer_ids:
            out[prefix + "layer_" + lid] = np.asarray(other[lid], dtype=np.float64)
        out[prefix + "m_traj_tail"] = np.asarray(other["m_traj"], dtype=np.float64)[..., -5:]


def pimm_fresh(m, P, batched):
    """2-layer MTJ, 5 field points, z pulse, every point from the same (un-normalised) initial state."""
    out = {}
    kw = dict(PIMM_KW, resistance_params=rparams(P, 2), disturbance=0, disable_tqdm=True)
    if batched:
        _pack_pimm("", out, *P.PIMM_procedure(helpers.two_layer_mtj(m), H_PIMM, carry_state=False, **kw))
    else:
        parts = [P.PIMM_procedure(helpers.two_layer_mtj(m), H_PIMM[i : i + 1], **kw) for i in range(len(H_PIMM))]
        other = {k: [p[2][k][0] for p in parts] for k in ("Rx", "Ry", "m_avg", "H")}
        _pack_pimm("", out, np.asarray([p[0] for p in parts]), parts[0][1], other)
    return out


def pimm_fresh_full(m, P, batched):
    """3-layer junction, y pulse, wait_time > 0, parallel resistance, full_output."""
    out = {}
    kw = dict(PIMM_KW, resistance_params=rparams(P, 3), disturbance=0, disable_tqdm=True, full_output=True, wait_time=0.1e-9,
              Hoe_direction=m.Axis.yaxis, resistance_fn=_parallel_fn(P))
    H = H_PIMM[1:4]
    ids = ["free", "middle", "bottom"]
    if batched:
        _pack_pimm("", out, *P.PIMM_procedure(helpers.tri_layer_mtj(m), H, carry_state=False, **kw), layer_ids=ids)
    else:
        parts = [P.PIMM_procedure(helpers.tri_layer_mtj(m), H[i : i + 1], **kw) for i in range(len(H))]
        other = {k: [p[2][k][0] for p in parts] for k in ("Rx", "Ry", "m_avg", "H", "m_traj")}
        for lid in ids:
            other[lid] = [p[2][lid] for p in parts]
        _pack_pimm("", out, np.asarray([p[0] for p in parts]), parts[0][1], other, layer_ids=ids)
    return out


def pimm_carried(m, P, batched):
    """The reference's own semantics: one call over the sweep, state carried, seeded disturbance."""
    out = {}
    kw = dict(PIMM_KW, resistance_params=rparams(P, 2), disable_tqdm=True)
    extra = {}  # carried state is the default of both implementations
    np.random.seed(11)
    _pack_pimm("", out, *P.PIMM_procedure(helpers.two_layer_mtj(m), H_PIMM[:4], disturbance=1e-3, **kw, **extra))
    return out


def pimm_static(m, P, batched):
    """static_only=True: no spectrum, resistances and m_avg only (AHE-loop use of the procedure)."""
    kw = dict(PIMM_KW, resistance_params=rparams(P, 2), disable_tqdm=True, disturbance=0, static_only=True)
    extra = {}  # carried state is the default of both implementations
    spectrum, freqs, other = P.PIMM_procedure(helpers.two_layer_mtj(m), H_PIMM[:3], **kw, **extra)
    assert freqs is None and np.asarray(spectrum).size == 0
    return {k: np.asarray(other[k], dtype=np.float64) for k in ("Rx", "Ry", "m_avg")}


def _vsd(m, P, batched, build, **kw):
    if batched:
        return {"spectrum": P.VSD_procedure(build(m), H_VSD, F_VSD, disturbance=0, disable_tqdm=True, carry_state=False, **VSD_KW, **kw)}
    s = np.zeros((len(H_VSD), len(F_VSD)))
    for i in range(len(H_VSD)):
        for k in range(len(F_VSD)):
            s[i, k] = P.VSD_procedure(build(m), H_VSD[i : i + 1], F_VSD[k : k + 1], disturbance=0, disable_tqdm=True,
                                      **VSD_KW, **kw)[0, 0]
    return {"spectrum": s}


def vsd_rz_fresh(m, P, batched):
    """2-layer MTJ, GMR channel mixed with the excitation (Rtype='Rz')."""
    return _vsd(m, P, batched, helpers.two_layer_mtj, resistance_params=rparams(P, 2), Rtype="Rz")


def vsd_rx_fresh(m, P, batched):
    """2-layer, Rx of the series strip model, x excitation."""
    return _vsd(m, P, batched, helpers.two_layer_mtj, resistance_params=rparams(P, 2), Rtype="Rx", Hoe_direction=m.Axis.xaxis)


def vsd_ry_parallel_single(m, P, batched):
    """1-layer junction, Ry of the parallel model."""
    return _vsd(m, P, batched, helpers.single_layer_mtj, resistance_params=rparams(P, 1), Rtype="Ry",
                resistance_fn=_parallel_fn(P))


def vsd_carried(m, P, batched):
    extra = {}  # carried state is the default of both implementations
    np.random.seed(5)
    return {"spectrum": P.VSD_procedure(helpers.two_layer_mtj(m), H_VSD, F_VSD[:2], disturbance=1e-3, disable_tqdm=True,
                                        resistance_params=rparams(P, 2), Rtype="Rz", **VSD_KW, **extra)}


SCENARIOS = {
    "pimm_fresh": pimm_fresh,
    "pimm_fresh_full": pimm_fresh_full,
    "pimm_carried": pimm_carried,
    "pimm_static": pimm_static,
    "vsd_rz_fresh": vsd_rz_fresh,
    "vsd_rx_fresh": vsd_rx_fresh,
    "vsd_ry_parallel_single": vsd_ry_parallel_single,
    "vsd_carried": vsd_carried,
}
