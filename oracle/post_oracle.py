"""CPU restatement of the reductions the reference's sweep procedures apply to ONE runSimulation
log -- TEST INFRASTRUCTURE ONLY (imported by tests/ only;
never by the product, which does these reductions on the GPU).

Follows cmtj/utils/procedures.py and cmtj/utils/resistance.py of the reference line by line, with
the same numpy / scipy calls (the FFT, `butter` and `lfilter` live in scipy, a third-party
dependency of the reference: scipy 1.18.1 in this image), so that applied to a log that is
bit-identical to the reference's it returns bit-identical results.  Pinned in
tests/test_post_oracle.py against goldens generated from the unmodified reference procedures
(tests/golden/make_golden_post.py).
"""
import math

import numpy as np
from scipy.fft import fft, fftfreq
from scipy.signal import butter, lfilter


def compute_spectrum_strip(input_m, int_step, max_frequency):
    """procedures.py:29-40"""
    yf = np.abs(fft(input_m))
    freqs = fftfreq(len(yf), int_step)
    freqs = freqs[: len(freqs) // 2]
    yf = yf[: len(yf) // 2]
    findx = np.argwhere(freqs <= max_frequency)
    return yf[findx], freqs[findx]


def compute_resistance(Rx0, Ry0, AMR, AHE, SMR, m, l, w):
    """resistance.py:25-66"""
    n = len(Rx0)
    m = np.asarray(m)
    shape = (n,) if m.ndim == 2 else (n, m.shape[2])
    Sx, Sy = np.zeros(shape), np.zeros(shape)
    for i in range(n):
        w_l = w[i] / l[i]
        Sx[i] = Rx0[i] + (AMR[i] * m[i, 0] ** 2 + SMR[i] * m[i, 1] ** 2)
        Sy[i] = Ry0[i] + 0.5 * AHE[i] * m[i, 2] + (w_l) * (SMR[i] + AMR[i]) * m[i, 0] * m[i, 1]
    return Sx, Sy


def calculate_resistance(params, m, parallel):
    """resistance.py:96-149; params = list of (Rxx0, Rxy0, Ramr, Rahe, Rsmr, w, l) per layer."""
    Sx, Sy = compute_resistance(
        [p[0] for p in params], [p[1] for p in params], [p[2] for p in params], [p[3] for p in params],
        [p[4] for p in params], m, l=[p[6] for p in params], w=[p[5] for p in params],
    )  # fmt: skip
    if parallel:
        return 1.0 / np.sum(1.0 / Sx, axis=0), 1.0 / np.sum(1.0 / Sy, axis=0)
    return np.sum(Sx, axis=0), np.sum(Sy, axis=0)


def compute_sd(dynamic_r, dynamic_i, integration_step):
    """resistance.py:12-22 + filters.py:33-44"""
    SD = -np.asarray(dynamic_i) * np.asarray(dynamic_r)
    fs = 1.0 / integration_step
    b, a = butter(3, 10e6 / (0.5 * fs), btype="low", analog=False)
    return np.mean(lfilter(b, a, SD, zi=None))


def pimm_point(log, layer_ids, thickness, Ms, comp, int_step, params, parallel=False, wait_time=0.0, max_frequency=80e9,
               take_last_n=100):
    """The body of PIMM_procedure's field loop after runSimulation (procedures.py:141-193) for one
    log dict.  Returns spectrum, freqs, per-layer spectra, Rx, Ry, m_avg."""
    nf = np.sum([t * ms for t, ms in zip(thickness, Ms)])
    indx = np.argwhere(np.asarray(log["time"]) >= wait_time).ravel()
    m_traj = np.asarray([
        np.asarray([log[f"{lid}_mx"], log[f"{lid}_my"], log[f"{lid}_mz"]]) * t * ms / nf
        for lid, t, ms in zip(layer_ids, thickness, Ms)
    ])
    m = m_traj[:, :, -take_last_n:]
    Rx, Ry = calculate_resistance(params, m, parallel)
    mixed = np.asarray([np.asarray(log[f"{lid}_m{comp}"])[indx] * t * ms / nf for lid, t, ms in zip(layer_ids, thickness, Ms)])
    yf, freqs = compute_spectrum_strip(mixed.sum(axis=0), int_step, max_frequency)
    per_layer = [compute_spectrum_strip(mixed[i], int_step, max_frequency)[0] for i in range(len(layer_ids))]
    return {"spectrum": yf, "freqs": freqs, "layers": per_layer, "Rx": Rx, "Ry": Ry, "m_avg": m_traj[:, :, -1].sum(0),
            "m_traj": m_traj}


def vsd_point(log, layer_ids, frequency, int_step, Rtype="Rz", params=None, parallel=False):
    """simulate_VSD after runSimulation (procedures.py:283-329) for one log dict."""
    if Rtype == "Rz":
        R = log[f"R_{layer_ids[0]}_{layer_ids[1]}"]
    else:
        m_traj = np.asarray([[log[f"{lid}_mx"], log[f"{lid}_my"], log[f"{lid}_mz"]] for lid in layer_ids])
        Rx, Ry = calculate_resistance(params, m_traj, parallel)
        R = Rx if Rtype == "Rx" else Ry
    dynamicI = np.sin(2 * math.pi * frequency * np.asarray(log["time"]))
    return compute_sd(R, dynamicI, int_step)
