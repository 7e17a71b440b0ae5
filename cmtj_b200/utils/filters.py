"""``cmtj.utils.filters.Filters`` (cmtj/utils/filters.py:5-58): the Butterworth low-pass of
``compute_sd`` runs on the device (``cmtjb_compute_sd_device`` / fused in the time loop); only the
coefficient design (scipy.signal.butter, a host-side closed form) happens here."""
from __future__ import annotations

import numpy as np


def butter_lowpass_coefficients(cutoff: float, fs: float, order: int = 3):
    """(b, a) of Filters.butter_lowpass_filter (filters.py:33-44): butter(order, cutoff/(fs/2))."""
    from scipy.signal import butter

    b, a = butter(order, cutoff / (0.5 * fs), btype="low", analog=False)
    return np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64)


class Filters:
    @staticmethod
    def butter_lowpass_filter(data: np.ndarray, cutoff: float, fs: float, order: int = 5):
        """Low-pass a host series (scipy ``lfilter``; host utility for plotting scripts)."""
        from scipy.signal import lfilter

        b, a = butter_lowpass_coefficients(cutoff, fs, order)
        return lfilter(b, a, data, zi=None)

    @staticmethod
    def detrend_axis(arr, axis):
        """Subtract the median along ``axis`` (spectrum plots)."""
        med = np.median(arr, axis=axis)
        return (arr.T - med).T if axis else arr - med
