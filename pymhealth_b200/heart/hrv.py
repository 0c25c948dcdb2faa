"""Frequency-domain PSD reducers of ``mhealth.heart.hrv`` (reference src/mhealth/heart/hrv.py:172-198).
The RR-interval time-domain metrics of that module are out of the hot path."""
import numpy as np

from .. import _lib as L
from ..spectral import psd_reduce


def power_band(psd, freqs, lower=None, upper=None):
    """sum |psd| over lower <= f <= upper, both inclusive (hrv.py:173-179)."""
    return float(psd_reduce(psd, freqs, [(L.S_BAND_POWER, lower, upper)])[0])


def relative_power_band(psd, freqs, lower=None, upper=None):
    """power_band / sum |psd| (hrv.py:192-198)."""
    return float(psd_reduce(psd, freqs, [(L.S_REL_BAND_POWER, lower, upper)])[0])


def peak_frequency(psd, freqs, lower=None, upper=None):
    """The reference's hrv.peak_frequency (hrv.py:182-189) indexes the UNMASKED frequency vector with
    the argmax of the MASKED psd and is wrong whenever lower > min(freqs).  This function
    implements what it plainly intends -- the frequency of the largest PSD value inside
    lower <= f <= upper -- and says so; use generic.frequency.density.peak_frequency for the
    reference's other (correct, upper-exclusive) variant."""
    f = np.asarray(freqs, dtype=np.float64)
    hi = None
    if upper is not None:
        above = f[f > upper]
        hi = float(above.min()) if above.size else None     # inclusive upper bound -> exclusive bound at the next bin
    return float(psd_reduce(psd, freqs, [(L.S_PEAK_FREQUENCY, lower, hi)])[0])
