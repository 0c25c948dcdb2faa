"""Spectrum-density reducers -- drop-in for ``mhealth.generic.frequency.density``
(reference src/mhealth/generic/frequency/density.py)."""
import numpy as np

from ... import _lib as L
from ...spectral import psd_reduce


def first_index(arr, x):
    """First i with x <= arr[i], else len(arr) (density.py:9-14) -- host-side index helper."""
    a = np.asarray(arr)
    hits = np.nonzero(x <= a)[0]
    return int(hits[0]) if hits.size else len(a)


def peak_frequency(psd, freqs, lower=None, upper=None):
    """freqs[lidx + argmax(psd[lidx:uidx])]: lower inclusive, upper EXCLUSIVE, first maximum
    (density.py:18-32)."""
    return float(psd_reduce(psd, freqs, [(L.S_PEAK_FREQUENCY, lower, upper)])[0])
