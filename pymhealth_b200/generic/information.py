"""Information-theory reducers -- drop-in for ``mhealth.generic.information.entropy``
(reference src/mhealth/generic/information.py:10-20).  ``sampen`` (O(n^2) sample entropy) is out
of the hot path."""
import numpy as np

from .. import _lib as L
from ..spectral import psd_reduce


def entropy(x):
    """Shannon entropy of counts / probabilities: p = x / sum(x); p += 1e-30; -sum(p ln p)."""
    a = np.asarray(x, dtype=np.float64).ravel()
    return float(psd_reduce(a, None, [(L.S_ENTROPY, None, None)])[0])
