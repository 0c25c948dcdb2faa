"""HRV time-domain metrics of the reference, restated.  TEST INFRASTRUCTURE ONLY.

Reference: ``src/mhealth/heart/hrv.py`` -- ``td_factor`` (:24-34), ``sdnn`` (:50-63), ``sdann`` (:66-86), ``sdnni``
(:89-108), ``pnn50`` / ``pnnx`` (:111-136), ``rmssd`` (:139-147), ``ssd`` (:150-158), ``sdsd`` (:161-170).
Pinned by tests/golden/ref_extra.npz except ``sdann`` / ``sdnni``: the reference's versions fail to compile under
numba 0.65 (nopython code calling a Python closure) -- parity unpinned for those two; the restatement
follows the source text line by line with the window helpers of oracle/windows.py.
"""
import numpy as np

from . import windows as OW


def td_factor(unit):
    return {"ns": 1., "us": 1e3, "ms": 1e6, "s": 1e9}[unit]


def sdnn(nni):
    return float(np.std(nni))


def pnnx(nni, unit="ms", x=50.0):
    thr = x * 1e6 / td_factor(unit)
    return float(np.sum(np.abs(np.diff(nni)) > thr) / (len(nni) - 1))


def rmssd(nni):
    return float(np.sqrt(np.mean(np.square(np.diff(nni)))))


def ssd(nni):
    return float(np.sum(np.diff(nni)))


def sdsd(nni):
    return float(np.std(np.diff(nni)))


def _index(nni, index, unit):
    if index is None:
        index = np.cumsum(nni) * td_factor(unit)
    return np.asarray(index).astype(np.int64)


def sdann(nni, index=None, interval=300.0, unit=None):
    """hrv.py:83-86: ``_window_mean(index.astype(int), nni, interval, interval).std()`` with interval in ns."""
    step = int(interval * 1e9)
    return float(np.std(OW.nonuniform_rolling("mean", _index(nni, index, unit), nni, step, step)))


def sdnni(nni, index=None, interval=300.0, unit=None):
    """hrv.py:105-108: ``_window_std(index, nni, interval, interval).mean()``."""
    step = int(interval * 1e9)
    return float(np.mean(OW.nonuniform_rolling("std", _index(nni, index, unit), nni, step, step)))
