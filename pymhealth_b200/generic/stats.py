"""Generic statistical reducers -- drop-in for ``mhealth.generic.stats``
(reference src/mhealth/generic/stats.py).  Same names; each object reduces one window on the GPU
when called, and selects a kernel column when handed to ``rolling_apply``."""
import numpy as np

from .. import _lib as L
from ..reducers import Reducer, register_numpy

mean = Reducer("mean", "stream", L.F_MEAN, doc="np.mean (stats.py:157)")
var = Reducer("var", "stream", L.F_VAR, doc="np.var, population (stats.py:160)")
std = Reducer("std", "stream", L.F_STD, doc="np.std (stats.py:159)")
dmin = Reducer("dmin", "stream", L.F_MIN, doc="np.min (stats.py:161)")
dmax = Reducer("dmax", "stream", L.F_MAX, doc="np.max (stats.py:162)")
drange = Reducer("drange", "stream", L.F_DRANGE, doc="max - min (stats.py:34-45)")
skewness = Reducer("skewness", "stream", L.F_SKEWNESS, doc="biased skewness, 0 for a constant window (stats.py:97-110)")
kurtosis = Reducer("kurtosis", "stream", L.F_KURTOSIS, doc="mu4 / mu2^2, 0 for a constant window (stats.py:113-126)")
kurtosis_excess = Reducer("kurtosis_excess", "stream", L.F_KURTOSIS_EXCESS, doc="kurtosis - 3 (stats.py:129-139)")
coeff_var = Reducer("coeff_var", "stream", L.F_COEFF_VAR, doc="std / mean (stats.py:142-153)")
median = Reducer("median", "order", L.F_MEDIAN, doc="np.median (stats.py:158)")
percentile = Reducer("percentile", "order", L.F_PERCENTILE, param_name="q",
                     doc="np.percentile(x, q) with numba's interpolation (stats.py:163)")
interquartile_range = Reducer("interquartile_range", "order", L.F_IQR, doc="p75 - p25 (stats.py:48-59)")
mode = Reducer("mode", "order", L.F_MODE, doc="most frequent element, jit semantics (stats.py:62-94)")
_sum = Reducer("sum", "stream", L.F_SUM, doc="np.sum")

for _np, _r in ((np.mean, mean), (np.var, var), (np.std, std), (np.min, dmin), (np.max, dmax),
                (np.amin, dmin), (np.amax, dmax), (np.sum, _sum), (np.median, median),
                (np.percentile, percentile)):
    register_numpy(_np, _r)

absolute = np.absolute     # elementwise alias kept for API parity (stats.py:156); not a reducer


def minmax(x):
    """(min, max) of an array (stats.py:12-31) -- one kernel pass."""
    from ..reducers import _one_window
    a = np.asarray(x).ravel()
    lo, hi = _one_window(a, [dmin.feature(), dmax.feature()])
    return (a.dtype.type(lo), a.dtype.type(hi))
