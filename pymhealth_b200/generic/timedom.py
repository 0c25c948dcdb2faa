"""Time-domain reducers -- drop-in for the window reductions of ``mhealth.generic.timedom``
(reference src/mhealth/generic/timedom.py:34-193).  dfa / hurst / o1fit are whole-signal
polynomial fits, not window reductions, and are out of scope."""
import numpy as np

from .. import _lib as L
from ..reducers import Reducer, _one_window

zero_crossing_count = Reducer("zero_crossing_count", "stream", L.F_ZERO_CROSSINGS, param_name="th", default=0.0,
                              integer=True, doc="number of sign changes after zeroing |x| <= th (timedom.py:52-64)")
line_length = Reducer("line_length", "stream", L.F_LINE_LENGTH, doc="sum |x[i+1] - x[i]| (timedom.py:67-78)")
hjorth_activity = Reducer("hjorth_activity", "stream", L.F_HJORTH_ACTIVITY, doc="variance (timedom.py:81-95)")
hjorth_mobility = Reducer("hjorth_mobility", "order", L.F_HJORTH_MOBILITY,
                          doc="sqrt(var(gradient(x)) / var(x)) (timedom.py:98-114)")
hjorth_complexity = Reducer("hjorth_complexity", "order", L.F_HJORTH_COMPLEXITY,
                            doc="mobility(gradient(x)) / mobility(x) (timedom.py:135-151)")


def hjorth_parameters(x):
    """(activity, mobility, complexity) (timedom.py:167-193) -- one staging of the window."""
    a = np.asarray(x).ravel()
    v = _one_window(a, [hjorth_activity.feature(), hjorth_mobility.feature(), hjorth_complexity.feature()])
    return (float(v[0]), float(v[1]), float(v[2]))
