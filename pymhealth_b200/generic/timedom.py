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


def _series(x):
    from ..engine import require_cuda
    torch = require_cuda()
    a = np.asarray(x)
    if a.ndim != 1:
        raise ValueError("1-D array expected")
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return torch, torch.from_numpy(np.array(a, order="C", copy=True)).cuda()


def gradient(x):
    """Derivative of the input: halved distance between x[i-1] and x[i+1], one-sided at the ends; float64 output,
    differences in the input type (timedom.py:11-31)."""
    from ..engine import _stream_ptr
    torch, t = _series(x)
    n = t.shape[0]
    if n < 2:
        raise ValueError("gradient needs at least two samples")
    out = torch.empty(n, dtype=torch.float64, device=t.device)
    L.check(L.load().mhb_gradient(1 if t.dtype == torch.float64 else 0, t.data_ptr(), n, out.data_ptr(), _stream_ptr(torch)),
            "gradient")
    return out.cpu().numpy()


def zero_crossings(x, th=0):
    """Boolean array: was there a zero crossing between samples i and i + 1, after zeroing |x| <= th
    (timedom.py:34-49)."""
    from ..engine import _stream_ptr
    torch, t = _series(x)
    n = t.shape[0]
    out = torch.zeros(max(0, n - 1), dtype=torch.uint8, device=t.device)
    L.check(L.load().mhb_zero_crossings(1 if t.dtype == torch.float64 else 0, t.data_ptr(), n, float(th), out.data_ptr(),
                                        _stream_ptr(torch)), "zero_crossings")
    return out.cpu().numpy().astype(bool)
