"""Window drivers of the reference, restated.  TEST INFRASTRUCTURE ONLY.

Reference: ``src/mhealth/util/windows.py`` -- ``rolling_apply`` (:54-95), ``view`` (:20-33),
``get_indices`` (:162-178), ``indices_rolling_apply`` (:122-159),
``nonuniform_rolling_apply`` (:181-216).

The reference builds one jitted closure per reducer; the oracle keeps a table of
pre-declared drivers (one ``prange`` loop per reducer, the reference's only parallel
construct, ``windows.py:68-72``) addressed by feature name.
"""
import numpy as np
from numba import njit, prange

from . import reducers as R


def n_windows(n, wsize, wstep):
    """windows.py:86 -- ``max(0, 1 + (len(arr) - wsize) // wstep)``; the tail is dropped."""
    return max(0, 1 + (n - wsize) // wstep)


def view(x, w, s):
    """windows.py:20-33 -- zero-copy (nw, w) strided view."""
    x = np.asarray(x)
    nw = (x.shape[0] - w) // s + 1
    st = x.strides[0]
    return np.lib.stride_tricks.as_strided(x, (nw, w), (s * st, st), writeable=False)


def _driver0(f):
    """Driver for reducers ``f(window) -> scalar`` (windows.py:68-91)."""
    @njit(parallel=True, cache=False)
    def run(arr, wsize, wstep):
        nw = max(0, 1 + (arr.shape[0] - wsize) // wstep)
        out = np.zeros(nw)           # float64 whatever the input dtype (windows.py:89)
        for i in prange(nw):
            out[i] = f(arr[i * wstep:i * wstep + wsize])
        return out
    return run


def _driver1(f):
    """Driver for reducers with one scalar parameter ``f(window, p)``."""
    @njit(parallel=True, cache=False)
    def run(arr, wsize, wstep, p):
        nw = max(0, 1 + (arr.shape[0] - wsize) // wstep)
        out = np.zeros(nw)
        for i in prange(nw):
            out[i] = f(arr[i * wstep:i * wstep + wsize], p)
        return out
    return run


_PLAIN = {
    "mean": R.w_mean, "var": R.w_var, "std": R.w_std, "min": R.w_min, "max": R.w_max,
    "drange": R.w_drange, "skewness": R.w_skewness, "kurtosis": R.w_kurtosis,
    "kurtosis_excess": R.w_kurtosis_excess, "coeff_var": R.w_coeff_var,
    "median": R.w_median, "iqr": R.w_iqr, "mode": R.w_mode,
    "line_length": R.w_line_length, "hjorth_activity": R.w_hjorth_activity,
    "hjorth_mobility": R.w_hjorth_mobility, "hjorth_complexity": R.w_hjorth_complexity,
}
_PARAM = {"percentile": R.w_percentile, "zero_crossing_count": R.w_zero_crossing_count}
_drivers = {}

FEATURES = tuple(_PLAIN) + tuple(_PARAM)


def rolling(name, arr, wsize, wstep, param=None):
    """Oracle of ``rolling_apply(f)(arr, wsize, wstep)`` for the reducer called ``name``.

    ``arr`` is promoted to float64 first: the reference is always fed float64 views of the
    same float32 values the GPU sees, so only arithmetic order differs.
    Returns float64[nw].
    """
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("oracle.rolling takes 1-D series")
    if a.shape[0] < wsize:
        # windows.py:86-90 would write out[0] out of bounds here; the
        # defined behaviour both sides agree on is "no windows".
        return np.zeros(0)
    if name in _PLAIN:
        if name not in _drivers:
            _drivers[name] = _driver0(_PLAIN[name])
        return _drivers[name](a, wsize, wstep)
    if name in _PARAM:
        if name not in _drivers:
            _drivers[name] = _driver1(_PARAM[name])
        p = 0.0 if param is None else float(param)
        return _drivers[name](a, wsize, wstep, p)
    raise KeyError(name)


def rolling_table(names, arr, wsize, wstep, params=None):
    """Column-stack of ``rolling`` over several reducers -> float64[nw, len(names)]."""
    params = params or {}
    cols = [rolling(n.split(":")[0], arr, wsize, wstep, params.get(n)) for n in names]
    return np.stack(cols, axis=1) if cols else np.zeros((n_windows(len(arr), wsize, wstep), 0))


# ----------------------------------------------------------------- non-uniform windows
def get_indices(index, wsize, wstep):
    """windows.py:162-178: starts = arange(index[0], index[-1], wstep); ends = starts + wsize;
    left ``searchsorted`` of both into ``index`` -> int64[2, n]."""
    index = np.asarray(index)
    starts = np.arange(index[0], index[-1], wstep)
    ends = starts + wsize
    both = np.concatenate((starts, ends))
    return np.searchsorted(index, both).reshape((2, len(starts)))


def indices_rolling(name, indices, arr, min_window_len=1, param=None):
    """windows.py:134-157: serial loop over [start, end) pairs; windows shorter than
    ``min_window_len`` give NaN.  Output dtype follows ``arr`` (windows.py:149) -- the oracle
    is only defined for floating inputs (NaN into an integer array is garbage)."""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    f = _PLAIN[name] if name in _PLAIN else _PARAM[name]
    n = indices.shape[1]
    out = np.zeros(n, dtype=np.float64)
    for i in range(n):
        si, ei = int(indices[0, i]), int(indices[1, i])
        if ei - si >= min_window_len:
            out[i] = f(a[si:ei]) if name in _PLAIN else f(a[si:ei], 0.0 if param is None else float(param))
        else:
            out[i] = np.nan
    return out.astype(np.asarray(arr).dtype if np.asarray(arr).dtype.kind == "f" else np.float64)


def nonuniform_rolling(name, index, arr, wsize, wstep, min_window_len=1, param=None):
    """windows.py:198-216."""
    return indices_rolling(name, get_indices(index, wsize, wstep), arr, min_window_len, param)
