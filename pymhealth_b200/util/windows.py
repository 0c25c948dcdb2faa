"""Rolling window operations -- drop-in for ``mhealth.util.windows``
(reference src/mhealth/util/windows.py).

``rolling_apply(func, wsize=None, wstep=1)`` returns a callable ``(arr, wsize, wstep)`` exactly as
the reference does (:54-95), but the callable launches the CUDA window kernels instead of a numba
``prange`` loop; the list / tuple / dict forms (:98-119) compute ALL their reducers in one pass over
the data instead of one pass per reducer.  Deviations, all documented in DESIGN.md:
  * the dict form returns ``dict(zip(names, values))`` -- the reference returns a set holding a
    zip object (``{zip(names, vals)}``, :116), an obvious bug;
  * ``len(arr) < wsize`` gives an empty result (the reference writes out of bounds, :86-90);
  * an unsupported callable raises NotImplementedError (no CPU fallback).
"""
from functools import lru_cache, singledispatch
from typing import Callable, Optional

import numpy as np
from numpy.lib.stride_tricks import as_strided

from .. import _lib as L
from ..engine import device_get_indices, segment_table, window_table
from ..reducers import resolve


def view(x: np.ndarray, w: int, s: int) -> np.ndarray:
    """Strided (nw, w) window view of ``x`` -- pure indexing, as in the reference (:20-33)."""
    x = np.asarray(x)
    stride = x.strides[0]
    n = x.shape[0]
    return as_strided(x, (((n - w) // s) + 1, w), (s * stride, stride))


def _zc_threshold(features):
    ths = {f.params[0] for f in features if f.family == "stream" and f.fid == L.F_ZERO_CROSSINGS}
    if len(ths) > 1:
        raise NotImplementedError("one zero-crossing threshold per rolling_apply call")
    return ths.pop() if ths else 0.0


def _apply(features, arr, wsize, wstep):
    if wsize is None:
        raise TypeError("wsize must be given (at rolling_apply() or at call time)")
    a = np.asarray(arr)
    if a.ndim == 2:
        # the reference slices axis 0 and hands (W, k) blocks to the reducer; for the
        # order-free reducers that is a 1-D roll over the flattened array
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        k = a.shape[1]
        bad = [f for f in features if f.fid in (L.F_ZERO_CROSSINGS, L.F_LINE_LENGTH, L.F_HJORTH_MOBILITY,
                                                  L.F_HJORTH_COMPLEXITY)]
        if bad:
            raise NotImplementedError("time-domain reducers are defined for 1-D series only")
        a = a.reshape(-1)
        wsize, wstep = int(wsize) * k, int(wstep) * k
    elif a.ndim != 1:
        raise ValueError("rolling_apply takes 1-D (or 2-D, axis-0 windows) arrays")
    fss = {f.fs for f in features if f.family == "spectral"}
    if len(fss) > 1:
        raise NotImplementedError("one sampling rate per rolling_apply call")
    if fss and a.dtype != np.float32:
        a = a.astype(np.float32)        # the spectral kernel computes in float32 (float64 sums)
    tab = window_table(a, int(wsize), int(wstep), features, zc_threshold=_zc_threshold(features),
                       fs=fss.pop() if fss else 1.0)
    return tab          # float64 [nw, n_features]  (windows.py:89: always float64)


@singledispatch
@lru_cache(256)
def rolling_apply(func: Callable, wsize: Optional[int] = None, wstep: int = 1) -> Callable:
    """Create a callable applying ``func`` to the windows of an array (windows.py:54-95)."""
    feature, _ = resolve(func)

    def loop_wrapper(arr, wsize=wsize, wstep=wstep):
        return np.ascontiguousarray(_apply([feature], arr, wsize, wstep)[:, 0])

    loop_wrapper.__doc__ = "Apply the function %s to windows in a given array (CUDA)." % getattr(func, "__name__", func)
    return loop_wrapper


@rolling_apply.register(list)
@rolling_apply.register(tuple)
def _rolling_apply_coll(funcs, wsize: Optional[int] = None, wstep: int = 1) -> Callable:
    features = [resolve(f)[0] for f in funcs]

    def multi_funcs_rolling_apply(arr, wsize=wsize, wstep=wstep):
        tab = _apply(features, arr, wsize, wstep)
        return [np.ascontiguousarray(tab[:, j]) for j in range(len(features))]
    return multi_funcs_rolling_apply


@rolling_apply.register(dict)
def _rolling_apply_dict(funcs, wsize: Optional[int] = None, wstep: int = 1) -> Callable:
    names = list(funcs.keys())
    inner = _rolling_apply_coll(list(funcs.values()), wsize, wstep)

    def dict_funcs_rolling_apply(arr, wsize=wsize, wstep=wstep):
        return dict(zip(names, inner(arr, wsize, wstep)))
    return dict_funcs_rolling_apply


def get_indices(index: np.ndarray, wsize, wstep) -> np.ndarray:
    """Start and end indices of the windows ``[t0 + k*wstep, t0 + k*wstep + wsize)`` over a sorted index
    (windows.py:162-178): device binary search, int64[2, n] back on the host."""
    return device_get_indices(index, wsize, wstep).cpu().numpy()


def _segment_apply(features, indices, arr, min_window_len):
    a = np.asarray(arr)
    if a.ndim != 1:
        raise ValueError("index-addressed windows take a 1-D array")
    tab = segment_table(a, indices, features, min_window_len=min_window_len, zc_threshold=_zc_threshold(features))
    res = tab.cpu().numpy()
    # windows.py:149: the output has the dtype of ``arr``; for integer input the reference stores NaN into an
    # integer array (garbage) -- float64 is returned instead
    if a.dtype == np.float32:
        res = res.astype(np.float32)
    return res


@lru_cache(256)
def indices_rolling_apply(func: Callable, min_window_len: int = 1) -> Callable:
    """Callable ``(indices[2, n], arr)`` applying ``func`` to ``arr[start:end]`` (windows.py:122-159)."""
    feature, _ = resolve(func)

    def windows_loop(indices, arr, min_window_len=min_window_len):
        return np.ascontiguousarray(_segment_apply([feature], indices, arr, min_window_len)[:, 0])
    windows_loop.__doc__ = "Apply the '%s' function to windows with known indices (CUDA)." % getattr(func, "__name__", func)
    return windows_loop


@singledispatch
def nonuniform_rolling_apply(func: Callable, min_window_len: int = 1) -> Callable:
    """Moving-window aggregation over a non-uniform (e.g. datetime) index (windows.py:181-216):
    returns ``moving_window(index, arr, wsize, wstep[, min_window_len])``."""
    feature, _ = resolve(func)

    def moving_window(index, arr, wsize, wstep, min_window_len=min_window_len):
        indices = device_get_indices(index, wsize, wstep)           # stays on the device
        return np.ascontiguousarray(_segment_apply([feature], indices, arr, min_window_len)[:, 0])
    moving_window.__doc__ = "Aggregate windows with the '%s' function (CUDA)." % getattr(func, "__name__", func)
    return moving_window


@nonuniform_rolling_apply.register(list)
@nonuniform_rolling_apply.register(tuple)
def _nu_rolling_apply_coll(funcs, min_window_len: int = 1) -> Callable:
    features = [resolve(f)[0] for f in funcs]

    def moving_window(index, arr, wsize, wstep):
        indices = device_get_indices(index, wsize, wstep)
        tab = _segment_apply(features, indices, arr, min_window_len)     # every reducer in one pass
        return [np.ascontiguousarray(tab[:, j]) for j in range(len(features))]
    return moving_window


@nonuniform_rolling_apply.register(dict)
def _nu_rolling_apply_dict(funcs, min_window_len: int = 1) -> Callable:
    names = list(funcs.keys())
    inner = _nu_rolling_apply_coll(list(funcs.values()), min_window_len)

    def moving_window(index, arr, wsize, wstep):
        return dict(zip(names, inner(index, arr, wsize, wstep)))
    return moving_window
