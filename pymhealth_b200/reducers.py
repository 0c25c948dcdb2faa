"""Registry that maps the reference's reducer callables to kernel feature columns.

The reference hands arbitrary Python callables to numba (``rolling_apply(func)``,
src/mhealth/util/windows.py:54-95).  Here a callable is only accepted when it is one of the
reducers the CUDA kernels implement, recognised BY IDENTITY: the numpy aliases the reference
exports (generic/stats.py:156-163), the ``Reducer`` objects of this package, or a
``functools.partial`` of either that binds the reducer's parameter (``q`` / ``th``).  Anything
else raises NotImplementedError -- there is no CPU fallback.
"""
import functools

import numpy as np

from . import _lib as L
from .engine import Feature, window_table


class Reducer:
    """A window reducer with the reference's name and call signature.

    Called directly on a 1-D array it reduces that array as ONE window on the GPU
    (e.g. ``stats.skewness(x)``); handed to ``rolling_apply`` it selects a feature column.
    """

    def __init__(self, name, family, fid, param_name=None, default=None, integer=False, doc=None):
        self.__name__ = name
        self.__qualname__ = name
        self.family = family
        self.fid = fid
        self.param_name = param_name
        self.default = default
        self.integer = integer
        self.__doc__ = doc

    def feature(self, *params):
        if self.param_name is None:
            return Feature(self.family, self.fid, (), self.__name__)
        p = params[0] if params else self.default
        if p is None:
            raise TypeError("%s needs its '%s' argument" % (self.__name__, self.param_name))
        return Feature(self.family, self.fid, (p,), self.__name__)

    def __call__(self, x, *args, **kwargs):
        params = list(args)
        if self.param_name and self.param_name in kwargs:
            params = [kwargs.pop(self.param_name)]
        if kwargs or len(params) > 1:
            raise TypeError("%s: unexpected arguments" % self.__name__)
        a = np.asarray(x)
        if a.ndim != 1:
            a = a.ravel()
        if a.shape[0] == 0:
            raise ValueError("%s of an empty array" % self.__name__)
        if self.param_name and params and np.ndim(params[0]) > 0:
            feats = [self.feature(float(q)) for q in np.asarray(params[0]).ravel()]
            return _one_window(a, feats)
        v = _one_window(a, [self.feature(*params)])[0]
        return int(v) if self.integer else float(v)

    def __repr__(self):
        return "<pymhealth_b200 reducer %s>" % self.__name__


def _one_window(a, feats):
    zc = 0.0
    for f in feats:
        if f.family == "stream" and f.fid == L.F_ZERO_CROSSINGS and f.params:
            zc = f.params[0]
    tab = window_table(a, a.shape[0], 1, feats, zc_threshold=zc)
    return tab[0]


# numpy callables the reference exports as aliases -> (family, id)
_NUMPY = {}


def register_numpy(fn, reducer):
    _NUMPY[fn] = reducer


def resolve(func):
    """callable -> (Feature, integer_flag) or raise NotImplementedError."""
    params = ()
    base = func
    if isinstance(func, functools.partial):
        base = func.func
        if func.args:
            params = tuple(func.args)
        elif func.keywords:
            if len(func.keywords) != 1:
                raise NotImplementedError("partial with several keywords is not a supported reducer: %r" % (func,))
            params = tuple(func.keywords.values())
    from .spectral import SpectralReducer
    from .generic import stats as _stats      # noqa: F401  (registers the numpy aliases)
    if isinstance(base, SpectralReducer):
        if params:
            raise NotImplementedError("spectral reducers carry their parameters; do not wrap them in partial()")
        return base.feature(), base.integer
    red = None
    if isinstance(base, Reducer):
        red = base
    else:
        try:
            red = _NUMPY.get(base)
        except TypeError:
            red = None
    if red is None:
        raise NotImplementedError(
            "rolling_apply: %r is not a reducer the B200 kernels implement (supported: the reducers of "
            "pymhealth_b200.generic.stats / generic.timedom and the numpy aliases np.mean, np.var, np.std, "
            "np.min, np.max, np.sum, np.median, np.percentile (via functools.partial(..., q=..)) ); "
            "there is no CPU fallback" % (func,))
    if params and red.param_name is None:
        raise NotImplementedError("%s takes no parameter" % red.__name__)
    return red.feature(*params), red.integer
