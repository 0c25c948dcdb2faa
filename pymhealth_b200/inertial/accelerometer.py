"""Accelerometer orientation / magnitude -- drop-in for the elementwise functions of
``mhealth.inertial.accelerometer`` (reference src/mhealth/inertial/accelerometer.py:13-74, 198-265).

The reference jit-compiles these with numba and dispatches on ``pd.DataFrame`` with
``functools.singledispatch``; here the array forms launch the CUDA pre-stage kernels (csrc/accel.cu) and the
DataFrame forms are the same thin wrappers.  Result types follow numba's [probed against the reference]:
``magnitude`` keeps the input float type (float32 in -> float32 arithmetic, bit-identical), ``roll`` / ``pitch``
return float64 degrees, integer input is promoted to float64.  The Butterworth ``linear_filter`` /
``gravity_filter`` of that module are scipy IIR recurrences and out of scope.
"""
from functools import singledispatch

import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr
from ..util.deps import pd

_MAGNITUDE, _ROLL, _PITCH = 0, 1, 2


def _prep(arrs):
    """numpy / torch / scalars -> (list of cuda 1-D tensors of one float dtype, shape, was_torch, scalar)."""
    torch = require_cuda()
    was_torch = any(isinstance(a, torch.Tensor) for a in arrs)
    if was_torch:
        ts = [a if isinstance(a, torch.Tensor) else torch.as_tensor(a) for a in arrs]
        dt = torch.float32 if all(t.dtype == torch.float32 for t in ts) else torch.float64
        ts = [t.to(device="cuda", dtype=dt) for t in ts]
        ts = torch.broadcast_tensors(*ts)
        shape = tuple(ts[0].shape)
        return [t.contiguous().reshape(-1) for t in ts], shape, True, False
    nps = [np.asarray(a) for a in arrs]
    scalar = all(a.ndim == 0 for a in nps)
    dt = np.float32 if all(a.dtype == np.float32 for a in nps) else np.float64
    nps = np.broadcast_arrays(*[a.astype(dt) for a in nps])
    shape = nps[0].shape
    return [torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).cuda() for a in nps], shape, False, scalar


def _elementwise(op, arrs, what):
    torch = require_cuda()
    ts, shape, was_torch, scalar = _prep(arrs)
    n = ts[0].shape[0]
    is64 = ts[0].dtype == torch.float64
    out_dt = ts[0].dtype if op == _MAGNITUDE else torch.float64
    out = torch.empty(n, dtype=out_dt, device=ts[0].device)
    x, y, z = (ts if len(ts) == 3 else [None] + ts)
    st = L.load().mhb_accel_elementwise(op, 1 if is64 else 0, x.data_ptr() if x is not None else None, y.data_ptr(),
                                        z.data_ptr(), n, out.data_ptr(), _stream_ptr(torch))
    L.check(st, what)
    out = out.reshape(shape)
    if was_torch:
        return out
    res = out.cpu().numpy()
    return res.item() if scalar else res


@singledispatch
def roll(y, z):
    """Angular roll (degrees) from gravitational acceleration: arctan2(y, z) * 180 / pi (accelerometer.py:13-25)."""
    return _elementwise(_ROLL, [y, z], "roll")


@roll.register(pd.DataFrame)
def _df_roll(df, ycol: str = 'y', zcol: str = 'z'):
    return pd.Series(roll(df[ycol].values, df[zcol].values), name='roll')


@singledispatch
def pitch(x, y, z):
    """Angular pitch (degrees): arctan2(-x, sqrt(y*y + z*z)) * 180 / pi (accelerometer.py:44-56)."""
    return _elementwise(_PITCH, [x, y, z], "pitch")


@pitch.register(pd.DataFrame)
def _df_pitch(df, xcol: str = 'x', ycol: str = 'y', zcol: str = 'z'):
    return pd.Series(pitch(df[xcol].values, df[ycol].values, df[zcol].values), name='pitch')


@singledispatch
def magnitude(x, y, z):
    """Magnitude of acceleration sqrt(x**2 + y**2 + z**2), elementwise (accelerometer.py:198-225)."""
    return _elementwise(_MAGNITUDE, [x, y, z], "magnitude")


@magnitude.register(pd.DataFrame)
def _pd_magnitude(df, xcol: str = 'x', ycol: str = 'y', zcol: str = 'z'):
    return pd.Series(magnitude(df[xcol].values, df[ycol].values, df[zcol].values), name='magnitude')


@singledispatch
def magnitude_dot(x, y, z):
    """sqrt(x.x + y.y + z.z): one scalar for three arrays (accelerometer.py:236-259).  The sums are
    accumulated in float64 in a fixed order (the reference's BLAS dot accumulates float32 input in float32)."""
    torch = require_cuda()
    ts, _, _, _ = _prep([x, y, z])
    n = ts[0].shape[0]
    lib = L.load()
    wlen = int(lib.mhb_accel_sumsq_workspace(n))
    ws = torch.empty(wlen, dtype=torch.float64, device=ts[0].device)
    out = torch.empty(1, dtype=torch.float64, device=ts[0].device)
    st = lib.mhb_accel_magnitude_dot(1 if ts[0].dtype == torch.float64 else 0, ts[0].data_ptr(), ts[1].data_ptr(),
                                     ts[2].data_ptr(), n, ws.data_ptr(), wlen, out.data_ptr(), _stream_ptr(torch))
    L.check(st, "magnitude_dot")
    return float(out.item())


@magnitude_dot.register(pd.DataFrame)
def _pd_magnitude_dot(df, xcol: str = 'x', ycol: str = 'y', zcol: str = 'z'):
    return magnitude_dot(df[xcol].values, df[ycol].values, df[zcol].values)


def rolling_magnitude(funcs, wsize=None, wstep: int = 1):
    """``rolling_apply(funcs, wsize, wstep)(magnitude(x, y, z))`` in one step: returns a callable
    ``(x, y, z, wsize=wsize, wstep=wstep)`` with rolling_apply's output forms (one reducer -> array, list / tuple ->
    list, dict -> dict).  Streaming reducers (mean .. line_length) never materialise the magnitude series: the axes
    are combined while kernel 1a stages its tile; results are bit-identical to the two-step form."""
    from ..engine import magnitude_window_table
    from ..reducers import resolve
    from ..util.windows import _zc_threshold
    if isinstance(funcs, dict):
        names, flist, form = list(funcs.keys()), list(funcs.values()), "dict"
    elif isinstance(funcs, (list, tuple)):
        names, flist, form = None, list(funcs), "list"
    else:
        names, flist, form = None, [funcs], "one"
    features = [resolve(f)[0] for f in flist]
    fss = {f.fs for f in features if f.family == "spectral"}
    if len(fss) > 1:
        raise NotImplementedError("one sampling rate per rolling_magnitude call")
    fs = fss.pop() if fss else 1.0

    def magnitude_windows(x, y, z, wsize=wsize, wstep=wstep):
        if wsize is None:
            raise TypeError("wsize must be given (at rolling_magnitude() or at call time)")
        arrs = [np.asarray(a) for a in (x, y, z)]
        if any(a.ndim != 1 for a in arrs):
            raise ValueError("rolling_magnitude takes three 1-D arrays")
        if fss or features and any(f.family == "spectral" for f in features):
            arrs = [a.astype(np.float32) for a in arrs]      # the spectral kernel computes in float32
        tab = magnitude_window_table(arrs[0], arrs[1], arrs[2], int(wsize), int(wstep), features,
                                     zc_threshold=_zc_threshold(features), fs=fs)
        cols = [np.ascontiguousarray(tab[:, j]) for j in range(len(features))]
        if form == "one":
            return cols[0]
        return dict(zip(names, cols)) if form == "dict" else cols
    return magnitude_windows
