"""Host-side engine: feature specs -> C-ABI calls on torch-owned device buffers.

PyTorch is used for device memory, streams and (in ``sharded``) torch.distributed only; every
number is produced by the CUDA kernels in libmhb200.so.  numpy in -> numpy out for the drop-in
modules (``util.windows`` ...), torch CUDA tensors in -> torch CUDA tensors out for resident data.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L


class Feature:
    """One output column: family 'stream' | 'order' | 'spectral', C-ABI feature id, parameters."""
    __slots__ = ("family", "fid", "params", "name", "fs")

    def __init__(self, family, fid, params=(), name=None):
        self.fs = None
        self.family = family
        self.fid = int(fid)
        self.params = tuple(float(p) if p is not None else math.nan for p in params)
        self.name = name or "f%d" % fid

    def key(self):
        return (self.family, self.fid, self.params)

    def __repr__(self):
        return "Feature(%s%s)" % (self.name, self.params if self.params else "")


def _torch():
    import torch
    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise L.MhbError("pymhealth_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    L.load()
    return torch


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device_series(x):
    """numpy / torch, 1-D or 2-D -> (cuda tensor [n_series, len] float32|float64 contiguous rows,
    was_numpy, was_1d)."""
    torch = require_cuda()
    was_numpy = not isinstance(x, torch.Tensor)
    if was_numpy:
        a = np.asarray(x)
        if a.dtype == np.float32 or a.dtype == np.float64:
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda(non_blocking=False)
        elif a.dtype in (np.int8, np.uint8, np.int16, np.uint16, np.float16, np.bool_):
            # raw sensor counts: upload as they are (2 bytes per sample over PCIe) and widen on the device; every value
            # is exact in float32, and the kernels accumulate in float64 either way
            raw = a.view(np.uint8) if a.dtype == np.bool_ else a
            if raw.dtype == np.uint16:
                raw = raw.astype(np.int32)                 # torch has no uint16 arithmetic
            t = torch.from_numpy(np.ascontiguousarray(raw)).cuda(non_blocking=False).float()
        elif a.dtype.kind in "iu":
            raw = a.astype(np.int64) if a.dtype in (np.uint32, np.uint64) else a
            t = torch.from_numpy(np.ascontiguousarray(raw)).cuda(non_blocking=False).double()   # exact for 32-bit integers
        else:
            raise TypeError("unsupported dtype %s" % a.dtype)
    else:
        t = x
        if not t.is_cuda:
            t = t.cuda()
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
    was_1d = t.dim() == 1
    if was_1d:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError("series must be 1-D or 2-D [n_series, len]")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t, was_numpy, was_1d


def n_windows(n, wsize, wstep):
    return max(0, 1 + (n - wsize) // wstep) if n >= wsize else 0


def window_table(x, wsize, wstep, features, zc_threshold=0.0, fs=1.0, out_dtype=None, out=None):
    """Feature table of every window of every series.

    x: numpy or torch, [len] or [n_series, len].  features: list of Feature.
    Returns [n_series, nw, len(features)] (numpy float64 for numpy input unless out_dtype is given;
    a CUDA tensor for CUDA input).  All streaming features share one kernel pass, all order
    features another, all spectral features a third.
    """
    torch = require_cuda()
    lib = L.load()
    wsize = int(wsize)
    wstep = int(wstep)
    if wsize < 1 or wstep < 1:
        raise ValueError("wsize and wstep must be >= 1")
    t, was_numpy, was_1d = to_device_series(x)
    ns, n = t.shape
    nw = n_windows(n, wsize, wstep)
    nf = len(features)
    if out_dtype is None:
        out_dtype = torch.float64 if was_numpy else torch.float32
    if out is None:
        out = torch.empty((ns, nw, nf), dtype=out_dtype, device=t.device)
    elif tuple(out.shape) != (ns, nw, nf):
        raise ValueError("out has shape %s, expected %s" % (tuple(out.shape), (ns, nw, nf)))
    _check_out(torch, out, t)
    if nw > 0 and ns > 0 and nf > 0 and t.device.index != torch.cuda.current_device():
        with torch.cuda.device(t.device):          # launch on the device (and its current stream) that holds the series
            res = window_table(t, wsize, wstep, features, zc_threshold=zc_threshold, fs=fs, out_dtype=out_dtype, out=out)
        if was_numpy:
            res = res.cpu().numpy()
        return res[0] if was_1d else res
    if nw > 0 and ns > 0 and nf > 0:
        geom = L.MhbWindows(ns, n, t.stride(0) if ns > 1 else n, wsize, wstep)
        stream = _stream_ptr(torch)
        f32_in = t.dtype == torch.float32
        by_family = {"stream": [], "order": [], "spectral": []}
        for j, f in enumerate(features):
            by_family[f.family].append(j)
        # statistical + spectral columns of the same float32 windows: ONE C-ABI call (kernel 1a and kernel 2 back to
        # back on the stream); the order family follows on its own
        s_runs, p_runs = _runs(by_family["stream"]), _runs(by_family["spectral"])
        if f32_in and len(s_runs) == 1 and len(p_runs) == 1:
            sr, pr = s_runs[0], p_runs[0]
            esz = out.stride(2) * out.element_size()
            f32o = 1 if out.dtype == torch.float32 else 0
            stab = L.MhbTable(out.data_ptr() + sr[0] * esz, f32o, out.stride(0), out.stride(1), out.stride(2))
            ptab = L.MhbTable(out.data_ptr() + pr[0] * esz, f32o, out.stride(0), out.stride(1), out.stride(2))
            flat = []
            for j in pr:
                p = features[j].params
                flat += [p[0] if len(p) > 0 else math.nan, p[1] if len(p) > 1 else math.nan]
            st = lib.mhb_window_features_f32(t.data_ptr(), C.byref(geom), L.i32_array([features[j].fid for j in sr]), len(sr),
                                             float(zc_threshold), C.byref(stab), float(fs),
                                             L.i32_array([features[j].fid for j in pr]), L.f64_array(flat), len(pr),
                                             C.byref(ptab), stream)
            L.check(st, "window_features")
            by_family = {"order": by_family["order"]}
        # columns of one family must be contiguous-strided for the table descriptor: launch each
        # family on maximal runs of consecutive columns
        for family, cols in by_family.items():
            for run in _runs(cols):
                j0 = run[0]
                tab = L.MhbTable(out.data_ptr() + j0 * out.stride(2) * out.element_size(), 1 if out.dtype == torch.float32 else 0,
                                 out.stride(0), out.stride(1), out.stride(2))
                ids = L.i32_array([features[j].fid for j in run])
                if family == "stream":
                    fn = lib.mhb_window_stats_f32 if f32_in else lib.mhb_window_stats_f64
                    st = fn(t.data_ptr(), C.byref(geom), ids, len(run), float(zc_threshold), C.byref(tab), stream)
                elif family == "order":
                    fn = lib.mhb_window_order_f32 if f32_in else lib.mhb_window_order_f64
                    pars = L.f64_array([features[j].params[0] if features[j].params else 0.0 for j in run])
                    st = fn(t.data_ptr(), C.byref(geom), ids, pars, len(run), C.byref(tab), stream)
                else:
                    if not f32_in:
                        raise NotImplementedError("spectral features take float32 series (convert with .astype(np.float32))")
                    flat = []
                    for j in run:
                        p = features[j].params
                        flat += [p[0] if len(p) > 0 else math.nan, p[1] if len(p) > 1 else math.nan]
                    st = lib.mhb_window_spectral_f32(t.data_ptr(), C.byref(geom), float(fs), ids, L.f64_array(flat),
                                                     len(run), C.byref(tab), stream)
                L.check(st, "window_%s" % family)
    if was_numpy:
        res = out.cpu().numpy()
        return res[0] if was_1d else res
    return out[0] if was_1d else out


def _check_out(torch, out, t):
    """A caller-supplied table must be float32 / float64 on the input's device (any element strides are fine)."""
    if out.dtype not in (torch.float32, torch.float64):
        raise TypeError("out must be float32 or float64, not %s" % out.dtype)
    if out.device != t.device:
        raise ValueError("out is on %s, the series on %s" % (out.device, t.device))


def magnitude_window_table(x, y, z, wsize, wstep, features, zc_threshold=0.0, fs=1.0, out_dtype=None, out=None):
    """Feature table of the windows of ``magnitude(x, y, z)`` (inertial/accelerometer.py:198-225 followed by
    ``rolling_apply``) for tri-axial series of one geometry ([len] or [n_series, len] each).

    The streaming family (kernel 1a) never materialises the magnitude: the axes are combined while a stage is copied
    into shared memory (``mhb_window_stats_magnitude_*``).  Order-statistic and spectral columns need the series more
    than once, so for them it is written to device memory once (``mhb_accel_elementwise``) and the ordinary kernels run
    on it.  Results are bit-identical to ``window_table(magnitude(x, y, z), ...)``."""
    torch = require_cuda()
    lib = L.load()
    wsize, wstep = int(wsize), int(wstep)
    if wsize < 1 or wstep < 1:
        raise ValueError("wsize and wstep must be >= 1")
    tx, was_numpy, was_1d = to_device_series(x)
    ty, _, _ = to_device_series(y)
    tz, _, _ = to_device_series(z)
    if not (tx.shape == ty.shape == tz.shape):
        raise ValueError("x, y and z must have the same shape")
    dt = torch.float32 if tx.dtype == ty.dtype == tz.dtype == torch.float32 else torch.float64
    tx, ty, tz = (t.to(dt) for t in (tx, ty, tz))
    if not (tx.stride() == ty.stride() == tz.stride()):
        tx, ty, tz = (t.contiguous() for t in (tx, ty, tz))
    ns, n = tx.shape
    row_stride = tx.stride(0) if ns > 1 else n       # e.g. the three axis planes of one [subjects, 3, len] tensor
    nw = n_windows(n, wsize, wstep)
    nf = len(features)
    if out_dtype is None:
        out_dtype = torch.float64 if was_numpy else torch.float32
    if out is None:
        out = torch.empty((ns, nw, nf), dtype=out_dtype, device=tx.device)
    elif tuple(out.shape) != (ns, nw, nf):
        raise ValueError("out has shape %s, expected %s" % (tuple(out.shape), (ns, nw, nf)))
    _check_out(torch, out, tx)
    if nw > 0 and ns > 0 and nf > 0:
        stream = _stream_ptr(torch)
        s_cols = [j for j, f in enumerate(features) if f.family == "stream"]
        o_cols = [j for j, f in enumerate(features) if f.family != "stream"]
        geom = L.MhbWindows(ns, n, row_stride, wsize, wstep)
        fn = lib.mhb_window_stats_magnitude_f32 if dt == torch.float32 else lib.mhb_window_stats_magnitude_f64
        for run in _runs(s_cols):
            tab = L.MhbTable(out.data_ptr() + run[0] * out.stride(2) * out.element_size(), 1 if out.dtype == torch.float32 else 0,
                             out.stride(0), out.stride(1), out.stride(2))
            ids = L.i32_array([features[j].fid for j in run])
            L.check(fn(tx.data_ptr(), ty.data_ptr(), tz.data_ptr(), C.byref(geom), ids, len(run), float(zc_threshold),
                       C.byref(tab), stream), "window_stats_magnitude")
        if o_cols:
            cx, cy, cz = (t.contiguous() for t in (tx, ty, tz))
            mag = torch.empty_like(cx)
            L.check(lib.mhb_accel_elementwise(0, 1 if dt == torch.float64 else 0, cx.data_ptr(), cy.data_ptr(), cz.data_ptr(),
                                              ns * n, mag.data_ptr(), stream), "magnitude")
            for run in _runs(o_cols):
                window_table(mag, wsize, wstep, [features[j] for j in run], zc_threshold=zc_threshold, fs=fs,
                             out=out[:, :, run[0]:run[-1] + 1])
    if was_numpy:
        res = out.cpu().numpy()
        return res[0] if was_1d else res
    return out[0] if was_1d else out


def n_index_windows(first, last, wstep, is_float):
    """len(np.arange(first, last, wstep)) without building the array (util/windows.py:175): numpy takes
    ceil((stop - start) / step) in float64 for floats and the exact integer count for integers."""
    if is_float:
        if not wstep > 0:
            raise ValueError("get_indices: wstep must be > 0")
        delta = float(last) - float(first)
        q = delta / float(wstep)
        if q == 0.0 and delta != 0.0:          # numpy's underflow rule (PyArray_Arange): a denormal span still holds one key
            return 0 if math.copysign(1.0, q) < 0 else 1
        return max(0, int(math.ceil(q)))
    if int(wstep) <= 0:
        raise ValueError("get_indices: wstep must be > 0")
    return len(range(int(first), int(last), int(wstep)))


def device_get_indices(index, wsize, wstep):
    """get_indices (util/windows.py:162-178) on the device -> int64 cuda tensor [2, n_windows].

    index: numpy / torch 1-D, integer, datetime64 / timedelta64 (same unit as wsize / wstep after conversion) or float.
    """
    torch = require_cuda()
    lib = L.load()
    if isinstance(index, torch.Tensor):
        it = index if index.is_cuda else index.cuda()
        if it.dtype.is_floating_point:
            it = it.double()
            first, last = float(it[0]), float(it[-1])
        else:
            it = it.long()
            first, last = int(it[0]), int(it[-1])
    else:
        a = np.asarray(index)
        if a.ndim != 1 or a.shape[0] == 0:
            raise ValueError("get_indices: index must be a non-empty 1-D array")
        if a.dtype.kind in "mM":
            unit = np.datetime_data(a.dtype)[0]
            if isinstance(wsize, np.timedelta64):
                wsize = int(wsize.astype("timedelta64[%s]" % unit).astype(np.int64))
            if isinstance(wstep, np.timedelta64):
                wstep = int(wstep.astype("timedelta64[%s]" % unit).astype(np.int64))
            a = a.view(np.int64)
        if a.dtype.kind == "f":
            a = a.astype(np.float64)
            first, last = float(a[0]), float(a[-1])
        elif a.dtype.kind in "iu":
            a = a.astype(np.int64)
            first, last = int(a[0]), int(a[-1])
        else:
            raise TypeError("get_indices: unsupported index dtype %s" % a.dtype)
        it = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    it = it.contiguous()
    n = it.shape[0]
    is_f = it.dtype == torch.float64
    if is_f:
        wsize, wstep = float(wsize), float(wstep)
    else:
        wsize, wstep = int(wsize), int(wstep)
    nwin = n_index_windows(first, last, wstep, is_f)
    out = torch.empty((2, nwin), dtype=torch.int64, device=it.device)
    if nwin:
        fn = lib.mhb_get_indices_f64 if is_f else lib.mhb_get_indices_i64
        # (float keys: the kernel reproduces numpy's arange fill, start + i * ((start + step) - start), itself)
        st = fn(it.data_ptr(), n, first, wsize, wstep, nwin, out.data_ptr(), _stream_ptr(torch))
        L.check(st, "get_indices")
    return out


def segment_table(x, indices, features, min_window_len=1, zc_threshold=0.0, out_dtype=None):
    """Feature table of arbitrary [start, end) windows of ONE series (indices_rolling_apply,
    util/windows.py:122-159).  indices: int64 [2, n] (numpy or cuda tensor).  Returns a cuda tensor
    [n, len(features)]; rows of windows shorter than min_window_len are NaN."""
    torch = require_cuda()
    lib = L.load()
    t, _, was_1d = to_device_series(x)
    if not was_1d and t.shape[0] != 1:
        raise ValueError("index-addressed windows take a 1-D array")
    t = t.reshape(-1)
    n = t.shape[0]
    if isinstance(indices, torch.Tensor):
        ind = indices.to(device=t.device, dtype=torch.int64)
    else:
        ia = np.asarray(indices)
        if ia.ndim != 2 or ia.shape[0] != 2:
            raise ValueError("indices must have shape (2, n)")
        ind = torch.from_numpy(np.ascontiguousarray(ia.astype(np.int64))).to(t.device)
    if ind.dim() != 2 or ind.shape[0] != 2:
        raise ValueError("indices must have shape (2, n)")
    ind = ind.contiguous()
    nwin = ind.shape[1]
    nf = len(features)
    if out_dtype is None:
        out_dtype = torch.float64
    out = torch.empty((nwin, nf), dtype=out_dtype, device=t.device)
    if nwin == 0 or nf == 0:
        return out
    starts_ptr = ind.data_ptr()
    ends_ptr = ind.data_ptr() + nwin * 8
    stream = _stream_ptr(torch)
    f32_in = t.dtype == torch.float32
    by_family = {"stream": [], "order": []}
    for j, f in enumerate(features):
        if f.family not in by_family:
            raise NotImplementedError("spectral reducers are defined on uniform windows only")
        by_family[f.family].append(j)
    max_len = None
    for family, cols in by_family.items():
        for run in _runs(cols):
            j0 = run[0]
            tab = L.MhbTable(out.data_ptr() + j0 * out.stride(1) * out.element_size(), 1 if out.dtype == torch.float32 else 0,
                             0, out.stride(0), out.stride(1))
            ids = L.i32_array([features[j].fid for j in run])
            if family == "stream":
                fn = lib.mhb_segment_stats_f32 if f32_in else lib.mhb_segment_stats_f64
                st = fn(t.data_ptr(), n, starts_ptr, ends_ptr, nwin, int(min_window_len), ids, len(run),
                        float(zc_threshold), C.byref(tab), stream)
            else:
                if max_len is None:
                    lens = ind[1].clamp(max=n) - ind[0].clamp(min=0)
                    max_len = max(0, int(lens.max().item()))
                fn = lib.mhb_segment_order_f32 if f32_in else lib.mhb_segment_order_f64
                pars = L.f64_array([features[j].params[0] if features[j].params else 0.0 for j in run])
                st = fn(t.data_ptr(), n, starts_ptr, ends_ptr, nwin, max_len, int(min_window_len), ids, pars,
                        len(run), C.byref(tab), stream)
            L.check(st, "segment_%s" % family)
    return out


def _runs(cols):
    runs, cur = [], []
    for c in cols:
        if cur and c != cur[-1] + 1:
            runs.append(cur)
            cur = []
        cur.append(c)
    if cur:
        runs.append(cur)
    return runs
