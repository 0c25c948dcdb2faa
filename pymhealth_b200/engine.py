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
            pass
        elif a.dtype.kind in "iub" or a.dtype == np.float16:
            a = a.astype(np.float64)          # exact for <= 32-bit integers
        else:
            raise TypeError("unsupported dtype %s" % a.dtype)
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda(non_blocking=False)
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
    if nw > 0 and ns > 0 and nf > 0:
        geom = L.MhbWindows(ns, n, t.stride(0) if ns > 1 else n, wsize, wstep)
        stream = _stream_ptr(torch)
        f32_in = t.dtype == torch.float32
        by_family = {"stream": [], "order": [], "spectral": []}
        for j, f in enumerate(features):
            by_family[f.family].append(j)
        # columns of one family must be contiguous-strided for the table descriptor: launch each
        # family on maximal runs of consecutive columns
        for family, cols in by_family.items():
            for run in _runs(cols):
                j0 = run[0]
                tab = L.MhbTable(out.data_ptr() + j0 * out.element_size(), 1 if out.dtype == torch.float32 else 0,
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
