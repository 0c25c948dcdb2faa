"""Windowed spectral features -- the GPU form of the chain the reference leaves to the user
(SURVEY 3.3): ``view`` -> ``mhealth.fft.fft`` -> ``|F|^2`` -> ``hrv.power_band`` /
``density.peak_frequency`` / ``information.entropy``.

``band_power(fs, lo, hi)`` & co. return reducer objects that ``rolling_apply`` accepts next to
the statistical reducers; ``window_psd`` returns the raw one-sided PSD rows; ``psd_reduce`` is the
engine behind the literal ``hrv.power_band(psd, freqs, lo, hi)``-style calls.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from .engine import Feature, require_cuda, to_device_series, n_windows, _stream_ptr


class SpectralReducer:
    """A PSD column bound to a sampling rate; accepted by ``rolling_apply`` (one fs per call)."""

    def __init__(self, name, fid, fs, lo=None, hi=None, integer=False):
        self.__name__ = name
        self.fid = fid
        self.fs = float(fs)
        self.lo = lo
        self.hi = hi
        self.integer = integer

    def feature(self):
        f = Feature("spectral", self.fid, (self.lo, self.hi), self.__name__)
        f.fs = self.fs
        return f

    def __repr__(self):
        return "<pymhealth_b200 spectral reducer %s fs=%g [%s, %s]>" % (self.__name__, self.fs, self.lo, self.hi)


def total_power(fs):
    return SpectralReducer("total_power", L.S_TOTAL_POWER, fs)


def band_power(fs, lower=None, upper=None):
    """hrv.power_band on each window's PSD: lower <= f <= upper, both inclusive."""
    return SpectralReducer("band_power", L.S_BAND_POWER, fs, lower, upper)


def relative_band_power(fs, lower=None, upper=None):
    return SpectralReducer("relative_band_power", L.S_REL_BAND_POWER, fs, lower, upper)


def peak_frequency(fs, lower=None, upper=None):
    """density.peak_frequency on each window's PSD: lower <= f < upper, first maximum."""
    return SpectralReducer("peak_frequency", L.S_PEAK_FREQUENCY, fs, lower, upper)


def peak_bin(fs, lower=None, upper=None):
    return SpectralReducer("peak_bin", L.S_PEAK_BIN, fs, lower, upper, integer=True)


def spectral_entropy(fs):
    """information.entropy of each window's PSD."""
    return SpectralReducer("spectral_entropy", L.S_ENTROPY, fs)


def window_psd(x, wsize, wstep, fs=1.0, out_dtype=None):
    """One-sided |FFT|^2 of every window: (psd [n_series?, nw, W//2+1], freqs [W//2+1])."""
    torch = require_cuda()
    lib = L.load()
    t, was_numpy, was_1d = to_device_series(x)
    if t.dtype != torch.float32:
        t = t.float()
    ns, n = t.shape
    nw = n_windows(n, wsize, wstep)
    nb = wsize // 2 + 1
    if out_dtype is None:
        out_dtype = torch.float64 if was_numpy else torch.float32
    out = torch.empty((ns, nw, nb), dtype=out_dtype, device=t.device)
    if nw > 0:
        if ns > 1 and t.stride(0) != n:
            t = t.contiguous()
        geom = L.MhbWindows(ns, n, n, int(wsize), int(wstep))
        L.check(lib.mhb_window_psd_f32(t.data_ptr(), C.byref(geom), out.data_ptr(),
                                       1 if out_dtype == torch.float32 else 0, _stream_ptr(torch)), "window_psd")
    freqs = np.fft.rfftfreq(int(wsize), 1.0 / float(fs))
    if was_numpy:
        res = out.cpu().numpy()
        return (res[0] if was_1d else res), freqs
    return (out[0] if was_1d else out), freqs


def psd_reduce(psd, freqs, columns):
    """columns: list of (column id, lo, hi).  psd: [nb] or [n_rows, nb] float64.  -> float64 [n_rows, n_cols]."""
    torch = require_cuda()
    lib = L.load()
    a = np.ascontiguousarray(np.asarray(psd, dtype=np.float64))
    one = a.ndim == 1
    if one:
        a = a[None, :]
    if a.ndim != 2 or a.shape[1] < 1:
        raise ValueError("psd must be [nb] or [n_rows, nb]")
    d = torch.from_numpy(a).cuda()
    fd = None
    if freqs is not None:
        f = np.ascontiguousarray(np.asarray(freqs, dtype=np.float64))
        if f.shape != (a.shape[1],):
            raise ValueError("freqs must have one entry per PSD bin")
        fd = torch.from_numpy(f).cuda()
    ids = L.i32_array([c[0] for c in columns])
    flat = []
    for c in columns:
        flat += [math.nan if c[1] is None else float(c[1]), math.nan if c[2] is None else float(c[2])]
    out = torch.empty((a.shape[0], len(columns)), dtype=torch.float64, device=d.device)
    L.check(lib.mhb_psd_reduce_f64(d.data_ptr(), fd.data_ptr() if fd is not None else None, a.shape[0], a.shape[1],
                                   ids, L.f64_array(flat), len(columns), out.data_ptr(), _stream_ptr(torch)), "psd_reduce")
    res = out.cpu().numpy()
    return res[0] if one else res
