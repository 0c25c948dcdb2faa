"""Spectral chain of the reference, restated.  TEST INFRASTRUCTURE ONLY.

The reference has no single windowed-spectrum entry point; users compose
``view`` -> ``mhealth.fft.fft`` -> ``|F|^2`` -> PSD reducers.  ``mhealth.fft.fft`` is FFTW through
a CFFI shim when that shim is compiled and ``numpy.fft.fft`` otherwise
(``src/mhealth/fft/__init__.py:3-7``); FFTW is absent from this image, so the oracle is the
documented fallback: numpy's pocketfft in float64, length == wsize exactly, unnormalised.

FFT third-party pin: numpy 2.3.5 pocketfft ("parity unpinned" against FFTW -- both are
float64 DFTs accurate to ~1e-15, far inside the 1e-5 contract).
"""
import numpy as np

from . import reducers as R
from .windows import view


def fft(a):
    """fft/_fft.py:18-29 / numpy fallback: unnormalised forward DFT, complex128."""
    return np.fft.fft(np.asarray(a, dtype=np.complex128))


def ifft(a):
    """fft/_fft.py:46-48: backward DFT divided by n."""
    return np.fft.ifft(np.asarray(a, dtype=np.complex128))


def window_psd(x, wsize, wstep, fs):
    """|FFT|^2 of every window, one-sided layout (bins 0..wsize//2) and its frequencies.

    Follows SURVEY 8c quick-start: ``np.abs(np.fft.fft(view(x, W, S), axis=1)[:, :W//2+1])**2``
    with ``freqs = np.fft.rfftfreq(W, 1/fs)``.  Returns (psd float64[nw, nb], freqs float64[nb]).
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    nb = wsize // 2 + 1
    freqs = np.fft.rfftfreq(wsize, 1.0 / fs)
    if x.shape[0] < wsize:
        return np.zeros((0, nb)), freqs
    spec = np.fft.fft(view(x, wsize, wstep), axis=1)[:, :nb]
    return np.abs(spec) ** 2, freqs


def power_band(psd, freqs, lower=None, upper=None):
    """heart/hrv.py:173-179: sum |psd| over lower <= f <= upper (BOTH ends inclusive)."""
    if lower is None:
        lower = np.min(freqs)
    if upper is None:
        upper = np.max(freqs)
    keep = np.logical_and(freqs >= lower, freqs <= upper)
    return np.sum(np.abs(psd[keep]))


def relative_power_band(psd, freqs, lower=None, upper=None):
    """heart/hrv.py:192-198: band power / total |psd|."""
    return power_band(psd, freqs, lower, upper) / np.sum(np.abs(psd))


def first_index(arr, x):
    """generic/frequency/density.py:9-14: first i with x <= arr[i], else len(arr)."""
    for i in range(len(arr)):
        if x <= arr[i]:
            return i
    return len(arr)


def peak_bin(psd, freqs, lower=None, upper=None):
    """Integer bin behind density.peak_frequency (density.py:18-32): lower bound inclusive,
    upper bound EXCLUSIVE, first maximum wins."""
    lidx = 0 if lower is None else first_index(freqs, lower)
    uidx = len(psd) if upper is None else first_index(freqs, upper)
    return lidx + int(np.argmax(psd[lidx:uidx]))


def peak_frequency(psd, freqs, lower=None, upper=None):
    """generic/frequency/density.py:18-32.  (heart/hrv.py:182-189 has a mask/index bug and
    is deliberately NOT the oracle -- SURVEY 8a.)"""
    return freqs[peak_bin(psd, freqs, lower, upper)]


def spectral_entropy(psd):
    """``information.entropy(psd)`` (generic/information.py:10-20) applied to one PSD row."""
    return R.entropy(np.ascontiguousarray(psd, dtype=np.float64))


def spectral_table(x, wsize, wstep, fs, bands, peak_lo=None, peak_hi=None):
    """Per-window spectral feature table used by the parity tests and the CPU baseline.

    Columns: total power, power_band for each (lo, hi) in ``bands``, relative power for each
    band, peak frequency, peak bin (integer), spectral entropy.
    Returns dict name -> float64[nw] (peak_bin is int64).
    """
    psd, freqs = window_psd(x, wsize, wstep, fs)
    nw = psd.shape[0]
    out = {"total_power": np.abs(psd).sum(axis=1)}
    for j, (lo, hi) in enumerate(bands):
        keep = np.logical_and(freqs >= lo, freqs <= hi)
        bp = np.abs(psd[:, keep]).sum(axis=1)
        out["band_power_%d" % j] = bp
        out["rel_band_power_%d" % j] = bp / out["total_power"]
    lidx = 0 if peak_lo is None else first_index(freqs, peak_lo)
    uidx = psd.shape[1] if peak_hi is None else first_index(freqs, peak_hi)
    pb = lidx + np.argmax(psd[:, lidx:uidx], axis=1) if nw else np.zeros(0, dtype=np.int64)
    out["peak_bin"] = pb.astype(np.int64)
    out["peak_frequency"] = freqs[pb] if nw else np.zeros(0)
    p = psd / psd.sum(axis=1, keepdims=True) + 1e-30
    out["spectral_entropy"] = -(p * np.log(p)).sum(axis=1)
    return out
