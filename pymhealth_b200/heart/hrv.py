"""Frequency-domain PSD reducers of ``mhealth.heart.hrv`` (reference src/mhealth/heart/hrv.py:172-198).
and, below them, its RR-interval time-domain metrics (hrv.py:24-170) as thin wrappers over the same kernels."""
import numpy as np

from .. import _lib as L
from ..spectral import psd_reduce


def power_band(psd, freqs, lower=None, upper=None):
    """sum |psd| over lower <= f <= upper, both inclusive (hrv.py:173-179)."""
    return float(psd_reduce(psd, freqs, [(L.S_BAND_POWER, lower, upper)])[0])


def relative_power_band(psd, freqs, lower=None, upper=None):
    """power_band / sum |psd| (hrv.py:192-198)."""
    return float(psd_reduce(psd, freqs, [(L.S_REL_BAND_POWER, lower, upper)])[0])


def peak_frequency(psd, freqs, lower=None, upper=None):
    """The reference's hrv.peak_frequency (hrv.py:182-189) indexes the UNMASKED frequency vector with
    the argmax of the MASKED psd and is wrong whenever lower > min(freqs).  This function
    implements what it plainly intends -- the frequency of the largest PSD value inside
    lower <= f <= upper -- and says so; use generic.frequency.density.peak_frequency for the
    reference's other (correct, upper-exclusive) variant."""
    f = np.asarray(freqs, dtype=np.float64)
    hi = None
    if upper is not None:
        above = f[f > upper]
        hi = float(above.min()) if above.size else None     # inclusive upper bound -> exclusive bound at the next bin
    return float(psd_reduce(psd, freqs, [(L.S_PEAK_FREQUENCY, lower, hi)])[0])


# ---------------------------------------------------------------------------------------------------------
# Time-domain metrics on normal R-peak intervals (hrv.py:24-170): thin host wrappers over the window-statistics and
# successive-difference kernels.
def td_factor(unit: str) -> float:
    """Nanoseconds per ``unit`` (hrv.py:24-34)."""
    try:
        return {"ns": 1., "us": 1e3, "ms": 1e6, "s": 1e9}[unit]
    except KeyError:
        raise ValueError('Unknown unit. Must be: "ns", "us", "ms", or "s"')


def nni_to_ms(nni, current_unit: str = 'ns'):
    """hrv.py:37-39."""
    return td_factor(current_unit) * np.asarray(nni).astype(float) / 1e6


def _diff_stats(nni, threshold=0.0):
    from ..engine import require_cuda, _stream_ptr
    torch = require_cuda()
    a = np.ascontiguousarray(np.asarray(nni, dtype=np.float64).ravel())
    if a.shape[0] < 2:
        raise ValueError("at least two intervals are needed")
    lib = L.load()
    d = torch.from_numpy(a).cuda()
    wlen = int(lib.mhb_diff_stats_workspace(a.shape[0] - 1))
    ws = torch.empty(wlen, dtype=torch.float64, device=d.device)
    out = torch.empty(7, dtype=torch.float64, device=d.device)
    L.check(lib.mhb_diff_stats_f64(d.data_ptr(), a.shape[0], float(threshold), ws.data_ptr(), wlen, out.data_ptr(),
                                   _stream_ptr(torch)), "diff_stats")
    return out.cpu().numpy()


def sdnn(nni) -> float:
    """Standard deviation of the intervals, np.std(nni) (hrv.py:50-63)."""
    from ..generic import stats
    return float(stats.std(np.asarray(nni)))


def pnnx(nni, unit: str = 'ms', x: float = 50.) -> float:
    """Proportion of successive differences over ``x`` ms (hrv.py:124-136)."""
    thr = x * 1e6 / td_factor(unit)
    r = _diff_stats(nni, thr)
    return float(r[4] / r[0])


def pnn50(nni, unit: str = 'ms') -> float:
    """Proportion of successive differences over 50 ms (hrv.py:111-121)."""
    return pnnx(nni, unit, 50.)


def rmssd(nni) -> float:
    """Root mean square of successive differences (hrv.py:139-147)."""
    return float(np.sqrt(_diff_stats(nni)[5]))


def ssd(nni) -> float:
    """Sum of successive differences (hrv.py:150-158)."""
    return float(_diff_stats(nni)[1])


def sdsd(nni) -> float:
    """Standard deviation of successive differences (hrv.py:161-170)."""
    return float(np.sqrt(_diff_stats(nni)[3]))


def _segment_index(nni, index, unit, who):
    if index is None:
        if unit is None:
            raise ValueError('%s: index or unit must be specified' % who)
        index = np.cumsum(np.asarray(nni)) * td_factor(unit)          # hrv.py:42-44, 80-82
    idx = np.asarray(index)
    if idx.dtype.kind == "M":
        idx = idx.astype("datetime64[ns]").view(np.int64)
    return idx.astype(np.int64)


def sdann(nni, index=None, interval: float = 60 * 5, unit=None) -> float:
    """Standard deviation of the per-segment means of the intervals (segments of ``interval`` seconds over a
    nanosecond index; hrv.py:66-86).  The reference's version does not compile under current numba (it calls a Python
    closure from nopython code); this is the computation it states, on the non-uniform window kernels."""
    from ..util.windows import nonuniform_rolling_apply
    from ..generic import stats
    idx = _segment_index(nni, index, unit, "sdann")
    step = int(interval * 1e9)
    means = nonuniform_rolling_apply(np.mean)(idx, np.asarray(nni, dtype=np.float64), step, step)
    return float(stats.std(means))


def sdnni(nni, index=None, interval: float = 60 * 5, unit=None) -> float:
    """Mean of the per-segment standard deviations (hrv.py:89-108); see ``sdann``."""
    from ..util.windows import nonuniform_rolling_apply
    from ..generic import stats
    idx = _segment_index(nni, index, unit, "sdnni")
    step = int(interval * 1e9)
    sds = nonuniform_rolling_apply(np.std)(idx, np.asarray(nni, dtype=np.float64), step, step)
    return float(stats.mean(sds))


# Non-linear (Poincare / Lorenz plot) indices, hrv.py:207-266
def csi_sd1(rri, factor: float = 1 / np.sqrt(2)) -> float:
    """Poincare plot width: factor * std(diff(rri)) (hrv.py:207-216)."""
    return float(factor * np.sqrt(_diff_stats(rri)[3]))


def csi_sd2(rri, factor: float = 1 / np.sqrt(2)) -> float:
    """Poincare plot length: factor * std(rri[1:] + rri[:-1]) (hrv.py:219-231)."""
    return float(factor * np.sqrt(_diff_stats(rri)[6]))


def _sd12(rri, factor):
    r = _diff_stats(rri)
    return factor * np.sqrt(r[3]), factor * np.sqrt(r[6])


def lorenz_csi(rri, factor: float = 1 / np.sqrt(2)) -> float:
    """Cardiac sympathetic index sd1 / sd2 (hrv.py:234-243)."""
    a, b = _sd12(rri, factor)
    return float(a / b)


def lorenz_cvi(rri, factor: float = 1 / np.sqrt(2)) -> float:
    """log10(sd1 * sd2) (hrv.py:246-250)."""
    a, b = _sd12(rri, factor)
    return float(np.log10(a * b))


def lorenz_mcsi(rri, factor: float = 1 / np.sqrt(2)) -> float:
    """Modified sympathetic index sd1^2 / sd2 (hrv.py:253-266)."""
    a, b = _sd12(rri, factor)
    return float(a ** 2 / b)
