"""Per-window reducers of the reference, restated with explicit loops (numba nopython).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each ``w_*`` function takes ONE window (1-D float64 array) and returns a scalar, exactly
like the callables a user hands to the reference's ``rolling_apply``.  Where the reference
simply aliases numpy (``src/mhealth/generic/stats.py:156-163``) the arithmetic that runs
inside a jitted loop is numba's own implementation; its loop order is restated here so the
oracle does not depend on which numba happens to be installed.
"""
import math

import numpy as np
from numba import njit


# --------------------------------------------------------------------------- moments
@njit(cache=True)
def w_mean(w):
    # stats.py:157 -> numba/np/arraymath.py:434-466: sequential sum in the array dtype, / size
    c = 0.0
    for i in range(w.shape[0]):
        c += w[i]
    return c / w.shape[0]


@njit(cache=True)
def w_var(w):
    # stats.py:160 -> arraymath.py:469-487: two-pass population variance (ddof = 0)
    m = w_mean(w)
    ssd = 0.0
    for i in range(w.shape[0]):
        d = w[i] - m
        ssd += d * d
    return ssd / w.shape[0]


@njit(cache=True)
def w_std(w):
    # stats.py:159 -> arraymath.py:490-496: var ** 0.5
    return w_var(w) ** 0.5


@njit(cache=True)
def w_min(w):
    # stats.py:161 (np.min): numba's array min propagates NaN (numba/np/arraymath.py, "if np.isnan(v): return v")
    m = w[0]
    if np.isnan(m):
        return m
    for i in range(1, w.shape[0]):
        v = w[i]
        if np.isnan(v):
            return v
        if v < m:
            m = v
    return m


@njit(cache=True)
def w_max(w):
    # stats.py:162 (np.max): NaN propagates, as in numba's array max
    m = w[0]
    if np.isnan(m):
        return m
    for i in range(1, w.shape[0]):
        v = w[i]
        if np.isnan(v):
            return v
        if v > m:
            m = v
    return m


@njit(cache=True)
def w_drange(w):
    # stats.py:34-45 via minmax :12-31 -> max - min in one sweep
    lo = w[0]
    hi = w[0]
    for i in range(1, w.shape[0]):
        if w[i] < lo:
            lo = w[i]
        if w[i] > hi:
            hi = w[i]
    return hi - lo


@njit(cache=True)
def w_skewness(w):
    # stats.py:97-110: sum(((x - mean)**3) / n) / std**3, 0 when std == 0
    sd = w_std(w)
    if sd == 0:
        return 0.0
    m = w_mean(w)
    n = w.shape[0]
    acc = 0.0
    for i in range(n):
        d = w[i] - m
        acc += (d * d * d) / n
    return acc / (sd * sd * sd)


@njit(cache=True)
def w_kurtosis(w):
    # stats.py:113-126: sum(((x - mean)**4) / n) / var**2, 0 when var == 0
    v = w_var(w)
    if v == 0:
        return 0.0
    m = w_mean(w)
    n = w.shape[0]
    acc = 0.0
    for i in range(n):
        d = w[i] - m
        d2 = d * d
        acc += (d2 * d2) / n
    return acc / (v * v)


@njit(cache=True)
def w_kurtosis_excess(w):
    # stats.py:129-139
    return w_kurtosis(w) - 3.0


@njit(cache=True)
def w_coeff_var(w):
    # stats.py:142-153: std / mean, no zero guard.  mean == 0 raises ZeroDivisionError under numba's
    # default error model; IEEE (inf / nan) is the defined behaviour here.
    m = w_mean(w)
    sd = w_std(w)
    if m == 0:
        return np.nan if sd == 0 else np.inf
    return sd / m


# --------------------------------------------------------------------------- order statistics
@njit(cache=True)
def _percentile_sorted(s, q):
    # numba/np/arraymath.py:1657-1703 (_collect_percentiles_inner): q == 0 / 100 short-circuit
    # to min / max; otherwise rank = 1 + (n-1) q/100, f = floor(rank), m = rank - f,
    # val = lower*(1-m) + upper*m with lower = s[f-1], upper = s[f].
    n = s.shape[0]
    if n == 1:
        return s[0]
    if q == 100:
        return s[n - 1]
    if q == 0:
        return s[0]
    rank = 1 + (n - 1) * (q / 100.0)
    f = math.floor(rank)
    m = rank - f
    lower = s[int(f) - 1]
    upper = s[int(f)]
    return lower * (1 - m) + upper * m


@njit(cache=True)
def w_percentile(w, q):
    # stats.py:163 (np.percentile) -- selection replaced by a full sort: the selected order
    # statistics are identical, only the interpolation arithmetic matters for parity.
    return _percentile_sorted(np.sort(w), q)


@njit(cache=True)
def w_median(w):
    # stats.py:158 -> arraymath.py:1620-1650: odd n -> middle, even n -> (a + b) / 2
    s = np.sort(w)
    n = s.shape[0]
    h = n >> 1
    if n & 1 == 0:
        return (s[h - 1] + s[h]) / 2
    return s[h]


@njit(cache=True)
def w_iqr(w):
    # stats.py:48-59: percentile 75 - percentile 25
    s = np.sort(w)
    return _percentile_sorted(s, 75.0) - _percentile_sorted(s, 25.0)


@njit(cache=True)
def w_mode(w):
    # stats.py:62-94 (jit overload): sort, then the first run that becomes strictly longer
    # than every earlier run wins.  NOTE the reference's quirk: the running counter c2 starts
    # at 0 (not 1) for the very first run, so the first run is under-counted by one.
    s = np.sort(w)
    best = s[0]
    c1 = 1
    c2 = 0
    for i in range(1, s.shape[0]):
        if s[i] == s[i - 1]:
            c2 += 1
            if c2 > c1:
                c1 = c2
                best = s[i]
        else:
            c2 = 1
    return best


# --------------------------------------------------------------------------- time domain
@njit(cache=True)
def w_zero_crossing_count(w, th):
    # timedom.py:34-64: samples with |x| <= th are zeroed, pos = x > 0, crossings = xor of
    # neighbours; zero counts as "not positive".
    n = w.shape[0]
    cnt = 0
    prev = w[0]
    if abs(prev) <= th:
        prev = 0.0
    prev_pos = prev > 0
    for i in range(1, n):
        v = w[i]
        if abs(v) <= th:
            v = 0.0
        pos = v > 0
        if pos != prev_pos:
            cnt += 1
        prev_pos = pos
    return cnt


@njit(cache=True)
def w_line_length(w):
    # timedom.py:67-78: sum |x[i+1] - x[i]|
    acc = 0.0
    for i in range(w.shape[0] - 1):
        acc += abs(w[i + 1] - w[i])
    return acc


@njit(cache=True)
def gradient(x):
    # timedom.py:11-31: one-sided at the ends, halved central difference inside
    n = x.shape[0]
    out = np.zeros(n)
    out[0] = x[1] - x[0]
    out[n - 1] = x[n - 1] - x[n - 2]
    for i in range(1, n - 1):
        out[i] = (x[i + 1] - x[i - 1]) / 2
    return out


@njit(cache=True)
def w_hjorth_activity(w):
    # timedom.py:81-95
    return w_var(w)


@njit(cache=True)
def w_hjorth_mobility(w):
    # timedom.py:98-114: sqrt(var(gradient(x)) / var(x)).  A constant window divides 0 by 0: numba's
    # default error model raises ZeroDivisionError there (and leaves garbage under prange); the defined
    # behaviour both sides of the parity tests use is IEEE: nan.
    vx = w_var(w)
    v1 = w_var(gradient(w))
    if vx == 0:
        return np.nan if v1 == 0 else np.inf
    return np.sqrt(v1 / vx)


@njit(cache=True)
def w_hjorth_complexity(w):
    # timedom.py:135-151: mobility(gradient(x)) / mobility(x)
    d1 = gradient(w)
    m0 = w_hjorth_mobility(w)
    m1 = w_hjorth_mobility(d1)
    if m0 == 0 or np.isnan(m0) or np.isnan(m1):
        return np.nan if (np.isnan(m0) or np.isnan(m1) or m1 == 0) else np.inf
    return m1 / m0


# --------------------------------------------------------------------------- information
@njit(cache=True)
def entropy(x):
    # generic/information.py:10-20: p = x / sum(x); p += 1e-30; -sum(p ln p)
    tot = 0.0
    for i in range(x.shape[0]):
        tot += x[i]
    acc = 0.0
    for i in range(x.shape[0]):
        p = x[i] / tot + 1e-30
        acc += p * math.log(p)
    return -acc
