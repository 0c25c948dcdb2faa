"""Size-independent properties at the full sizes of BASELINE configs 4 and 5 (no oracle pass needed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config4_stats_one_subject_day():
    """PPG 64 Hz x 24 h (5 529 600 samples), W = 1920, S = 64: 86 371 windows, 30x overlap."""
    import torch
    from pymhealth_b200 import synth, engine
    from pymhealth_b200.generic import stats, timedom
    n, W, S = 5_529_600, 1920, 64
    x = synth.ppg(11, n)
    xd = torch.from_numpy(x).cuda()
    feats = [stats.mean.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(),
             timedom.zero_crossing_count.feature(0.0), timedom.line_length.feature()]
    tab = engine.window_table(xd, W, S, feats, out_dtype=torch.float64).cpu().numpy()
    nw = 1 + (n - W) // S
    assert tab.shape == (nw, 6)
    x64 = x.astype(np.float64)
    cs = np.concatenate([[0.0], np.cumsum(x64)])
    starts = np.arange(nw) * S
    np.testing.assert_allclose(tab[:, 0], (cs[starts + W] - cs[starts]) / W, rtol=1e-9, atol=1e-11)       # prefix-sum identity
    cs2 = np.concatenate([[0.0], np.cumsum(x64 * x64)])
    var = (cs2[starts + W] - cs2[starts]) / W - tab[:, 0] ** 2
    np.testing.assert_allclose(tab[:, 1], var, rtol=1e-6, atol=1e-9)
    # extrema: exact, via block minima (hop 64 blocks, 30 per window)
    bmin = x[: (n // S) * S].reshape(-1, S).min(axis=1)
    bmax = x[: (n // S) * S].reshape(-1, S).max(axis=1)
    k = W // S
    wmin = np.lib.stride_tricks.sliding_window_view(bmin, k).min(axis=1)[:nw]
    wmax = np.lib.stride_tricks.sliding_window_view(bmax, k).max(axis=1)[:nw]
    np.testing.assert_array_equal(tab[:, 2], wmin.astype(np.float64))
    np.testing.assert_array_equal(tab[:, 3], wmax.astype(np.float64))
    # zero crossings / line length from prefix sums of the pair terms (pairs inside the window: W - 1 of them)
    pos = x > 0
    zc_pair = (pos[1:] != pos[:-1]).astype(np.int64)
    czc = np.concatenate([[0], np.cumsum(zc_pair)])
    np.testing.assert_array_equal(tab[:, 4], (czc[starts + W - 1] - czc[starts]).astype(np.float64))
    ll_pair = np.abs(np.diff(x64))
    cll = np.concatenate([[0.0], np.cumsum(ll_pair)])
    np.testing.assert_allclose(tab[:, 5], cll[starts + W - 1] - cll[starts], rtol=1e-5)


def test_config4_spectral_parseval_and_shift():
    """W = 1920 / S = 64 fast path on a long series: Parseval and shift invariance of the window grid."""
    import torch
    from pymhealth_b200 import synth, engine, spectral as SP
    n, W, S, fs = 600_000, 1920, 64, 64.0
    x = synth.ppg(12, n)
    xd = torch.from_numpy(x).cuda()
    feats = [SP.total_power(fs).feature(), SP.peak_bin(fs, 0.5, 4.0).feature(), SP.spectral_entropy(fs).feature()]
    tab = engine.window_table(xd, W, S, feats, fs=fs, out_dtype=torch.float64).cpu().numpy()
    nw = 1 + (n - W) // S
    assert tab.shape == (nw, 3)
    # Parseval on the one-sided PSD: sum_k c_k |X_k|^2 = W sum x^2 with c_0 = c_{W/2} = 1, c_k = 2; total power counts
    # every bin once, so  W sum x^2 = 2 total - |X_0|^2 - |X_{W/2}|^2
    x64 = x.astype(np.float64)
    cs2 = np.concatenate([[0.0], np.cumsum(x64 * x64)])
    cs = np.concatenate([[0.0], np.cumsum(x64)])
    alt = x64 * np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    csa = np.concatenate([[0.0], np.cumsum(alt)])
    starts = np.arange(nw) * S
    e = W * (cs2[starts + W] - cs2[starts])
    x0 = cs[starts + W] - cs[starts]
    xn = csa[starts + W] - csa[starts]                          # S even: the alternating sign pattern is window-invariant
    total = (e + x0 ** 2 + xn ** 2) / 2
    np.testing.assert_allclose(tab[:, 0], total, rtol=2e-6)
    assert np.all((tab[:, 1] >= 15) & (tab[:, 1] < 120))          # bins of 0.5 .. 4 Hz at 1/30 Hz per bin
    assert np.all((tab[:, 2] > 0) & (tab[:, 2] < np.log(961)))
    tab2 = engine.window_table(xd[S * 7:], W, S, feats, fs=fs, out_dtype=torch.float64).cpu().numpy()
    np.testing.assert_array_equal(tab2, tab[7:])                  # same windows, other batches / lanes: bit-identical


def test_config5_month_of_1hz_gps():
    """One subject-month at 1 Hz (2 592 000 points, 30 day segments): kernels agree with each other and with the oracle
    on two of the days."""
    from oracle import location as OL
    from pymhealth_b200 import synth
    from pymhealth_b200.location import features
    day, ndays = 86400, 30
    lat, lon, t, home = synth.gps(9, day * ndays, 1)
    offs = np.arange(ndays + 1) * day
    rows = features.segment_rows(lat, lon, t, offs, [home] * ndays, limit=0.1)
    assert rows.shape == (ndays, 11)
    assert np.all(rows[:, 0] == day)
    d = features.arr_successive_distance(lat, lon)
    assert d[0] == 0.0
    for k in range(ndays):
        seg = d[k * day + 1:(k + 1) * day]                        # the step INTO a day belongs to no day
        np.testing.assert_allclose(rows[k, 1], seg.sum(), rtol=1e-9)
    dh = features.arr_distance_from_home(lat, lon, home)
    np.testing.assert_array_equal(rows[:, 5], (dh.reshape(ndays, day) < 0.1).sum(axis=1).astype(np.float64))   # counts: exact
    np.testing.assert_allclose(rows[:, 4], dh.reshape(ndays, day).max(axis=1), rtol=1e-12)
    assert np.all((rows[:, 6] >= 0) & (rows[:, 6] <= 1))
    for k in (0, 17):
        sl = slice(k * day, (k + 1) * day)
        np.testing.assert_allclose(d[sl][1:], OL.arr_successive_distance(lat[sl], lon[sl])[1:], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rows[k, 2], OL.arr_location_variance(lat[sl], lon[sl]), rtol=1e-9)


@pytest.mark.gpu
def test_series_longer_than_2_31_samples():
    """One 8.6 GB series (2^31 + 1,000,327 samples): 64-bit sample offsets INSIDE a series for kernels 1a, 1b and 2,
    checked against the oracle at the head, across the 2^31 crossing and at the tail (tools/big_series_check.py)."""
    import os
    import runpy
    import torch
    if torch.cuda.mem_get_info()[0] < 24 * 2**30:
        pytest.skip("needs ~20 GB of free HBM")
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "big_series_check.py")
    with pytest.raises(SystemExit) as ex:
        runpy.run_path(tool, run_name="__main__")
    assert ex.value.code == 0


def test_config2_order_statistics_one_subject_week():
    """Accelerometer 50 Hz x 7 d (30 240 000 samples), W = 500, S = 250: 120 959 windows through the streaming order
    kernel.  Properties that need no oracle pass: the 0th / 100th percentiles ARE the window extrema (bit for bit against
    kernel 1a and against a numpy block reduction), quantiles are monotone in q, IQR = p75 - p25, the median is odd under
    negation and homogeneous under scaling by 2 (both exact), and a sorted-window recomputation of 64 windows
    picked across the week (chunk boundaries included) reproduces every column."""
    import torch
    from pymhealth_b200 import synth, engine
    from pymhealth_b200.generic import stats
    n, W, S = 30_240_000, 500, 250
    x = synth.accelerometer(5, n)[2]
    xd = torch.from_numpy(x).cuda()
    qs = [0.0, 10.0, 25.0, 50.0, 75.0, 90.0, 100.0]
    feats = [stats.percentile.feature(q) for q in qs] + [stats.median.feature(), stats.interquartile_range.feature()]
    tab = engine.window_table(xd, W, S, feats, out_dtype=torch.float64).cpu().numpy()
    nw = 1 + (n - W) // S
    assert tab.shape == (nw, 9)
    ext = engine.window_table(xd, W, S, [stats.dmin.feature(), stats.dmax.feature()], out_dtype=torch.float64).cpu().numpy()
    np.testing.assert_array_equal(tab[:, 0], ext[:, 0])
    np.testing.assert_array_equal(tab[:, 6], ext[:, 1])
    bmin = x[: (n // S) * S].reshape(-1, S).min(axis=1)
    np.testing.assert_array_equal(tab[:, 0], np.minimum(bmin[:-1], bmin[1:])[:nw].astype(np.float64))
    assert np.all(np.diff(tab[:, :7], axis=1) >= 0)                       # monotone in q
    np.testing.assert_array_equal(tab[:, 7], tab[:, 3])                   # median == p50 (even W: the same two samples, weight 1/2)
    np.testing.assert_allclose(tab[:, 8], tab[:, 4] - tab[:, 2], rtol=1e-15, atol=0)
    neg = engine.window_table(-xd, W, S, [stats.median.feature()], out_dtype=torch.float64).cpu().numpy()[:, 0]
    np.testing.assert_array_equal(neg, -tab[:, 7])
    # scaling the data by 2 is exact, so every selected sample -- and the interpolation -- scales exactly
    sc = engine.window_table(xd * 2.0, W, S, [stats.median.feature(), stats.percentile.feature(90.0)],
                             out_dtype=torch.float64).cpu().numpy()
    np.testing.assert_array_equal(sc[:, 0], 2.0 * tab[:, 7])
    np.testing.assert_allclose(sc[:, 1], 2.0 * tab[:, 5], rtol=1e-15, atol=0)
    pick = np.unique(np.concatenate([np.arange(0, 6), np.arange(124, 132), np.arange(252, 258), [nw - 3, nw - 2, nw - 1],
                                     np.random.default_rng(3).integers(0, nw, 41)]))
    for w in pick:
        seg = x[w * S: w * S + W].astype(np.float64)
        want = [np.percentile(seg, q) for q in qs] + [np.median(seg), np.percentile(seg, 75) - np.percentile(seg, 25)]
        np.testing.assert_allclose(tab[w], want, rtol=1e-13, atol=0)
