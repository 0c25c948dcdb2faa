"""Kernel 1a (streaming window statistics) through the drop-in API vs the oracle and the
reference-generated fixtures.  Tolerances: integer-valued features bit-exact; floats rtol 1e-5
(contract) -- the float64-accumulating kernel is in practice ~1e-12."""
import functools

import numpy as np
import pytest

from conftest import split_feature, WINDOW_CASES

pytestmark = pytest.mark.gpu

STREAM = ["mean", "var", "std", "min", "max", "drange", "skewness", "kurtosis", "kurtosis_excess",
          "coeff_var", "zero_crossing_count", "line_length", "hjorth_activity"]
RTOL = 1e-5           # the stated contract (BASELINE.json north_star)
RTOL_TIGHT = 1e-9     # what float64 accumulation actually delivers on float32 inputs


def _reducers():
    from pymhealth_b200.generic import stats, timedom
    return {"mean": np.mean, "var": np.var, "std": np.std, "min": np.min, "max": np.max,
            "drange": stats.drange, "skewness": stats.skewness, "kurtosis": stats.kurtosis,
            "kurtosis_excess": stats.kurtosis_excess, "coeff_var": stats.coeff_var,
            "zero_crossing_count": timedom.zero_crossing_count, "line_length": timedom.line_length,
            "hjorth_activity": timedom.hjorth_activity}


def _scale(x):
    return float(np.sqrt(np.mean(np.asarray(x, dtype=np.float64) ** 2)))


def assert_feature_close(name, got, want, x, rtol=RTOL):
    if name in ("zero_crossing_count", "min", "max", "drange"):
        np.testing.assert_array_equal(got, want, err_msg=name)
        return
    # |a-b| <= rtol * max(|b|, scale): scale = series RMS for level-like features, 1 for shape ratios
    scale = {"mean": _scale(x), "line_length": 0.0, "var": 0.0, "std": 0.0, "hjorth_activity": 0.0}.get(name, 1.0)
    if name == "line_length":
        rtol = max(rtol, 1e-6)      # |dx| is formed and summed per 8..32-sample cell in float32, float64 across cells
    tol = rtol * np.maximum(np.abs(want), scale)
    bad = np.abs(got - want) > tol
    assert not bad.any(), "%s: %d/%d outside tolerance, worst %g" % (
        name, bad.sum(), bad.size, np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))


@pytest.mark.parametrize("case", WINDOW_CASES)
def test_golden_streaming(ref_windows, case):
    from pymhealth_b200.util import rolling_apply
    R = _reducers()
    x = ref_windows[case + "/x"]
    W, S = (int(v) for v in ref_windows[case + "/ws"])
    names = [n for n in STREAM if n != "zero_crossing_count"]
    got = rolling_apply([R[n] for n in names] + [R["zero_crossing_count"]])(x, W, S)
    for n, g in zip(names, got[:-1]):
        want = ref_windows[case + "/" + n]
        assert g.dtype == np.float64 and g.shape == want.shape
        assert_feature_close(n, g, want, x, RTOL_TIGHT)
    np.testing.assert_array_equal(got[-1], ref_windows[case + "/zero_crossing_count:0"])
    zc_th = rolling_apply(functools.partial(R["zero_crossing_count"], th=0.05))(x, W, S)
    np.testing.assert_array_equal(zc_th, ref_windows[case + "/zero_crossing_count:0.05"])
    # float64 input path
    g64 = rolling_apply([np.mean, np.var])(x.astype(np.float64), W, S)
    np.testing.assert_allclose(g64[0], ref_windows[case + "/mean"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(g64[1], ref_windows[case + "/var"], rtol=1e-9)


@pytest.mark.parametrize("n,W,S", [(4321, 500, 250), (10000, 1920, 64), (999, 64, 48), (257, 7, 3),
                                   (5000, 100, 100), (3000, 90, 120), (2048, 256, 32), (1500, 33, 1),
                                   (700, 700, 5), (40000, 500, 250), (100, 101, 1), (100, 100, 1),
                                   # co-prime / awkward geometries: no block sharing -> the direct (warp per window) fallback
                                   (30000, 479, 713), (9000, 1021, 7), (20000, 997, 2), (50000, 613, 411)])
def test_oracle_random_geometries(n, W, S):
    from oracle import windows as OW
    from pymhealth_b200.util import rolling_apply
    R = _reducers()
    rng = np.random.default_rng(n * 7 + W)
    x = (0.8 + 0.3 * rng.standard_normal(n) + np.sin(np.arange(n) * 0.05)).astype(np.float32)
    got = rolling_apply({k: R[k] for k in STREAM})(x, W, S)
    assert isinstance(got, dict) and list(got) == STREAM
    for name in STREAM:
        want = OW.rolling(name, x, W, S, 0.0 if name == "zero_crossing_count" else None)
        assert got[name].shape == want.shape
        assert_feature_close(name, got[name], want, x, RTOL_TIGHT)


def test_multi_series_device_tensor_and_f32_table():
    import torch
    from oracle import windows as OW
    from pymhealth_b200 import synth, engine
    from pymhealth_b200.generic import stats, timedom
    x = np.stack([synth.accelerometer(s, 30011) for s in range(3)]).reshape(9, -1)      # 9 series, odd length
    feats = [stats.mean.feature(), stats.std.feature(), stats.dmin.feature(), stats.dmax.feature(),
             stats.skewness.feature(), stats.kurtosis.feature(), timedom.zero_crossing_count.feature(0.0),
             timedom.line_length.feature()]
    xt = torch.from_numpy(x).cuda()
    tab = engine.window_table(xt, 500, 250, feats)
    assert tab.is_cuda and tab.dtype == torch.float32 and tuple(tab.shape) == (9, 119, 8)
    tab = tab.cpu().numpy().astype(np.float64)
    names = ["mean", "std", "min", "max", "skewness", "kurtosis", "zero_crossing_count", "line_length"]
    for s in range(9):
        for j, name in enumerate(names):
            want = OW.rolling(name, x[s], 500, 250, 0.0)
            assert_feature_close(name, tab[s, :, j], want.astype(np.float32).astype(np.float64) if name in ("min", "max") else want, x[s], 2e-6)
    # a strided view (rows of a wider buffer) is consumed in place
    wide = torch.zeros((4, 30100), dtype=torch.float32, device="cuda")
    wide[:, :30011] = xt[:4]
    tab2 = engine.window_table(wide[:, :30011], 500, 250, feats)
    torch.testing.assert_close(tab2, engine.window_table(xt[:4].contiguous(), 500, 250, feats), rtol=0, atol=0)


def test_edge_cases():
    from pymhealth_b200.util import rolling_apply
    from pymhealth_b200.generic import stats
    assert rolling_apply(np.mean)(np.ones(5, np.float32), 8, 2).shape == (0,)          # fewer samples than a window
    const = np.full(1000, 3.25, np.float32)
    out = rolling_apply([np.var, stats.skewness, stats.kurtosis, stats.kurtosis_excess])(const, 100, 50)
    assert np.all(out[0] == 0) and np.all(out[1] == 0) and np.all(out[2] == 0) and np.all(out[3] == -3)
    with pytest.raises(NotImplementedError):
        rolling_apply(lambda w: w.sum())
    with pytest.raises(ValueError):
        rolling_apply(np.mean)(np.ones(10, np.float32), 0, 1)
    # integer input is promoted exactly
    xi = np.arange(100, dtype=np.int32)
    np.testing.assert_array_equal(rolling_apply(np.mean)(xi, 10, 10), np.arange(10) * 10 + 4.5)
    # direct call = one window
    x = np.random.default_rng(1).standard_normal(257).astype(np.float32)
    assert stats.skewness(x) == pytest.approx(float(__import__("oracle.reducers", fromlist=["x"]).w_skewness(x.astype(np.float64))), rel=1e-9)


def test_size_independent_properties_full_config2():
    """Config-2 size (3 x 4 320 000, W=500, S=250): properties that need no oracle pass."""
    from pymhealth_b200 import synth, engine
    from pymhealth_b200.generic import stats
    import torch
    x = torch.from_numpy(synth.accelerometer(0, 4_320_000)).cuda()
    feats = [stats.mean.feature(), stats.var.feature(), stats.dmin.feature(), stats.dmax.feature(), stats._sum.feature()]
    tab = engine.window_table(x, 500, 250, feats, out_dtype=torch.float64)
    assert tuple(tab.shape) == (3, 17279, 5)
    # (1) non-overlapping windows tile the series: their sums add up to the series sum
    tot = tab[:, 0::2, 4].sum(dim=1)
    ref = x[:, :17280 // 2 * 500].double().sum(dim=1)
    torch.testing.assert_close(tot, ref, rtol=1e-12, atol=1e-9)
    # (2) min <= mean <= max, var >= 0
    assert bool((tab[..., 2] <= tab[..., 0]).all() and (tab[..., 0] <= tab[..., 3]).all() and (tab[..., 1] >= 0).all())
    # (3) affine covariance: features of a*x+b
    y = (x * 2.0 + 1.0)
    tab2 = engine.window_table(y, 500, 250, feats, out_dtype=torch.float64)
    torch.testing.assert_close(tab2[..., 0], tab[..., 0] * 2 + 1, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(tab2[..., 1], tab[..., 1] * 4, rtol=1e-5, atol=1e-9)
    # (4) shift invariance of the window grid: dropping the first hop shifts the table by one row
    tab3 = engine.window_table(x[:, 250:], 500, 250, feats, out_dtype=torch.float64)
    torch.testing.assert_close(tab3[:, :, :4], tab[:, 1:, :4], rtol=1e-10, atol=1e-12)


def test_integer_sensor_counts():
    """int16 / uint8 / int32 input (raw counts): widened on the device, same results as the float64 oracle."""
    from oracle import windows as OW
    from pymhealth_b200.util import rolling_apply
    from pymhealth_b200.generic import stats, timedom
    rng = np.random.default_rng(3)
    for dt, lo, hi in ((np.int16, -2000, 2000), (np.uint8, 0, 255), (np.int32, -10**6, 10**6)):
        x = rng.integers(lo, hi, 7000).astype(dt)
        got = rolling_apply([np.mean, np.var, np.min, np.max, stats.skewness, np.median,
                             functools.partial(timedom.zero_crossing_count, th=0.0)])(x, 128, 32)
        for g, name in zip(got, ["mean", "var", "min", "max", "skewness", "median", "zero_crossing_count"]):
            want = OW.rolling(name, x.astype(np.float64), 128, 32, 0.0)
            assert g.dtype == np.float64
            if name in ("min", "max", "median", "zero_crossing_count"):
                np.testing.assert_array_equal(g, want, err_msg="%s %s" % (dt.__name__, name))
            else:
                np.testing.assert_allclose(g, want, rtol=1e-9, atol=1e-9, err_msg="%s %s" % (dt.__name__, name))


@pytest.mark.parametrize("n,W,S", [(40000, 500, 250), (60000, 1920, 64), (9000, 64, 48), (5000, 100, 100)])
def test_non_finite_samples_stay_local(n, W, S):
    """Sensor gaps stored as NaN (and a stray inf): only the windows that CONTAIN such a sample may be affected, and they
    must propagate exactly as the reference's numpy reducers do (np.min / np.max return NaN, a crossing with NaN is
    "not positive").  The first sample of the series (the pivot of the shifted power sums) is one of them."""
    from oracle import windows as OW
    from pymhealth_b200.util import rolling_apply
    R = _reducers()
    rng = np.random.default_rng(W + S)
    x = (0.8 + 0.3 * rng.standard_normal(n)).astype(np.float32)
    bad = np.array([0, 7 * S + 3, n // 2, n // 2 + 1, n - 1])
    x[bad] = np.nan
    x[n // 3] = np.inf
    x[2 * n // 3] = -np.inf
    names = ["mean", "var", "std", "min", "max", "drange", "skewness", "kurtosis", "zero_crossing_count", "line_length"]
    got = rolling_apply([R[k] for k in names])(x, W, S)
    nw = 1 + (n - W) // S
    starts = np.arange(nw) * S
    touched = np.zeros(nw, dtype=bool)
    for b in list(bad) + [n // 3, 2 * n // 3]:
        touched |= (starts <= b) & (b < starts + W)
    assert touched.sum() < nw // 2
    for name, g in zip(names, got):
        want = OW.rolling(name, x, W, S, 0.0 if name == "zero_crossing_count" else None)
        clean = ~touched
        assert np.all(np.isfinite(g[clean])), name                       # nothing leaks into the other windows
        assert_feature_close(name, g[clean], want[clean], x[np.isfinite(x)], RTOL_TIGHT)
        if name in ("min", "max", "zero_crossing_count"):                 # exact propagation where the reference defines it
            np.testing.assert_array_equal(g[touched], want[touched], err_msg=name)
        elif name in ("mean", "line_length"):
            np.testing.assert_array_equal(np.isnan(g[touched]), np.isnan(want[touched]), err_msg=name)
            np.testing.assert_array_equal(np.isposinf(g[touched]), np.isposinf(want[touched]), err_msg=name)
