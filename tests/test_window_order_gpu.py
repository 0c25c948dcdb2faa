"""Kernel 1b (order statistics, mode, Hjorth) vs the reference-generated fixtures and the oracle."""
import functools

import numpy as np
import pytest

from conftest import WINDOW_CASES

pytestmark = pytest.mark.gpu
RTOL = 1e-12     # selected order statistics are exact; interpolation / Hjorth are float64 on both sides


def _funcs():
    from pymhealth_b200.generic import stats, timedom
    f = {"median": np.median, "iqr": stats.interquartile_range, "hjorth_mobility": timedom.hjorth_mobility,
         "hjorth_complexity": timedom.hjorth_complexity}
    for q in (10, 25, 50, 90, 99.5, 0, 100):
        f["percentile:%g" % q] = functools.partial(np.percentile, q=q)
    return f


@pytest.mark.parametrize("case", WINDOW_CASES)
def test_golden_order(ref_windows, case):
    from pymhealth_b200.util import rolling_apply
    from pymhealth_b200.generic import stats
    x = ref_windows[case + "/x"]
    W, S = (int(v) for v in ref_windows[case + "/ws"])
    F = _funcs()
    got = rolling_apply(F)(x, W, S)
    for name, g in got.items():
        want = ref_windows[case + "/" + name]
        assert g.shape == want.shape and g.dtype == np.float64
        if name.startswith("hjorth"):
            np.testing.assert_allclose(g, want, rtol=1e-9, err_msg=name)
        else:
            np.testing.assert_allclose(g, want, rtol=RTOL, atol=0, err_msg=name)
    if case + "/mode" in ref_windows.files:
        np.testing.assert_array_equal(rolling_apply(stats.mode)(x, W, S), ref_windows[case + "/mode"])
    # float64 input goes through the double kernels
    g64 = rolling_apply(np.median)(x.astype(np.float64), W, S)
    np.testing.assert_array_equal(g64, ref_windows[case + "/median"])


@pytest.mark.parametrize("n,W,S", [(3000, 500, 250), (9000, 1920, 640), (999, 33, 7), (300, 2, 1), (50000, 5000, 2500),
                                   (70000, 30000, 10000)])
def test_oracle_order_random(n, W, S):
    from oracle import windows as OW
    from pymhealth_b200.util import rolling_apply
    from pymhealth_b200.generic import stats, timedom
    rng = np.random.default_rng(n + W)
    x = np.round(rng.standard_normal(n) * 50).astype(np.float32) / 8      # many ties
    got = rolling_apply([np.median, functools.partial(np.percentile, q=37.5), stats.interquartile_range, stats.mode,
                         timedom.hjorth_mobility, timedom.hjorth_complexity])(x, W, S)
    names = ["median", "percentile", "iqr", "mode", "hjorth_mobility", "hjorth_complexity"]
    for nme, g in zip(names, got):
        want = OW.rolling(nme, x, W, S, 37.5)
        if nme.startswith("hjorth"):
            np.testing.assert_allclose(g, want, rtol=1e-9, err_msg=nme, equal_nan=True)
        else:
            np.testing.assert_allclose(g, want, rtol=1e-13, atol=0, err_msg=nme)


def test_direct_calls_and_mixed_families(ref_windows):
    from pymhealth_b200.generic import stats, timedom
    from pymhealth_b200.util import rolling_apply
    w = ref_windows["direct/x"]
    np.testing.assert_allclose(stats.percentile(w, [5, 50, 95]), ref_windows["direct/percentile_multi"], rtol=1e-13)
    act, mob, cpx = timedom.hjorth_parameters(w)
    np.testing.assert_allclose([act, mob, cpx], ref_windows["direct/hjorth_parameters"], rtol=1e-9)
    lo, hi = stats.minmax(w)
    assert (lo, hi) == tuple(ref_windows["direct/minmax"].astype(np.float32))
    assert timedom.zero_crossing_count(w, 0.9) == int(ref_windows["direct/zero_crossings_0.9"].sum())
    # one call mixing the streaming and the order kernels keeps the requested column order
    x = ref_windows["acc_z_500_250/x"]
    got = rolling_apply([np.median, np.mean, stats.interquartile_range, np.std])(x, 500, 250)
    for g, k in zip(got, ["median", "mean", "iqr", "std"]):
        np.testing.assert_allclose(g, ref_windows["acc_z_500_250/" + k], rtol=1e-9)
    with pytest.raises(ValueError):
        rolling_apply(functools.partial(np.percentile, q=101))(x, 500, 250)


@pytest.mark.parametrize("W,S", [(64, 16), (256, 256), (500, 250), (33, 7)])
def test_percentile_next_to_100(W, S):
    """q one ulp below 100: numba's rank 1 + (n - 1) q / 100 rounds to n, i.e. the largest element with weight 1 --
    never a read past the window.  Also q next to 0."""
    from pymhealth_b200.util import rolling_apply
    rng = np.random.default_rng(W)
    x = rng.standard_normal(4000).astype(np.float32)
    qhi, qlo = float(np.nextafter(100.0, 0.0)), float(np.nextafter(0.0, 1.0))
    hi, lo, mx, mn = rolling_apply([functools.partial(np.percentile, q=qhi), functools.partial(np.percentile, q=qlo),
                                    np.max, np.min])(x, W, S)
    assert np.all(np.isfinite(hi)) and np.all(np.isfinite(lo))
    np.testing.assert_allclose(hi, mx, rtol=1e-12)
    np.testing.assert_allclose(lo, mn, rtol=1e-12, atol=1e-300)
