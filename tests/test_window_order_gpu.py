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


# ---- the streaming kernel (window_order_stream.cu): float32, W = g or 2 g with S = g, only median / percentile / IQR in
# the call -- every lane-group size (g <= 32, 64, 128, 256), odd block lengths, chunk boundaries (127 / 128 windows per
# warp chunk), tiny series, several series of one launch
@pytest.mark.parametrize("n,W,S", [(40_000, 500, 250), (64_250, 250, 250), (9_000, 64, 32), (7_001, 96, 48), (30_011, 200, 100),
                                   (5_000, 34, 17), (70_000, 512, 256), (2_600, 32, 16), (3_333, 33, 33), (500, 500, 250),
                                   (750, 500, 250), (1_249, 500, 250), (32_500, 500, 250), (32_750, 500, 250)])
def test_stream_kernel_vs_oracle(n, W, S):
    from oracle import windows as OW
    from pymhealth_b200 import engine
    from pymhealth_b200.generic import stats
    rng = np.random.default_rng(n * 7 + W)
    x = np.round(rng.standard_normal((3, n)) * 40).astype(np.float32) / 16      # many ties
    x[1] = rng.standard_normal(n).astype(np.float32) * 1e-3 + 1.0
    x[2, ::7] = -x[2, ::7]
    feats = [stats.median.feature(), stats.percentile.feature(0.0), stats.percentile.feature(100.0),
             stats.percentile.feature(37.5), stats.percentile.feature(99.99), stats.interquartile_range.feature(),
             stats.percentile.feature(90.0)]
    got = engine.window_table(x, W, S, feats)                       # numpy in -> float64 table [3, nw, 7]
    names = [("median", None), ("percentile", 0.0), ("percentile", 100.0), ("percentile", 37.5), ("percentile", 99.99),
             ("iqr", None), ("percentile", 90.0)]
    for s in range(3):
        for j, (nme, q) in enumerate(names):
            want = OW.rolling(nme, x[s], W, S, q)
            assert got[s, :, j].shape == want.shape
            if j < 3:           # selected samples (the even-length median is an exact float64 mean of two float32 values)
                np.testing.assert_array_equal(got[s, :, j], want, err_msg="%s q=%s series %d" % (nme, q, s))
            else:
                np.testing.assert_allclose(got[s, :, j], want, rtol=1e-13, atol=1e-300, err_msg="%s q=%s series %d" % (nme, q, s))


def test_stream_kernel_matches_batch_kernel(monkeypatch):
    """Same table, bit for bit, as the batch kernel it replaced on this geometry (window_order_blocks.cu)."""
    from pymhealth_b200 import engine, synth
    from pymhealth_b200.generic import stats
    x = synth.accelerometer(3, 400_000)
    feats = [stats.median.feature(), stats.percentile.feature(90.0), stats.interquartile_range.feature()]
    a = engine.window_table(x, 500, 250, feats)
    monkeypatch.setenv("MHB_ORDER_NOSTREAM", "1")
    b = engine.window_table(x, 500, 250, feats)
    np.testing.assert_allclose(a, b, rtol=1e-15, atol=0)
    np.testing.assert_array_equal(a[:, :, 0], b[:, :, 0])
