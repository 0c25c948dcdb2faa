"""Non-uniform (timestamp-indexed) windows: get_indices / indices_rolling_apply / nonuniform_rolling_apply
(reference src/mhealth/util/windows.py:122-249) vs the reference-generated fixtures and the oracle."""
import functools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_golden_nonuniform(ref_windows):
    from pymhealth_b200.util.windows import get_indices, nonuniform_rolling_apply
    idx, vals = ref_windows["nonuniform/index"], ref_windows["nonuniform/vals"]
    ind = get_indices(idx, 60, 20)
    assert ind.dtype == np.int64
    np.testing.assert_array_equal(ind, ref_windows["nonuniform/indices_60_20"])        # indices: bit-exact
    got = nonuniform_rolling_apply(np.mean)(idx, vals, 60, 20)
    assert got.dtype == vals.dtype
    np.testing.assert_allclose(got, ref_windows["nonuniform/mean_60_20"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(nonuniform_rolling_apply(np.max)(idx, vals, 60, 20), ref_windows["nonuniform/max_60_20"])
    got = nonuniform_rolling_apply(np.std, 5)(idx, vals, 3, 20)
    want = ref_windows["nonuniform/std_60_20_min5"]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))                       # short windows -> NaN
    np.testing.assert_allclose(got, want, rtol=1e-10, equal_nan=True)


def _gappy_index(rng, n, dtype=np.int64):
    gaps = rng.integers(1, 4, n)
    gaps[rng.random(n) < 0.01] += rng.integers(50, 400)          # recording gaps
    return np.cumsum(gaps).astype(dtype) + 1000


@pytest.mark.parametrize("n,wsize,wstep,minlen", [(5000, 120, 40, 1), (20000, 600, 60, 10), (300, 7, 3, 1),
                                                  (3000, 5000, 1000, 1), (40000, 90000, 30000, 1)])
def test_oracle_nonuniform_all_reducers(n, wsize, wstep, minlen):
    from oracle import windows as OW
    from pymhealth_b200.util.windows import nonuniform_rolling_apply, get_indices
    from pymhealth_b200.generic import stats, timedom
    rng = np.random.default_rng(n + wsize)
    idx = _gappy_index(rng, n)
    x = (np.round(rng.standard_normal(n) * 20) / 4 + 9.81).astype(np.float32)
    np.testing.assert_array_equal(get_indices(idx, wsize, wstep), OW.get_indices(idx, wsize, wstep))
    funcs = {"mean": np.mean, "var": np.var, "std": np.std, "min": np.min, "max": np.max, "drange": stats.drange,
             "skewness": stats.skewness, "kurtosis": stats.kurtosis, "kurtosis_excess": stats.kurtosis_excess,
             "coeff_var": stats.coeff_var, "zero_crossing_count": functools.partial(timedom.zero_crossing_count, th=9.5),
             "line_length": timedom.line_length, "median": np.median,
             "percentile": functools.partial(np.percentile, q=80), "iqr": stats.interquartile_range}
    got = nonuniform_rolling_apply(funcs, minlen)(idx, x, wsize, wstep)
    assert list(got) == list(funcs)
    for name, g in got.items():
        p = 9.5 if name == "zero_crossing_count" else 80.0
        want = OW.nonuniform_rolling(name, idx, x.astype(np.float64), wsize, wstep, minlen, p)
        assert g.dtype == np.float32 and g.shape == want.shape
        np.testing.assert_array_equal(np.isnan(g), np.isnan(want), err_msg=name)
        ok = ~np.isnan(want)
        if name in ("min", "max", "zero_crossing_count", "median"):
            np.testing.assert_array_equal(g[ok], want[ok].astype(np.float32), err_msg=name)
        else:
            # float32 output cells (the array's dtype, windows.py:149) of float64 arithmetic
            np.testing.assert_allclose(g[ok], want[ok].astype(np.float32), rtol=2e-7, atol=1e-12, err_msg=name)


def test_float64_series_and_datetime_index():
    from oracle import windows as OW
    from pymhealth_b200.util.windows import nonuniform_rolling_apply, indices_rolling_apply, get_indices
    from pymhealth_b200.generic import timedom
    rng = np.random.default_rng(5)
    n = 8000
    secs = _gappy_index(rng, n)
    t = secs.astype("datetime64[s]")
    x = rng.standard_normal(n)
    wsize, wstep = np.timedelta64(5, "m"), np.timedelta64(60, "s")
    ind = get_indices(t, wsize, wstep)
    np.testing.assert_array_equal(ind, OW.get_indices(secs, 300, 60))
    got = nonuniform_rolling_apply([np.mean, np.std, timedom.hjorth_mobility])(t, x, wsize, wstep)
    for g, name in zip(got, ["mean", "std", "hjorth_mobility"]):
        want = OW.indices_rolling(name, ind, x, 1)
        assert g.dtype == np.float64
        if name == "hjorth_mobility":                     # windows of one sample have no gradient: NaN here
            ok = (ind[1] - ind[0]) >= 2
            np.testing.assert_allclose(g[ok], want[ok], rtol=1e-9, equal_nan=True)
        else:
            np.testing.assert_allclose(g, want, rtol=1e-10, atol=1e-14, equal_nan=True)
    # the indices form, arbitrary (unordered, overlapping, empty) windows
    ar = np.array([[0, 10, 5, 7999, 4000, 100], [n, 10, 900, n, 4001, 50]])
    got = indices_rolling_apply(np.var, 1)(ar, x)
    want = OW.indices_rolling("var", ar, x, 1)
    np.testing.assert_allclose(got, want, rtol=1e-10, equal_nan=True)
    assert np.isnan(got[1]) and np.isnan(got[5])


def test_float_index_and_unsupported():
    from oracle import windows as OW
    from pymhealth_b200.util.windows import get_indices, nonuniform_rolling_apply
    rng = np.random.default_rng(9)
    idx = np.cumsum(rng.random(3000) + 0.01)
    np.testing.assert_array_equal(get_indices(idx, 12.5, 2.25), OW.get_indices(idx, 12.5, 2.25))
    with pytest.raises(NotImplementedError):
        nonuniform_rolling_apply(lambda w: w.sum())


def test_long_windows():
    """Day-long windows of a 1 Hz series (86 400 samples each)."""
    from oracle import windows as OW
    from pymhealth_b200.util.windows import nonuniform_rolling_apply
    rng = np.random.default_rng(11)
    n = 3 * 86400
    idx = np.arange(n, dtype=np.int64)
    x = rng.standard_normal(n) + 50.0
    got = nonuniform_rolling_apply([np.mean, np.var, np.min, np.max])(idx, x, 86400, 43200)
    for g, name in zip(got, ["mean", "var", "min", "max"]):
        want = OW.nonuniform_rolling(name, idx, x, 86400, 43200)
        np.testing.assert_allclose(g, want, rtol=1e-10, err_msg=name)


def test_get_indices_float_keys_on_window_boundaries():
    """Regularly sampled FLOAT timestamps put index values exactly on window boundaries, where the left searchsorted
    depends on the last bit of every key: numpy fills arange(start, stop, step) as start + i * delta with
    delta = (start + step) - start (float64), which is not always the double `step` (0.1, 0.2 -> 0.20000000000000004)."""
    from oracle import windows as OW
    from pymhealth_b200.util.windows import get_indices
    for start, dt, wsize, wstep in [(0.1, 0.1, 0.6, 0.2), (0.1, 0.05, 1.0, 0.2), (1e6 + 0.3, 0.1, 0.7, 0.3), (0.0, 0.25, 2.0, 0.5),
                                    (17.7, 0.02, 0.3, 0.06)]:
        idx = start + dt * np.arange(20_000)
        want = OW.get_indices(idx, wsize, wstep)
        got = get_indices(idx, wsize, wstep)
        np.testing.assert_array_equal(got, want, err_msg=str((start, dt, wsize, wstep)))
