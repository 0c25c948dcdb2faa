"""Host-side logic that needs no GPU: window counts, index-window counts, reducer resolution, shard ranges."""
import functools

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5000), w=st.integers(1, 600), s=st.integers(1, 600))
def test_window_count_matches_the_strided_view(n, w, s):
    from pymhealth_b200 import engine
    from oracle import windows as OW
    nw = engine.n_windows(n, w, s)
    assert nw == (OW.n_windows(n, w, s) if n >= w else 0)           # windows.py:86, tail dropped
    if n >= w:
        assert nw == OW.view(np.zeros(n), w, s).shape[0]


@settings(max_examples=200, deadline=None)
@given(first=st.integers(-10**6, 10**6), span=st.integers(0, 10**6), step=st.integers(1, 10**5))
def test_index_window_count_integers(first, span, step):
    from pymhealth_b200 import engine
    assert engine.n_index_windows(first, first + span, step, False) == len(np.arange(first, first + span, step))


@settings(max_examples=200, deadline=None)
@given(first=st.floats(-1e6, 1e6), span=st.floats(0, 1e3), step=st.floats(1e-3, 1e4))      # <= 1e6 keys per case
def test_index_window_count_floats(first, span, step):
    from pymhealth_b200 import engine
    assert engine.n_index_windows(first, first + span, step, True) == len(np.arange(first, first + span, step))


def test_reducer_resolution():
    from pymhealth_b200 import reducers, _lib as L
    from pymhealth_b200.generic import stats, timedom
    f, _ = reducers.resolve(np.mean)
    assert (f.family, f.fid) == ("stream", L.F_MEAN)
    f, _ = reducers.resolve(functools.partial(np.percentile, q=90))
    assert (f.family, f.fid, f.params) == ("order", L.F_PERCENTILE, (90.0,))
    f, integer = reducers.resolve(functools.partial(timedom.zero_crossing_count, th=0.25))
    assert (f.fid, f.params, integer) == (L.F_ZERO_CROSSINGS, (0.25,), True)
    assert reducers.resolve(stats.kurtosis_excess)[0].fid == L.F_KURTOSIS_EXCESS
    for bad in (lambda w: w.sum(), np.ptp, functools.partial(np.mean, axis=0), "mean"):
        with pytest.raises(NotImplementedError):
            reducers.resolve(bad)


def test_no_cuda_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pymhealth_b200 import _lib
    from pymhealth_b200.util import rolling_apply
    with pytest.raises(_lib.MhbError):
        rolling_apply(np.mean)(np.arange(100.0), 10, 5)


def test_product_view_is_the_reference_view(ref_windows):
    """pymhealth_b200.util.windows.view (util/windows.py:20-33 of the reference): same shape, same values, zero-copy --
    against the reference-generated fixture and the strided definition, for 1-D float / int input."""
    from pymhealth_b200.util.windows import view
    from oracle import windows as OW
    xv = ref_windows["view/x"]
    got = view(xv, 5, 3)
    np.testing.assert_array_equal(got, ref_windows["view/5_3"])
    assert np.shares_memory(got, xv) and got.dtype == xv.dtype
    rng = np.random.default_rng(5)
    for n, w, s in [(10, 3, 1), (10, 10, 4), (6137, 500, 250), (100, 7, 13), (64, 64, 1)]:
        for x in (rng.standard_normal(n), rng.integers(-9, 9, n).astype(np.int16)):
            v = view(x, w, s)
            assert v.shape == (1 + (n - w) // s, w)
            np.testing.assert_array_equal(v, OW.view(x, w, s))
            for i in (0, v.shape[0] - 1):
                np.testing.assert_array_equal(v[i], x[i * s:i * s + w])
