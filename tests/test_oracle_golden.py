"""The oracle against fixtures generated from the unmodified reference (tests/golden/make_golden.py)
and against the reference's own known-answer tests.  CPU only."""
import numpy as np
import pytest

from oracle import windows as OW, reducers as OR, spectral as OS, location as OL
from conftest import split_feature, WINDOW_CASES

RTOL = 1e-12      # oracle vs reference: same float64 algorithm, only summation order may differ


@pytest.mark.parametrize("case", WINDOW_CASES)
def test_window_reducers_match_reference(ref_windows, case):
    x = ref_windows[case + "/x"]
    W, S = (int(v) for v in ref_windows[case + "/ws"])
    keys = [k.split("/", 1)[1] for k in ref_windows.files if k.startswith(case + "/")]
    checked = 0
    for key in keys:
        if key in ("x", "ws"):
            continue
        name, p = split_feature(key)
        got = OW.rolling(name, x, W, S, p)
        want = ref_windows[case + "/" + key]
        assert got.shape == want.shape and got.dtype == np.float64
        if name in ("zero_crossing_count", "mode", "min", "max", "median" if W % 2 else "min"):
            np.testing.assert_array_equal(got, want, err_msg=key)
        else:
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13, err_msg=key)
        checked += 1
    assert checked >= 25


def test_window_count_and_view(ref_windows):
    xv = ref_windows["view/x"]
    np.testing.assert_array_equal(OW.view(xv, 5, 3), ref_windows["view/5_3"])
    for n, w, s in [(10, 3, 1), (10, 10, 4), (9, 10, 1), (0, 1, 1), (6137, 500, 250)]:
        assert OW.n_windows(n, w, s) == max(0, 1 + (n - w) // s)
    assert OW.rolling("mean", np.ones(5, np.float32), 8, 2).shape == (0,)


def test_direct_calls(ref_windows):
    w = ref_windows["direct/x"].astype(np.float64)
    lo, hi = ref_windows["direct/minmax"]
    assert OR.w_min(w) == lo and OR.w_max(w) == hi
    np.testing.assert_allclose(OR.gradient(w), ref_windows["direct/gradient"], rtol=0, atol=0)
    act, mob, cpx = ref_windows["direct/hjorth_parameters"]
    assert OR.w_hjorth_activity(w) == pytest.approx(act, rel=RTOL)
    assert OR.w_hjorth_mobility(w) == pytest.approx(mob, rel=RTOL)
    assert OR.w_hjorth_complexity(w) == pytest.approx(cpx, rel=RTOL)
    zc = ref_windows["direct/zero_crossings_0.9"]
    assert OR.w_zero_crossing_count(w, 0.9) == int(zc.sum())
    pm = ref_windows["direct/percentile_multi"]
    for q, v in zip((5, 50, 95), pm):
        assert OR.w_percentile(w, float(q)) == pytest.approx(v, rel=RTOL)


def test_nonuniform(ref_windows):
    idx, vals = ref_windows["nonuniform/index"], ref_windows["nonuniform/vals"]
    np.testing.assert_array_equal(OW.get_indices(idx, 60, 20), ref_windows["nonuniform/indices_60_20"])
    np.testing.assert_allclose(OW.nonuniform_rolling("mean", idx, vals, 60, 20),
                               ref_windows["nonuniform/mean_60_20"], rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(OW.nonuniform_rolling("max", idx, vals, 60, 20),
                               ref_windows["nonuniform/max_60_20"], rtol=0, equal_nan=True)
    np.testing.assert_allclose(OW.nonuniform_rolling("std", idx, vals, 3, 20, min_window_len=5),
                               ref_windows["nonuniform/std_60_20_min5"], rtol=RTOL, equal_nan=True)


@pytest.mark.parametrize("case", ["acc", "ppg", "odd"])
def test_spectral_chain(ref_spectral, case):
    x = ref_spectral[case + "/x"]
    W, S, fs = ref_spectral[case + "/wsf"]
    W, S = int(W), int(S)
    np.testing.assert_allclose(OS.fft(x[:W].astype(np.float64)), ref_spectral[case + "/fft0"], rtol=1e-13, atol=1e-10)
    np.testing.assert_allclose(OS.ifft(ref_spectral[case + "/fft0"]), ref_spectral[case + "/ifft0"], rtol=1e-13, atol=1e-12)
    bands = [tuple(b) for b in ref_spectral[case + "/bands"]]
    tab = OS.spectral_table(x, W, S, fs, bands, 0.3, 12.0)
    psd, freqs = OS.window_psd(x, W, S, fs)
    for j in range(len(bands)):
        np.testing.assert_allclose(tab["band_power_%d" % j], ref_spectral[case + "/band_power"][:, j], rtol=1e-12)
        np.testing.assert_allclose(tab["rel_band_power_%d" % j], ref_spectral[case + "/rel_band_power"][:, j], rtol=1e-12)
    np.testing.assert_array_equal(tab["peak_frequency"], ref_spectral[case + "/peak_frequency_0.3_12"])
    np.testing.assert_allclose(tab["spectral_entropy"], ref_spectral[case + "/entropy"], rtol=1e-12)
    np.testing.assert_allclose(tab["total_power"], ref_spectral[case + "/total_power"], rtol=1e-12)
    # scalar forms agree with the table
    for i in (0, psd.shape[0] - 1):
        assert OS.power_band(psd[i], freqs, *bands[0]) == pytest.approx(ref_spectral[case + "/band_power"][i, 0], rel=1e-12)
        assert OS.peak_frequency(psd[i], freqs) == ref_spectral[case + "/peak_frequency_all"][i]
        assert OS.spectral_entropy(psd[i]) == pytest.approx(ref_spectral[case + "/entropy"][i], rel=1e-12)


def test_haversine_known_answers(ref_location):
    """The reference's own test vectors (tests/location/test_distance.py:16-58).  They were written
    for 2r = 12742.0 while the code uses 12742.018: rescaled, they pin the
    formula to ~1e-11; the live-reference outputs pin the constant."""
    pts = ref_location["points"]
    la, lo = pts[:, 0].copy(), pts[:, 1].copy()
    k = 12742.018 / 12742.0
    assert OL.haversine(la[0], lo[0], la[1], lo[1]) == pytest.approx(float(ref_location["stale/scalar_0_1"]) * k, rel=1e-12)
    np.testing.assert_almost_equal(OL.haversine_elementwise(la[:-1], lo[:-1], la[1:], lo[1:]) / k,
                                   ref_location["stale/elementwise"], decimal=7)
    np.testing.assert_almost_equal(OL.haversine_vector(la[0], lo[0], la[1:], lo[1:]) / k,
                                   ref_location["stale/vector"], decimal=7)
    assert OL.haversine(la[0], lo[0], la[1], lo[1]) == pytest.approx(float(ref_location["ref/scalar_0_1"]), rel=1e-15)
    np.testing.assert_allclose(OL.haversine_elementwise(la[:-1], lo[:-1], la[1:], lo[1:]), ref_location["ref/elementwise"], rtol=1e-14)
    np.testing.assert_allclose(OL.haversine_vector(la[0], lo[0], la[1:], lo[1:]), ref_location["ref/vector"], rtol=1e-14)
    np.testing.assert_allclose(OL.haversine_outer_product(la, lo, la, lo), ref_location["ref/outer"], rtol=1e-14, atol=1e-9)


def test_location_features(ref_location):
    from pymhealth_b200 import synth
    n, period = (int(v) for v in ref_location["gps/n_period"])
    lat, lon, t, home = synth.gps(0, n, period)
    np.testing.assert_allclose(home, ref_location["gps/home"], rtol=0)
    np.testing.assert_allclose(OL.arr_successive_distance(lat, lon), ref_location["gps/successive_distance"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(OL.arr_distance_from_home(lat, lon, home), ref_location["gps/distance_from_home"], rtol=1e-13, atol=1e-15)
    assert OL.arr_proportion_home_stay(lat, lon, 0.1, home) == float(ref_location["gps/proportion_home_stay_0.1"])
    assert OL.arr_proportion_home_stay(lat, lon, 5.0, home) == float(ref_location["gps/proportion_home_stay_5"])
    assert OL.arr_location_variance(lat, lon) == pytest.approx(float(ref_location["gps/location_variance"]), rel=1e-12)
    labels = ref_location["labels/x"]
    assert OL.num_clusters(labels) == int(ref_location["labels/num_clusters"])
    tot = OL.cluster_totals(labels)
    np.testing.assert_array_equal(np.array(sorted(tot)), ref_location["labels/totals_keys"])
    np.testing.assert_array_equal(np.array([tot[k] for k in sorted(tot)]), ref_location["labels/totals_vals"])
    assert OL.cluster_entropy(labels) == pytest.approx(float(ref_location["labels/entropy"]), rel=1e-13)
    assert OL.normalized_cluster_entropy(labels) == pytest.approx(float(ref_location["labels/normalized_entropy"]), rel=1e-13)
    assert OL.normalized_cluster_entropy(labels, 8) == pytest.approx(float(ref_location["labels/normalized_entropy_n8"]), rel=1e-13)
    assert OR.entropy(ref_location["entropy/counts"]) == pytest.approx(float(ref_location["entropy/value"]), rel=1e-13)


def test_extension_oracle_sanity():
    """radius of gyration / stay points have no reference (parity unpinned): property checks only."""
    from oracle import location_ext as OX
    from pymhealth_b200 import synth
    lat, lon, t, home = synth.gps(1, 2880, 60)
    lab = OX.stay_points(lat, lon, t, 0.2, 1800)
    assert lab.min() >= -1 and lab.max() >= 0
    # labels are non-decreasing over the stay ids and every stay lasts >= min_dur
    for k in range(lab.max() + 1):
        idx = np.nonzero(lab == k)[0]
        assert idx[-1] - idx[0] + 1 == len(idx)
        assert t[idx[-1]] - t[idx[0]] >= 1800
    rg = OX.radius_of_gyration(lat, lon)
    assert 0 < rg < 60
    tab, labels = OX.segment_features(lat, lon, t, np.array([0, 1440, 2880]), np.array([home, home]), 0.1, 0.2, 1800)
    assert tab.shape == (2, len(OX.SEG_COLUMNS))
    assert tab[:, 0].sum() == 2880


def test_extra_oracles_pinned():
    """oracle/accel.py and oracle/hrv.py against fixtures produced by the live reference (ref_extra.npz)."""
    import os
    from conftest import GOLDEN
    from oracle import accel as OA, hrv as OH
    ref = np.load(os.path.join(GOLDEN, "ref_extra.npz"))
    a = ref["acc/xyz"]
    got = OA.magnitude(a[0], a[1], a[2])
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, ref["acc/magnitude_f32"])
    np.testing.assert_allclose(OA.roll(a[1], a[2]), ref["acc/roll_f32"], rtol=0, atol=180 * 2.0 ** -22)
    np.testing.assert_allclose(OA.pitch(a[0], a[1], a[2]), ref["acc/pitch_f32"], rtol=0, atol=180 * 2.0 ** -22)
    a64 = a.astype(np.float64)
    np.testing.assert_array_equal(OA.magnitude(a64[0], a64[1], a64[2]), ref["acc/magnitude_f64"])
    np.testing.assert_allclose(OA.roll(a64[1], a64[2]), ref["acc/roll_f64"], rtol=1e-15, atol=1e-13)
    np.testing.assert_allclose(OA.pitch(a64[0], a64[1], a64[2]), ref["acc/pitch_f64"], rtol=1e-15, atol=1e-13)
    assert OA.magnitude_dot(a64[0], a64[1], a64[2]) == float(ref["acc/magnitude_dot_f64"])
    rr = ref["hrv/rr_ms"]
    assert OH.sdnn(rr) == float(ref["hrv/sdnn"])
    assert OH.pnnx(rr) == float(ref["hrv/pnn50"])
    assert OH.pnnx(rr, "ms", 20.0) == float(ref["hrv/pnnx_20"])
    assert OH.pnnx(rr / 1e3, "s") == float(ref["hrv/pnn50_s"])
    assert abs(OH.rmssd(rr) - float(ref["hrv/rmssd"])) <= 1e-12 * float(ref["hrv/rmssd"])
    assert abs(OH.ssd(rr) - float(ref["hrv/ssd"])) <= 1e-9
    assert abs(OH.sdsd(rr) - float(ref["hrv/sdsd"])) <= 1e-12 * float(ref["hrv/sdsd"])


def test_oracle_min_max_propagate_nan_like_numba():
    """The reference's np.min / np.max run inside numba, whose array min / max return NaN as soon as one is met; the
    restatement must do the same (the GPU kernels use FMNMX.NAN for it)."""
    import numba

    @numba.njit
    def ref(a):
        return np.min(a), np.max(a)
    rng = np.random.default_rng(3)
    for trial in range(20):
        w = rng.standard_normal(17)
        if trial % 4:
            w[rng.integers(0, 17)] = np.nan
        lo, hi = ref(w)
        np.testing.assert_array_equal([OR.w_min(w), OR.w_max(w)], [lo, hi])


def test_location_extensions_against_independent_implementations():
    """The north-star's extensions (radius of gyration, stay points) have no reference implementation, so
    oracle/location_ext.py is their definition ("parity unpinned").  What CAN be pinned: the radius of gyration against
    scikit-learn's great-circle distance (an independent implementation of the same haversine formula), and the
    stay-point scan against a plain-Python restatement of the published anchor scan written from the docstring alone."""
    import math
    from sklearn.metrics.pairwise import haversine_distances
    from pymhealth_b200 import synth
    from oracle import location_ext as OX
    lat, lon, t, _ = synth.gps(3, 4000, 60)
    r_km = 12742.018 / 2.0                                    # location/distance.py:8,18
    c = np.radians([[lat.mean(), lon.mean()]])
    d = haversine_distances(np.radians(np.stack([lat, lon], axis=1)), c)[:, 0] * r_km
    np.testing.assert_allclose(OX.radius_of_gyration(lat, lon), math.sqrt(np.mean(d * d)), rtol=1e-9)

    def hav(a1, o1, a2, o2):
        a1, o1, a2, o2 = map(math.radians, (a1, o1, a2, o2))
        h = math.sin((a2 - a1) / 2) ** 2 + math.cos(a1) * math.cos(a2) * math.sin((o2 - o1) / 2) ** 2
        return 2 * r_km * math.asin(math.sqrt(h))
    for dist_km, min_dur in ((0.2, 1800), (0.05, 600), (1.0, 7200)):
        want = np.full(len(lat), -1, dtype=np.int64)
        i = k = 0
        while i < len(lat):
            j = i + 1
            while j < len(lat) and hav(lat[i], lon[i], lat[j], lon[j]) <= dist_km:
                j += 1
            if t[j - 1] - t[i] >= min_dur:
                want[i:j] = k
                k += 1
            i = j
        np.testing.assert_array_equal(OX.stay_points(lat, lon, t, dist_km, min_dur), want)
