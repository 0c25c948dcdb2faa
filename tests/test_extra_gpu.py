"""Accelerometer pre-stage, HRV time-domain metrics and DataFrame front-ends vs fixtures generated from the live
reference (tests/golden/ref_extra.npz) and the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_extra():
    return np.load(os.path.join(GOLDEN, "ref_extra.npz"))


def test_accelerometer_golden(ref_extra):
    import pandas as pd
    from pymhealth_b200.inertial import accelerometer as acc
    a = ref_extra["acc/xyz"]
    m = acc.magnitude(a[0], a[1], a[2])
    assert m.dtype == np.float32
    np.testing.assert_array_equal(m, ref_extra["acc/magnitude_f32"])            # float32 arithmetic, bit-identical
    r, p = acc.roll(a[1], a[2]), acc.pitch(a[0], a[1], a[2])
    assert r.dtype == np.float64 and p.dtype == np.float64
    # arctan2 in float32 (CUDA atan2f vs libm atan2f: a few float32 ulp), tolerance on degrees
    np.testing.assert_allclose(r, ref_extra["acc/roll_f32"], rtol=0, atol=180 * 4 * 2.0 ** -23)
    np.testing.assert_allclose(p, ref_extra["acc/pitch_f32"], rtol=0, atol=180 * 4 * 2.0 ** -23)
    a64 = a.astype(np.float64)
    np.testing.assert_array_equal(acc.magnitude(a64[0], a64[1], a64[2]), ref_extra["acc/magnitude_f64"])
    np.testing.assert_allclose(acc.roll(a64[1], a64[2]), ref_extra["acc/roll_f64"], rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(acc.pitch(a64[0], a64[1], a64[2]), ref_extra["acc/pitch_f64"], rtol=1e-14, atol=1e-13)
    assert acc.magnitude_dot(a64[0], a64[1], a64[2]) == pytest.approx(float(ref_extra["acc/magnitude_dot_f64"]), rel=1e-13)
    s = ref_extra["acc/scalars"]
    assert acc.magnitude(1.0, 2.0, 2.0) == s[0]
    assert acc.roll(1.0, 1.0) == pytest.approx(s[1], rel=1e-15)
    assert acc.pitch(1.0, 0.0, 1.0) == pytest.approx(s[2], rel=1e-15)
    df = pd.DataFrame({"x": a64[0], "y": a64[1], "z": a64[2]})
    got = acc.magnitude(df)
    assert isinstance(got, pd.Series) and got.name == "magnitude"
    np.testing.assert_array_equal(got.values, ref_extra["acc/df_magnitude"])
    np.testing.assert_allclose(acc.roll(df).values, ref_extra["acc/df_roll"], rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(acc.pitch(df).values, ref_extra["acc/df_pitch"], rtol=1e-14, atol=1e-13)
    assert acc.magnitude_dot(df) == pytest.approx(float(ref_extra["acc/magnitude_dot_f64"]), rel=1e-13)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 100003])
def test_accelerometer_oracle_sizes(n):
    import torch
    from oracle import accel as OA
    from pymhealth_b200.inertial import accelerometer as acc
    rng = np.random.default_rng(n)
    x, y, z = (rng.standard_normal(n).astype(np.float32) * 3 for _ in range(3))
    np.testing.assert_array_equal(acc.magnitude(x, y, z), OA.magnitude(x, y, z))
    np.testing.assert_allclose(acc.roll(y, z), OA.roll(y, z), rtol=0, atol=180 * 4 * 2.0 ** -23)
    np.testing.assert_allclose(acc.pitch(x, y, z), OA.pitch(x, y, z), rtol=0, atol=180 * 4 * 2.0 ** -23)
    # unaligned device views take the scalar kernel; device tensors stay on the device
    xt = torch.from_numpy(np.concatenate([[0], x]).astype(np.float32)).cuda()[1:]
    got = acc.magnitude(xt, torch.from_numpy(y).cuda(), torch.from_numpy(z).cuda())
    assert got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), OA.magnitude(x, y, z))
    xi = np.arange(n)
    assert acc.magnitude(xi, xi, xi).dtype == np.float64       # integers are promoted, as numba does
    assert acc.magnitude_dot(x, y, z) == pytest.approx(OA.magnitude_dot(x.astype(np.float64), y.astype(np.float64),
                                                                        z.astype(np.float64)), rel=1e-12)


def test_hrv_time_domain_golden(ref_extra):
    from oracle import hrv as OH
    from pymhealth_b200.heart import hrv
    rr = ref_extra["hrv/rr_ms"]
    assert hrv.sdnn(rr) == pytest.approx(float(ref_extra["hrv/sdnn"]), rel=1e-12)
    assert hrv.pnn50(rr, 'ms') == float(ref_extra["hrv/pnn50"])                 # counts: exact
    assert hrv.pnnx(rr, 'ms', 20.0) == float(ref_extra["hrv/pnnx_20"])
    assert hrv.pnn50(rr / 1e3, 's') == float(ref_extra["hrv/pnn50_s"])
    assert hrv.rmssd(rr) == pytest.approx(float(ref_extra["hrv/rmssd"]), rel=1e-12)
    assert hrv.ssd(rr) == pytest.approx(float(ref_extra["hrv/ssd"]), rel=1e-9, abs=1e-9)
    assert hrv.sdsd(rr) == pytest.approx(float(ref_extra["hrv/sdsd"]), rel=1e-12)
    np.testing.assert_array_equal(hrv.nni_to_ms(rr[:16] * 1e6, 'ns'), ref_extra["hrv/nni_to_ms"])
    got = [hrv.csi_sd1(rr), hrv.csi_sd2(rr), hrv.lorenz_csi(rr), hrv.lorenz_cvi(rr), hrv.lorenz_mcsi(rr), hrv.csi_sd2(rr, 0.5)]
    np.testing.assert_allclose(got, ref_extra["hrv/poincare"], rtol=1e-11)                 # Poincare / Lorenz indices
    with pytest.raises(ValueError):
        hrv.td_factor("h")
    # segment metrics: the reference's own versions do not compile under numba 0.65 (parity unpinned) -> oracle
    assert hrv.sdann(rr, unit='ms', interval=120.0) == pytest.approx(OH.sdann(rr, unit='ms', interval=120.0), rel=1e-10)
    assert hrv.sdnni(rr, unit='ms', interval=120.0) == pytest.approx(OH.sdnni(rr, unit='ms', interval=120.0), rel=1e-10)
    big = np.random.default_rng(1).normal(800, 50, 300001)
    assert hrv.rmssd(big) == pytest.approx(OH.rmssd(big), rel=1e-12)
    assert hrv.pnn50(big) == OH.pnnx(big)


def test_location_dataframe_forms(ref_extra):
    import pandas as pd
    from pymhealth_b200.location import features, distribution
    lat, lon, t = ref_extra["gps/lat"], ref_extra["gps/lon"], ref_extra["gps/t"]
    gdf = pd.DataFrame({"latitude": lat, "longitude": lon}, index=pd.to_datetime(t, unit="s"))
    home = features.determine_home_coords(gdf)
    np.testing.assert_array_equal(np.array(home), ref_extra["gps/home"])
    d = features.distance_from_home(gdf)
    assert isinstance(d, pd.Series) and d.name == "home_distance" and d.index.equals(gdf.index)
    np.testing.assert_allclose(d.values, ref_extra["gps/distance_from_home"], rtol=1e-9, atol=1e-12)
    assert features.proportion_home_stay(gdf, 0.5) == float(ref_extra["gps/proportion_home_stay_0.5"])
    sd = features.successive_distance(gdf)
    assert isinstance(sd, pd.Series)
    # the reference's DataFrame form is broken under pandas 3 (see make_golden.py); the array form is the golden
    assert sd.index.equals(gdf.index)
    np.testing.assert_allclose(sd.values, ref_extra["gps/arr_successive_distance"], rtol=1e-9, atol=1e-12)
    assert distribution.location_variance(gdf) == pytest.approx(float(ref_extra["gps/location_variance"]), rel=1e-10)


def test_ppg_slope_sum(ref_extra):
    from pymhealth_b200.heart import ppg
    x = ref_extra["ppg/x"]
    for w in (9, 1):
        got = ppg.slope_sum(x, w)
        want = ref_extra["ppg/slope_sum_%d" % w]
        assert got.dtype == np.float64 and got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(x).max())
        assert np.all(got[:w] == 0) and got[-1] == 0
    got32 = ppg.slope_sum(x.astype(np.float32), 9)                 # float32 samples, float64 sums
    np.testing.assert_allclose(got32, ref_extra["ppg/slope_sum_9"], rtol=0, atol=1e-6 * np.abs(x).max())
    assert ppg.slope_sum(np.zeros(0), 3).shape == (0,)
    assert np.all(ppg.slope_sum(np.arange(5.0), 9) == 0)           # window longer than the signal


def test_gradient_and_zero_crossings(ref_extra):
    from pymhealth_b200.generic import timedom
    g = ref_extra["td/x"]
    np.testing.assert_array_equal(timedom.gradient(g), ref_extra["td/gradient_f64"])          # same operations: exact
    got32 = timedom.gradient(g.astype(np.float32))
    assert got32.dtype == np.float64
    np.testing.assert_array_equal(got32, ref_extra["td/gradient_f32"])                         # float32 differences
    c = g - g.mean()
    z0, zt = timedom.zero_crossings(c, 0.0), timedom.zero_crossings(c, 0.05)
    assert z0.dtype == bool and z0.shape == (len(g) - 1,)
    np.testing.assert_array_equal(z0, ref_extra["td/zero_crossings_0"])
    np.testing.assert_array_equal(zt, ref_extra["td/zero_crossings_th"])
    assert int(z0.sum()) == timedom.zero_crossing_count(c, 0.0)
    assert int(zt.sum()) == timedom.zero_crossing_count(c, 0.05)
    np.testing.assert_array_equal(timedom.gradient(np.array([1.0, 4.0])), [3.0, 3.0])
    with pytest.raises(ValueError):
        timedom.gradient(np.array([1.0]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("W,S,n", [(500, 250, 40000), (128, 32, 9000), (1920, 64, 30000), (300, 300, 5000), (50, 7, 3000)])
def test_rolling_magnitude_fused(dtype, W, S, n):
    """window statistics of magnitude(x, y, z) with the magnitude formed inside kernel 1a's staging copy
    -- bit-identical to materialising the magnitude and rolling over it, and equal to the oracle chain."""
    from oracle import accel as OA, windows as OW
    from pymhealth_b200 import synth, spectral as SP
    from pymhealth_b200.generic import stats, timedom
    from pymhealth_b200.inertial import accelerometer as acc
    from pymhealth_b200.util import rolling_apply
    x, y, z = (a.astype(dtype) for a in synth.accelerometer(17, n))
    funcs = {"mean": np.mean, "std": np.std, "min": np.min, "max": np.max, "skew": stats.skewness, "kurt": stats.kurtosis,
             "zc": timedom.zero_crossing_count, "ll": timedom.line_length, "median": np.median}
    got = acc.rolling_magnitude(funcs, W, S)(x, y, z)
    mag = acc.magnitude(x, y, z)
    two_step = rolling_apply(funcs, W, S)(mag)
    want_mag = OA.magnitude(x, y, z)
    np.testing.assert_array_equal(mag, want_mag)
    for name in funcs:
        # the same kernel arithmetic on the same magnitudes; only the cell partition / pivot of the float64 partial sums
        # may differ (bit-identical when the block size is not a multiple of 8, e.g. W=500 / S=250)
        if name in ("min", "max", "zc", "median") or (W, S) == (500, 250):
            np.testing.assert_array_equal(got[name], two_step[name], err_msg=name)
        else:
            np.testing.assert_allclose(got[name], two_step[name], rtol=1e-6 if name == "ll" else 1e-10, atol=1e-12, err_msg=name)
    for name, oname in [("mean", "mean"), ("std", "std"), ("max", "max"), ("kurt", "kurtosis"), ("median", "median")]:
        np.testing.assert_allclose(got[name], OW.rolling(oname, want_mag.astype(np.float64) if dtype is np.float64 else want_mag,
                                                         W, S), rtol=1e-9, atol=1e-12, err_msg=name)
    if dtype is np.float32 and W == 500:
        h = acc.rolling_magnitude([np.mean, SP.spectral_entropy(50.0)], W, S)(x, y, z)
        h2 = rolling_apply([np.mean, SP.spectral_entropy(50.0)], W, S)(mag)
        np.testing.assert_array_equal(h[1], h2[1])
    # one reducer form, ragged shapes
    one = acc.rolling_magnitude(np.var)(x[:W + 3], y[:W + 3], z[:W + 3], W, S)
    assert one.shape == (1,)
    assert acc.rolling_magnitude(np.var, W, S)(x[:W - 1], y[:W - 1], z[:W - 1]).shape == (0,)
    with pytest.raises(ValueError):
        acc.rolling_magnitude(np.var, W, S)(x, y[:-1], z)
