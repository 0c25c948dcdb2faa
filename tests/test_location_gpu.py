"""Kernel 3 (location traces): the reference's own haversine known-answer tests, fixtures generated
from the live reference, and the extension oracle (radius of gyration / stay points)."""
import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu
K = 12742.018 / 12742.0      # the reference's goldens were written for 2r = 12742.0


def test_haversine_like_the_reference_tests(ref_location):
    """Mirrors reference tests/location/test_distance.py:16-58 (same POINTS, same assertions), with the
    stale-radius goldens rescaled by 12742.018/12742.0."""
    from pymhealth_b200.location import distance
    POINTS = ref_location["points"]
    lat1, lon1 = POINTS[0]
    lat2, lon2 = POINTS[1]
    assert distance.haversine(lat1, lon1, lat2, lon2) == pytest.approx(7704.777296228049 * K)
    lats, lons = POINTS[:, 0], POINTS[:, 1]
    out = distance.haversine_elementwise(lats[:-1], lons[:-1], lats[1:], lons[1:])
    npt.assert_almost_equal(out / K, ref_location["stale/elementwise"])
    out = distance.haversine_vector(lats[0], lons[0], lats[1:], lons[1:])
    npt.assert_almost_equal(out / K, ref_location["stale/vector"])
    out = distance.haversine_outer_product(lats, lons, lats, lons)
    npt.assert_allclose(out, ref_location["ref/outer"], rtol=1e-13, atol=1e-9)
    # against the live reference (pins the constant)
    npt.assert_allclose(distance.haversine_elementwise(lats[:-1], lons[:-1], lats[1:], lons[1:]),
                        ref_location["ref/elementwise"], rtol=1e-14)
    # float32 input is promoted like the reference's float64-only gufunc does
    out32 = distance.haversine_vector(lats[0], lons[0], lats[1:].astype(np.float32), lons[1:].astype(np.float32))
    assert out32.dtype == np.float64


def test_trace_features_match_reference(ref_location):
    from pymhealth_b200 import synth
    from pymhealth_b200.location import features, distribution
    from pymhealth_b200.generic import information
    n, period = (int(v) for v in ref_location["gps/n_period"])
    lat, lon, t, home = synth.gps(0, n, period)
    lat0, lon0 = lat.copy(), lon.copy()
    npt.assert_allclose(features.arr_successive_distance(lat, lon), ref_location["gps/successive_distance"], rtol=1e-11, atol=1e-13)
    npt.assert_allclose(features.arr_distance_from_home(lat, lon, home), ref_location["gps/distance_from_home"], rtol=1e-11, atol=1e-13)
    assert features.arr_proportion_home_stay(lat, lon, 0.1, home) == float(ref_location["gps/proportion_home_stay_0.1"])
    assert features.arr_proportion_home_stay(lat, lon, 5.0, home) == float(ref_location["gps/proportion_home_stay_5"])
    assert distribution.arr_location_variance(lat, lon) == pytest.approx(float(ref_location["gps/location_variance"]), rel=1e-9)
    npt.assert_array_equal(lat, lat0)
    npt.assert_array_equal(lon, lon0)                     # inputs untouched (features.py:107)
    labels = ref_location["labels/x"]
    assert distribution.num_clusters(labels) == int(ref_location["labels/num_clusters"])
    tot = distribution.cluster_totals(labels)
    npt.assert_array_equal(np.array(sorted(tot)), ref_location["labels/totals_keys"])
    npt.assert_array_equal(np.array([tot[k] for k in sorted(tot)]), ref_location["labels/totals_vals"])
    assert distribution.cluster_entropy(labels) == pytest.approx(float(ref_location["labels/entropy"]), rel=1e-13)
    assert distribution.normalized_cluster_entropy(labels) == pytest.approx(float(ref_location["labels/normalized_entropy"]), rel=1e-13)
    assert distribution.normalized_cluster_entropy(labels, 8) == pytest.approx(float(ref_location["labels/normalized_entropy_n8"]), rel=1e-13)
    assert information.entropy(ref_location["entropy/counts"]) == pytest.approx(float(ref_location["entropy/value"]), rel=1e-13)
    assert np.isnan(distribution.normalized_cluster_entropy(np.zeros(10, dtype=np.int64)))     # one label: 0/0


def test_segment_rows_vs_oracle():
    """Config-1 (one subject, 7 days at 1/min) and a multi-subject table against the oracle; counts and
    stay-point labels bit-exact, floats 1e-9."""
    from oracle import location_ext as OX
    from pymhealth_b200 import synth, _lib as L
    from pymhealth_b200.location import features
    lats, lons, ts, offs, homes = [], [], [], [0], []
    for sid in range(5):
        n = 10080 if sid == 0 else 4000 + 317 * sid
        lat, lon, t, home = synth.gps(sid, n, 60)
        day = 1440
        for a in range(0, n, day):
            offs.append(offs[-1] + min(day, n - a))
            homes.append(home)
        lats.append(lat), lons.append(lon), ts.append(t)
    offs.insert(3, offs[3])                     # an EMPTY segment in the middle
    homes.insert(3, homes[3])
    lat, lon, t = np.concatenate(lats), np.concatenate(lons), np.concatenate(ts)
    home = np.array(homes)
    rows, labels = features.segment_rows(lat, lon, t, offs, home, 0.1, 0.2, 1800, labels=True)
    want, wlab = OX.segment_features(lat, lon, t, np.array(offs), home, 0.1, 0.2, 1800)
    assert rows.shape == want.shape == (len(offs) - 1, len(L.SEG_COLUMNS))
    npt.assert_array_equal(labels, wlab)
    for c, name in enumerate(L.SEG_COLUMNS):
        if name in ("n_points", "home_stay_count", "n_stay_points", "n_labels"):
            npt.assert_array_equal(rows[:, c], want[:, c], err_msg=name)
        else:
            npt.assert_allclose(rows[:, c], want[:, c], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=name)
    # the single-trace extension helpers
    assert features.radius_of_gyration(lats[1], lons[1]) == pytest.approx(float(OX.radius_of_gyration(lats[1], lons[1])), rel=1e-10)
    npt.assert_array_equal(features.stay_points(lats[1], lons[1], ts[1], 0.2, 1800), OX.stay_points(lats[1], lons[1], ts[1], 0.2, 1800))


def test_successive_distance_segments_and_properties():
    import torch
    from pymhealth_b200 import synth, _lib as L
    from pymhealth_b200.engine import _stream_ptr
    lat, lon, t, home = synth.gps(3, 50000, 1)
    la, lo = torch.from_numpy(lat).cuda(), torch.from_numpy(lon).cuda()
    offs = torch.tensor([0, 20000, 20000, 50000], dtype=torch.int64, device="cuda")
    out = torch.empty_like(la)
    L.check(L.load().mhb_successive_distance(la.data_ptr(), lo.data_ptr(), offs.data_ptr(), 3, 50000, out.data_ptr(),
                                             _stream_ptr(torch)), "sd")
    out = out.cpu().numpy()
    assert out[0] == 0 and out[20000] == 0 and out[19999] > 0
    from oracle import location as OL
    ref = OL.arr_successive_distance(lat, lon)
    ref[20000] = 0
    npt.assert_allclose(out, ref, rtol=1e-9, atol=1e-13)
    # symmetry / identity properties at size
    from pymhealth_b200.location import distance
    d1 = distance.haversine_elementwise(lat[:-1], lon[:-1], lat[1:], lon[1:])
    d2 = distance.haversine_elementwise(lat[1:], lon[1:], lat[:-1], lon[:-1])
    npt.assert_allclose(d1, d2, rtol=1e-12, atol=1e-15)
    assert np.all(distance.haversine_elementwise(lat, lon, lat, lon) == 0)
