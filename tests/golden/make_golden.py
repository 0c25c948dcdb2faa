#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (the reference checkout is not present on the GPU box):

    PYTHONPATH=/root/reference/src PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports callumstew/pymhealth (``mhealth``) from /root/reference/src, feeds it seeded
synthetic inputs (pymhealth_b200.synth) and stores inputs' parameters + the reference's outputs
as small .npz files.  The compatibility shims of SURVEY section 8c live HERE, never in the
reference: ``.py_func`` for @jit dispatchers, thin wrappers instead of bare numpy aliases, a
stub ``hdbscan`` module, ``numpy.fft`` over ``util.windows.view`` for the spectral chain.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True
sys.modules.setdefault("hdbscan", types.ModuleType("hdbscan"))

import mhealth                                                       # noqa: E402
import mhealth.fft                                                   # noqa: E402  (prints the numpy-fallback notice)
from mhealth.util.windows import (rolling_apply, view, get_indices,  # noqa: E402
                                  nonuniform_rolling_apply, indices_rolling_apply)
from mhealth.generic import stats, timedom, information              # noqa: E402
from mhealth.generic.frequency import density                        # noqa: E402
from mhealth.heart import hrv                                        # noqa: E402
from mhealth.location import distance, features, distribution        # noqa: E402

from pymhealth_b200 import synth                                     # noqa: E402


# ---- shims: never pass a bare numpy alias or a CPUDispatcher to rolling_apply
def wmean(w): return np.mean(w)
def wvar(w): return np.var(w)
def wstd(w): return np.std(w)
def wmin(w): return np.min(w)
def wmax(w): return np.max(w)
def wmedian(w): return np.median(w)
def wp10(w): return np.percentile(w, 10)
def wp25(w): return np.percentile(w, 25)
def wp50(w): return np.percentile(w, 50)
def wp90(w): return np.percentile(w, 90)
def wp99_5(w): return np.percentile(w, 99.5)
def wp0(w): return np.percentile(w, 0)
def wp100(w): return np.percentile(w, 100)
def wzc0(w): return _zc(w, 0.0)
def wzc_th(w): return _zc(w, 0.05)


_zc = timedom.zero_crossing_count       # a dispatcher may be CALLED from a jitted closure

REDUCERS = {
    "mean": wmean, "var": wvar, "std": wstd, "min": wmin, "max": wmax, "median": wmedian,
    "percentile:10": wp10, "percentile:25": wp25, "percentile:50": wp50, "percentile:90": wp90,
    "percentile:99.5": wp99_5, "percentile:0": wp0, "percentile:100": wp100,
    "drange": stats.drange, "iqr": stats.interquartile_range,
    "kurtosis_excess": stats.kurtosis_excess, "mode": stats.mode,
    "skewness": stats.skewness.py_func, "kurtosis": stats.kurtosis.py_func,
    "coeff_var": stats.coeff_var.py_func,
    "zero_crossing_count:0": wzc0, "zero_crossing_count:0.05": wzc_th,
    "line_length": timedom.line_length.py_func,
    "hjorth_activity": timedom.hjorth_activity.py_func,
    "hjorth_mobility": timedom.hjorth_mobility.py_func,
    "hjorth_complexity": timedom.hjorth_complexity.py_func,
}

# (case name, series builder, wsize, wstep)
CASES = [
    ("acc_z_500_250", lambda: synth.accelerometer(7, 6137)[2], 500, 250),     # config-2 geometry, ragged tail
    ("acc_x_500_250", lambda: synth.accelerometer(7, 6137)[0], 500, 250),     # near-zero-mean axis
    ("ppg_1920_64", lambda: synth.ppg(3, 6000), 1920, 64),                    # config-4 geometry
    ("acc_y_64_48", lambda: synth.accelerometer(11, 1000)[1], 64, 48),        # gcd 16
    ("acc_z_7_3", lambda: synth.accelerometer(12, 200)[2], 7, 3),             # coprime, tiny windows
    ("quant_50_10", lambda: np.round(synth.accelerometer(13, 700)[0] * 8).astype(np.float32) / 8, 50, 10),  # ties for mode / zc
    ("exact_fit_100_100", lambda: synth.ppg(5, 1000), 100, 100),              # no overlap, no tail
    ("single_window", lambda: synth.ppg(6, 333), 333, 1),                     # nw == 1
]


def windows_fixture():
    out = {}
    for cname, build, W, S in CASES:
        x32 = np.ascontiguousarray(build(), dtype=np.float32)
        x = x32.astype(np.float64)
        out[cname + "/x"] = x32
        out[cname + "/ws"] = np.array([W, S], dtype=np.int64)
        for rname, f in REDUCERS.items():
            if rname == "mode" and cname not in ("quant_50_10", "acc_z_7_3"):
                continue
            try:
                r = rolling_apply(f)(x, W, S)
            except Exception as e:                      # pragma: no cover - report and skip
                print("SKIP", cname, rname, type(e).__name__, str(e)[:80])
                continue
            out[cname + "/" + rname] = np.asarray(r, dtype=np.float64)
        print(cname, "nw =", len(out[cname + "/mean"]))
    # view()
    xv = np.arange(23, dtype=np.float32)
    out["view/x"] = xv
    out["view/5_3"] = np.array(view(xv, 5, 3))
    # direct (un-rolled) calls on one window, incl. tuple-returning / array-returning ones
    w = synth.accelerometer(21, 257)[2].astype(np.float64)
    out["direct/x"] = w.astype(np.float32)
    w = out["direct/x"].astype(np.float64)
    out["direct/minmax"] = np.array(stats.minmax(w))
    out["direct/zero_crossings_0.9"] = np.asarray(timedom.zero_crossings(w, 0.9))
    out["direct/gradient"] = np.asarray(timedom.gradient(w))
    out["direct/hjorth_parameters"] = np.array(timedom.hjorth_parameters(w))
    out["direct/percentile_multi"] = np.asarray(np.percentile(w, [5, 50, 95]))
    # non-uniform windows
    rng = np.random.default_rng(5)
    idx = np.cumsum(rng.integers(1, 5, 400)).astype(np.int64) + 1000
    vals = rng.standard_normal(400)
    out["nonuniform/index"] = idx
    out["nonuniform/vals"] = vals
    out["nonuniform/indices_60_20"] = get_indices(idx, 60, 20)
    out["nonuniform/mean_60_20"] = nonuniform_rolling_apply(np.mean)(idx, vals, 60, 20)
    out["nonuniform/std_60_20_min5"] = nonuniform_rolling_apply(np.std, 5)(idx, vals, 3, 20)
    out["nonuniform/max_60_20"] = nonuniform_rolling_apply(np.max)(idx, vals, 60, 20)
    np.savez_compressed(os.path.join(HERE, "ref_windows.npz"), **out)


def spectral_fixture():
    out = {}
    for cname, x32, W, S, fs in [("acc", synth.accelerometer(7, 6137)[2], 500, 250, 50.0),
                                 ("ppg", synth.ppg(3, 6000), 1920, 64, 64.0),
                                 ("odd", synth.ppg(4, 900), 45, 20, 64.0)]:
        x = x32.astype(np.float64)
        spec = mhealth.fft.fft(view(x, W, S)[0])             # numpy fallback, 1 window
        out[cname + "/x"] = x32
        out[cname + "/wsf"] = np.array([W, S, fs])
        out[cname + "/fft0"] = np.asarray(spec)
        out[cname + "/ifft0"] = np.asarray(mhealth.fft.ifft(spec))
        nb = W // 2 + 1
        freqs = np.fft.rfftfreq(W, 1 / fs)
        psd = np.abs(np.fft.fft(view(x, W, S), axis=1)[:, :nb]) ** 2
        bands = [(0.5, 3.0), (3.0, 8.0), (0.0, 0.4)]
        out[cname + "/bands"] = np.array(bands)
        nw = psd.shape[0]
        bp = np.zeros((nw, len(bands)))
        rbp = np.zeros((nw, len(bands)))
        pk = np.zeros(nw)
        pk_all = np.zeros(nw)
        ent = np.zeros(nw)
        tot = np.zeros(nw)
        for i in range(nw):
            for j, (lo, hi) in enumerate(bands):
                bp[i, j] = hrv.power_band(psd[i], freqs, lo, hi)
                rbp[i, j] = hrv.relative_power_band(psd[i], freqs, lo, hi)
            pk[i] = density.peak_frequency(psd[i], freqs, 0.3, 12.0)
            pk_all[i] = density.peak_frequency(psd[i], freqs, None, None)
            ent[i] = information.entropy(psd[i])
            tot[i] = hrv.power_band(psd[i], freqs)
        out[cname + "/band_power"] = bp
        out[cname + "/rel_band_power"] = rbp
        out[cname + "/peak_frequency_0.3_12"] = pk
        out[cname + "/peak_frequency_all"] = pk_all
        out[cname + "/entropy"] = ent
        out[cname + "/total_power"] = tot
        print(cname, "nw =", nw)
    np.savez_compressed(os.path.join(HERE, "ref_spectral.npz"), **out)


def location_fixture():
    out = {}
    # the reference's own test vectors (tests/location/test_distance.py:7-13)
    pts = np.array([(0.1532, 86.675), (33.123, 21.541), (41.507483, -99.436554),
                    (38.504048, -98.315949), (51.5074, 0.1278), (41.3851, 2.1734)])
    out["points"] = pts
    # stale goldens as written in the reference's tests (2r = 12742.0) -- see BASELINE.md section 4
    out["stale/scalar_0_1"] = np.array(7704.777296228049)
    out["stale/elementwise"] = np.array([7704.77729623, 9756.94118642, 347.32834804,
                                         7275.82114826, 1136.28562666])
    out["stale/vector"] = np.array([7704.77729623, 15341.98217643, 15686.42408015,
                                    9755.32422594, 9537.84258146])
    lats, lons = pts[:, 0].copy(), pts[:, 1].copy()
    out["ref/scalar_0_1"] = np.array(distance.haversine(lats[0], lons[0], lats[1], lons[1]))
    out["ref/elementwise"] = distance.haversine_elementwise(lats[:-1], lons[:-1], lats[1:], lons[1:])
    out["ref/vector"] = distance.haversine_vector(lats[0], lons[0], lats[1:], lons[1:])
    out["ref/outer"] = distance.haversine_outer_product(lats, lons, lats, lons)
    # config-1 style trace (shortened) through the arr_* feature functions
    lat, lon, t, home = synth.gps(0, 3000)
    out["gps/n_period"] = np.array([3000, 60])
    out["gps/home"] = np.array(home)
    out["gps/successive_distance"] = features.arr_successive_distance(lat, lon)
    out["gps/distance_from_home"] = features.arr_distance_from_home(lat, lon, home)
    out["gps/proportion_home_stay_0.1"] = np.array(features.arr_proportion_home_stay(lat, lon, 0.1, home))
    out["gps/proportion_home_stay_5"] = np.array(features.arr_proportion_home_stay(lat, lon, 5.0, home))
    out["gps/location_variance"] = np.array(distribution.arr_location_variance(lat, lon))
    rng = np.random.default_rng(9)
    labels = rng.choice(np.array([-1, 0, 1, 2, 5, 9]), size=3000, p=[.3, .3, .2, .1, .05, .05])
    out["labels/x"] = labels.astype(np.int64)
    out["labels/num_clusters"] = np.array(distribution.num_clusters(labels))
    tot = distribution.cluster_totals(labels)
    out["labels/totals_keys"] = np.array(sorted(tot), dtype=np.int64)
    out["labels/totals_vals"] = np.array([tot[k] for k in sorted(tot)], dtype=np.int64)
    out["labels/entropy"] = np.array(distribution.cluster_entropy(labels))
    out["labels/normalized_entropy"] = np.array(distribution.normalized_cluster_entropy(labels))
    out["labels/normalized_entropy_n8"] = np.array(distribution.normalized_cluster_entropy(labels, 8))
    out["entropy/counts"] = np.array([5.0, 1.0, 0.0, 17.0, 2.5])
    out["entropy/value"] = np.array(information.entropy(out["entropy/counts"]))
    np.savez_compressed(os.path.join(HERE, "ref_location.npz"), **out)


def extra_fixture():
    """Accelerometer pre-stage (inertial/accelerometer.py), HRV time-domain metrics (heart/hrv.py:50-170) and the
    DataFrame forms of the location features (location/features.py:11-40, 56-68, 87-95)."""
    import pandas as pd
    from mhealth.inertial import accelerometer as acc
    out = {}
    a = synth.accelerometer(7, 5003)                       # float32 [3, n]
    out["acc/xyz"] = a
    out["acc/magnitude_f32"] = acc.magnitude(a[0], a[1], a[2])
    out["acc/roll_f32"] = acc.roll(a[1], a[2])
    out["acc/pitch_f32"] = acc.pitch(a[0], a[1], a[2])
    a64 = a.astype(np.float64)
    out["acc/magnitude_f64"] = acc.magnitude(a64[0], a64[1], a64[2])
    out["acc/roll_f64"] = acc.roll(a64[1], a64[2])
    out["acc/pitch_f64"] = acc.pitch(a64[0], a64[1], a64[2])
    out["acc/magnitude_dot_f64"] = np.array(acc.magnitude_dot(a64[0], a64[1], a64[2]))
    out["acc/scalars"] = np.array([acc.magnitude(1.0, 2.0, 2.0), acc.roll(1.0, 1.0), acc.pitch(1.0, 0.0, 1.0)])
    df = pd.DataFrame({"x": a64[0], "y": a64[1], "z": a64[2]})
    out["acc/df_magnitude"] = acc.magnitude(df).values
    out["acc/df_roll"] = acc.roll(df).values
    out["acc/df_pitch"] = acc.pitch(df).values
    # RR intervals (ms): 0.8 s +- respiration + noise, a few ectopic jumps
    rng = np.random.default_rng(4242)
    n = 4000
    rr = 800 + 60 * np.sin(np.arange(n) * 0.21) + rng.normal(0, 25, n)
    rr[rng.integers(0, n, 25)] += rng.normal(0, 150, 25)
    out["hrv/rr_ms"] = rr
    out["hrv/sdnn"] = np.array(hrv.sdnn(rr))
    out["hrv/pnn50"] = np.array(hrv.pnn50(rr, 'ms'))
    out["hrv/pnnx_20"] = np.array(hrv.pnnx(rr, 'ms', 20.0))
    out["hrv/pnn50_s"] = np.array(hrv.pnn50(rr / 1e3, 's'))
    out["hrv/rmssd"] = np.array(hrv.rmssd(rr))
    out["hrv/ssd"] = np.array(hrv.ssd(rr))
    out["hrv/sdsd"] = np.array(hrv.sdsd(rr))
    out["hrv/nni_to_ms"] = hrv.nni_to_ms(rr[:16] * 1e6, 'ns')
    out["hrv/poincare"] = np.array([hrv.csi_sd1(rr), hrv.csi_sd2(rr), hrv.lorenz_csi(rr), hrv.lorenz_cvi(rr),
                                    hrv.lorenz_mcsi(rr), hrv.csi_sd2(rr, 0.5)])
    g = a64[2][:2001].copy()
    out["td/x"] = g
    out["td/gradient_f64"] = timedom.gradient(g)
    out["td/gradient_f32"] = timedom.gradient(g.astype(np.float32))
    out["td/zero_crossings_0"] = timedom.zero_crossings(g - g.mean(), 0.0)
    out["td/zero_crossings_th"] = timedom.zero_crossings(g - g.mean(), 0.05)
    from mhealth.heart import ppg as rppg
    sig = synth.ppg(3, 4000).astype(np.float64)
    out["ppg/x"] = sig
    out["ppg/slope_sum_9"] = rppg.slope_sum(sig, 9)
    out["ppg/slope_sum_1"] = rppg.slope_sum(sig, 1)
    # location DataFrame forms
    lat, lon, t, _ = synth.gps(5, 3000)
    gdf = pd.DataFrame({"latitude": lat, "longitude": lon}, index=pd.to_datetime(t, unit="s"))
    out["gps/lat"], out["gps/lon"], out["gps/t"] = lat, lon, t
    home = features.determine_home_coords(gdf)
    out["gps/home"] = np.array(home)
    out["gps/distance_from_home"] = np.asarray(features.distance_from_home(gdf))
    out["gps/proportion_home_stay_0.5"] = np.array(features.proportion_home_stay(gdf, 0.5))
    # features.successive_distance(df) is NOT recorded: with a datetime index its ``dist[0] = 0`` on a Series appends a
    # label under pandas 3 (3001 garbage values); the array form is pinned in ref_location.npz instead
    out["gps/arr_successive_distance"] = features.arr_successive_distance(lat, lon)
    out["gps/location_variance"] = np.array(distribution.location_variance(gdf))
    np.savez_compressed(os.path.join(HERE, "ref_extra.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["windows", "spectral", "location", "extra"]
    if "extra" in which:
        extra_fixture()
    if "windows" in which:
        windows_fixture()
    if "spectral" in which:
        spectral_fixture()
    if "location" in which:
        location_fixture()
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
