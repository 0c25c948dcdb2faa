#!/usr/bin/env python3
"""Tiny invocation of every kernel family (for compute-sanitizer memcheck / racecheck)."""
import functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymhealth_b200 import synth, spectral as SP
from pymhealth_b200.util import rolling_apply
from pymhealth_b200.generic import stats, timedom
from pymhealth_b200.location import features, distribution, distance
from pymhealth_b200 import fft as F

x = synth.accelerometer(3, 9137)
for W, S in ((500, 250), (64, 48), (1920, 64), (33, 7)):
    rolling_apply([np.mean, np.std, stats.skewness, stats.kurtosis, timedom.zero_crossing_count, timedom.line_length,
                   np.median, functools.partial(np.percentile, q=90), stats.mode, timedom.hjorth_mobility])(x[2], W, S)
rolling_apply([SP.total_power(50.0), SP.band_power(50.0, 0.5, 3), SP.peak_frequency(50.0, 0.3, 12), SP.spectral_entropy(50.0)])(x[1], 500, 250)
rolling_apply([SP.total_power(50.0), SP.spectral_entropy(50.0)])(x[1], 90, 30)
SP.window_psd(x[0], 500, 250, 50.0)
F.fft(x[0][:500].astype(np.float64)); F.ifft(np.fft.fft(x[0][:60]))
lat, lon, t, home = synth.gps(0, 4000, 60)
features.segment_rows(lat, lon, t, [0, 1440, 2880, 2880, 4000], [home] * 4, labels=True)
features.arr_successive_distance(lat, lon); features.arr_proportion_home_stay(lat, lon, 0.1, home)
distance.haversine_outer_product(lat[:10], lon[:10], lat[:7], lon[:7])
distribution.cluster_entropy(np.array([-1, 0, 0, 2, 5, 5, 5])); distribution.arr_location_variance(lat, lon)
print("sanitize target ok")
