"""EXTENSIONS with no reference implementation -- **parity unpinned**.

TEST INFRASTRUCTURE ONLY.  BASELINE.json's north_star names radius of gyration and
stay-point / cluster assignment; callumstew/pymhealth has neither.  They are
specified here, in the style of ``src/mhealth/location/features.py:43-113`` (numba, float64,
haversine from ``location/distance.py:4-19``), and this file is the definition the CUDA
kernels are checked against.

Definitions
-----------
radius_of_gyration(lat, lon)
    sqrt(mean_i haversine(p_i, c)^2) in km, c = (mean(lat), mean(lon)) in degrees
    (Gonzalez et al. 2008, with the great-circle distance of the reference).
stay_points(lat, lon, t, dist_km, min_dur)
    Anchor scan (Li et al. 2008, jump variant): anchor i; j runs forward while
    haversine(p_i, p_j) <= dist_km; the run [i, j) is a stay point when
    t[j-1] - t[i] >= min_dur, and gets the next label 0, 1, 2, ...; otherwise its points are
    labelled -1 (noise, the convention of location/distribution.py:6-8); the next anchor is j.
segment_features(lat, lon, t, offsets, home, limit, dist_km, min_dur)
    one row per [offsets[k], offsets[k+1]) segment (a subject-day), columns SEG_COLUMNS.
"""
import math

import numpy as np
from numba import njit

from .location import haversine
from . import reducers as R

SEG_COLUMNS = ("n_points", "total_distance", "location_variance", "radius_of_gyration",
               "max_home_distance", "home_stay_count", "proportion_home_stay",
               "n_stay_points", "n_labels", "label_entropy", "normalized_label_entropy")


@njit(cache=True)
def radius_of_gyration(lat, lon):
    n = lat.shape[0]
    clat = R.w_mean(lat)
    clon = R.w_mean(lon)
    acc = 0.0
    for i in range(n):
        d = haversine(lat[i], lon[i], clat, clon)
        acc += d * d
    return math.sqrt(acc / n)


@njit(cache=True)
def stay_points(lat, lon, t, dist_km, min_dur):
    n = lat.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    i = 0
    k = 0
    while i < n:
        j = i + 1
        while j < n and haversine(lat[i], lon[i], lat[j], lon[j]) <= dist_km:
            j += 1
        if t[j - 1] - t[i] >= min_dur:
            for q in range(i, j):
                labels[q] = k
            k += 1
        i = j
    return labels


def segment_features(lat, lon, t, offsets, home, limit, dist_km, min_dur):
    """float64[n_segments, len(SEG_COLUMNS)]; ``home`` is float64[n_segments, 2] (lat, lon)."""
    from . import location as L
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.int64)
    ns = len(offsets) - 1
    out = np.zeros((ns, len(SEG_COLUMNS)))
    labels_all = np.full(lat.shape[0], -1, dtype=np.int64)
    for k in range(ns):
        a, b = int(offsets[k]), int(offsets[k + 1])
        n = b - a
        out[k, 0] = n
        if n == 0:
            out[k, 1:] = np.nan
            out[k, 5] = 0
            out[k, 7] = 0
            out[k, 8] = 0
            continue
        la, lo, tt = lat[a:b], lon[a:b], t[a:b]
        out[k, 1] = L.arr_successive_distance(la, lo).sum()
        out[k, 2] = L.arr_location_variance(la, lo)
        out[k, 3] = radius_of_gyration(la, lo)
        dh = L.arr_distance_from_home(la, lo, (home[k, 0], home[k, 1]))
        out[k, 4] = dh.max()
        cnt = int((dh < limit).sum())
        out[k, 5] = cnt
        out[k, 6] = cnt / n
        lab = stay_points(la, lo, tt, dist_km, min_dur)
        labels_all[a:b] = lab
        out[k, 7] = lab.max() + 1
        nl = L.num_clusters(lab)
        out[k, 8] = nl
        out[k, 9] = L.cluster_entropy(lab)
        out[k, 10] = out[k, 9] / np.log(nl) if nl > 1 else np.nan
    return out, labels_all
