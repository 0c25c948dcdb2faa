"""Location-trace features of the reference, restated.  TEST INFRASTRUCTURE ONLY.

Reference: ``src/mhealth/location/distance.py`` (haversine family, :4-59),
``location/features.py`` (:43-53, :71-84, :98-113), ``location/distribution.py``
(:28-39, :58-102).  All arithmetic is float64, as in the reference's gufunc signatures.
"""
import math

import numpy as np
from numba import njit

from . import reducers as R

EARTH_DIAMETER_KM = 12742.018     # distance.py:8,18 -- 2 * 6371.009
_DEG = math.pi / 180.0


@njit(cache=True)
def haversine(lat1, lon1, lat2, lon2):
    # distance.py:4-19: every input converted with np.radians BEFORE the subtraction;
    # 2r * asin(sqrt(sin^2(dlat/2) + cos(lat1) cos(lat2) sin^2(dlon/2)))
    a1 = np.radians(lat1)
    a2 = np.radians(lat2)
    o1 = np.radians(lon1)
    o2 = np.radians(lon2)
    s_lat = np.sin((a2 - a1) / 2.0)
    s_lon = np.sin((o2 - o1) / 2.0)
    h = s_lat ** 2 + (np.cos(a1) * np.cos(a2) * s_lon ** 2)
    return 12742.018 * np.arcsin(np.sqrt(h))


@njit(cache=True)
def haversine_elementwise(lat1, lon1, lat2, lon2):
    # distance.py:22-33, gufunc (n),(n),(n),(n)->(n)
    n = lat1.shape[0]
    out = np.empty(n)
    for i in range(n):
        out[i] = haversine(lat1[i], lon1[i], lat2[i], lon2[i])
    return out


@njit(cache=True)
def haversine_vector(lat, lon, latcol, loncol):
    # distance.py:36-44, gufunc (),(),(n),(n)->(n)
    n = latcol.shape[0]
    out = np.empty(n)
    for i in range(n):
        out[i] = haversine(lat, lon, latcol[i], loncol[i])
    return out


@njit(cache=True)
def haversine_outer_product(lat1, lon1, lat2, lon2):
    # distance.py:47-59, gufunc (n),(n),(m),(m)->(n,m)
    n = lat1.shape[0]
    m = lat2.shape[0]
    out = np.empty((n, m))
    for i in range(n):
        for j in range(m):
            out[i, j] = haversine(lat1[i], lon1[i], lat2[j], lon2[j])
    return out


def arr_successive_distance(latitude, longitude):
    """features.py:98-113: dist[0] = 0, dist[i] = haversine(p[i-1], p[i]); inputs untouched."""
    lat = np.ascontiguousarray(latitude, dtype=np.float64)
    lon = np.ascontiguousarray(longitude, dtype=np.float64)
    out = np.zeros(lat.shape[0])
    if lat.shape[0] > 1:
        out[1:] = haversine_elementwise(lat[:-1], lon[:-1], lat[1:], lon[1:])
    return out


def arr_distance_from_home(latitude, longitude, home_coords):
    """features.py:43-53."""
    hlat, hlon = home_coords
    return haversine_vector(float(hlat), float(hlon),
                            np.ascontiguousarray(latitude, dtype=np.float64),
                            np.ascontiguousarray(longitude, dtype=np.float64))


def arr_proportion_home_stay(latitude, longitude, limit, home_coords):
    """features.py:71-84: (d < limit).sum() / n -- strict inequality, integer count."""
    d = arr_distance_from_home(latitude, longitude, home_coords)
    return (d < limit).sum() / len(d)


def home_stay_count(latitude, longitude, limit, home_coords):
    """The integer numerator of arr_proportion_home_stay (bit-exact contract)."""
    return int((arr_distance_from_home(latitude, longitude, home_coords) < limit).sum())


def arr_location_variance(latitude, longitude):
    """distribution.py:28-39: var(lat) + var(lon), two-pass population variances (no log)."""
    return R.w_var(np.ascontiguousarray(latitude, dtype=np.float64)) + \
        R.w_var(np.ascontiguousarray(longitude, dtype=np.float64))


def num_clusters(cluster_labels):
    """distribution.py:58-65: number of distinct labels; noise (-1) counts as a label."""
    return len(np.unique(cluster_labels))


def cluster_totals(cluster_labels):
    """distribution.py:68-76: {label: occurrences}."""
    labs, cnts = np.unique(cluster_labels, return_counts=True)
    return {int(c): int(n) for c, n in zip(labs, cnts)}


def cluster_entropy(cluster_labels):
    """distribution.py:79-89: entropy of the label counts (sorted-label order)."""
    cnts = np.unique(cluster_labels, return_counts=True)[1]
    return R.entropy(cnts.astype(np.float64))


def normalized_cluster_entropy(cluster_labels, n_clusters=None):
    """distribution.py:92-102: entropy / ln(n_clusters)."""
    if n_clusters is None:
        n_clusters = len(np.unique(cluster_labels))
    return cluster_entropy(cluster_labels) / np.log(n_clusters)
