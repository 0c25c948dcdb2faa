"""Location distance features -- drop-in for the array forms of ``mhealth.location.features``
(reference src/mhealth/location/features.py:43-53, 71-84, 98-113).  The pandas DataFrame forms
(features.py:11-40, 56-68, 87-95) are thin wrappers around them."""

import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr
from . import distance
from .distance import _dev


def determine_home_coords(df, start_time='23:00', end_time='06:00'):
    """Median latitude / longitude during night-time (features.py:11-24): pandas glue, no kernel involved."""
    night = df[['latitude', 'longitude']].between_time(start_time, end_time)
    lat, lon = night.median().values
    return (lat, lon)


def distance_from_home(df, home_coords=None):
    """DataFrame form of ``arr_distance_from_home`` -> pd.Series 'home_distance' (features.py:27-40)."""
    from ..util.deps import pd
    if home_coords is None:
        home_coords = determine_home_coords(df)
    out = arr_distance_from_home(df['latitude'].values, df['longitude'].values, home_coords)
    return pd.Series(out, index=df.index, name='home_distance')


def proportion_home_stay(df, limit=0.1, home_coords=None):
    """DataFrame form of ``arr_proportion_home_stay`` (features.py:56-68)."""
    if home_coords is None:
        home_coords = determine_home_coords(df)
    return arr_proportion_home_stay(df['latitude'].values, df['longitude'].values, limit, home_coords)


def successive_distance(df):
    """DataFrame form of ``arr_successive_distance`` -> pd.Series (features.py:87-95)."""
    from ..util.deps import pd
    out = arr_successive_distance(df['latitude'].values, df['longitude'].values)
    return pd.Series(out, index=df.index, name='latitude')


def arr_distance_from_home(latitude, longitude, home_coords):
    """Distance (km) of every point from ``home_coords`` = (lat, lon) (features.py:43-53)."""
    lat, lon = home_coords
    return distance.haversine_vector(lat, lon, latitude, longitude)


def arr_proportion_home_stay(latitude, longitude, limit, home_coords):
    """Share of points strictly within ``limit`` km of home (features.py:71-84).  One fused pass:
    the distances are not materialised on the host."""
    torch = require_cuda()
    la, lo = _dev(latitude), _dev(longitude)
    n = la.shape[0]
    if n == 0:
        raise ZeroDivisionError("arr_proportion_home_stay of an empty trace")
    rows = segment_rows(la, lo, torch.zeros(n, dtype=torch.int64, device=la.device), [0, n], [home_coords], limit)
    return float(rows[0, 5]) / n


def arr_successive_distance(latitude, longitude):
    """Distance between successive points; the first distance is 0; inputs untouched
    (features.py:98-113)."""
    torch = require_cuda()
    la, lo = _dev(latitude), _dev(longitude)
    if la.shape != lo.shape or la.dim() != 1:
        raise ValueError("arr_successive_distance: two 1-D arrays of equal length")
    out = torch.empty_like(la)
    L.check(L.load().mhb_successive_distance(la.data_ptr(), lo.data_ptr(), None, 0, la.shape[0], out.data_ptr(),
                                             _stream_ptr(torch)), "arr_successive_distance")
    return out if isinstance(latitude, torch.Tensor) else out.cpu().numpy()


def segment_rows(lat, lon, t, offsets, home, limit=0.1, stay_dist_km=0.2, stay_min_seconds=1800, labels=False):
    """Per-segment (subject-day) feature rows, float64 [n_segments, 11], columns ``_lib.SEG_COLUMNS``:
    the reference's per-trace features evaluated on every [offsets[k], offsets[k+1]) slice in one
    launch, plus the radius-of-gyration / stay-point extensions (no reference implementation;
    defined by oracle/location_ext.py).  ``home`` is [n_segments, 2] (lat, lon)."""
    torch = require_cuda()
    la, lo = _dev(lat), _dev(lon)
    tt = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(t, dtype=np.int64)))
    tt = tt.to(device="cuda", dtype=torch.int64).contiguous()
    offs = torch.as_tensor(np.asarray(offsets, dtype=np.int64)).cuda() if not isinstance(offsets, torch.Tensor) \
        else offsets.to(device="cuda", dtype=torch.int64).contiguous()
    ns = offs.shape[0] - 1
    hm = _dev(np.asarray(home, dtype=np.float64).reshape(-1, 2) if not isinstance(home, torch.Tensor) else home)
    if hm.shape[0] != ns:
        raise ValueError("home must have one (lat, lon) row per segment")
    rows = torch.empty((ns, len(L.SEG_COLUMNS)), dtype=torch.float64, device=la.device)
    lab = torch.empty(la.shape[0], dtype=torch.int64, device=la.device) if labels else None
    L.check(L.load().mhb_location_segments(la.data_ptr(), lo.data_ptr(), tt.data_ptr(), offs.data_ptr(), ns,
                                           hm.data_ptr(), float(limit), float(stay_dist_km), int(stay_min_seconds),
                                           rows.data_ptr(), lab.data_ptr() if labels else None, _stream_ptr(torch)),
            "location_segments")
    if isinstance(lat, torch.Tensor):
        return (rows, lab) if labels else rows
    return (rows.cpu().numpy(), lab.cpu().numpy()) if labels else rows.cpu().numpy()


def radius_of_gyration(latitude, longitude):
    """[extension] sqrt(mean haversine(p_i, centroid)^2), km."""
    n = len(latitude)
    rows = segment_rows(latitude, longitude, np.zeros(n, dtype=np.int64), [0, n], [(0.0, 0.0)])
    return float(rows[0, 3])


def stay_points(latitude, longitude, t, dist_km=0.2, min_seconds=1800):
    """[extension] stay-point label of every point (-1 = none); see oracle/location_ext.py."""
    n = len(latitude)
    _, lab = segment_rows(latitude, longitude, t, [0, n], [(0.0, 0.0)], 0.1, dist_km, min_seconds, labels=True)
    return lab
