"""Location distribution features -- drop-in for the array forms of
``mhealth.location.distribution`` (reference src/mhealth/location/distribution.py:28-39, 58-102).
``cluster_locations`` (third-party HDBSCAN, distribution.py:42-55) is not ported: use
``location.features.stay_points`` for label assignment."""

import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr, window_table, Feature

_MAX_LABEL_RANGE = 1 << 26


def location_variance(df):
    """DataFrame form of ``arr_location_variance`` (distribution.py:16-25)."""
    return arr_location_variance(df['latitude'].values, df['longitude'].values)


def arr_location_variance(latitude, longitude):
    """var(latitude) + var(longitude), population variances (distribution.py:28-39)."""
    la = np.ascontiguousarray(np.asarray(latitude, dtype=np.float64))
    lo = np.ascontiguousarray(np.asarray(longitude, dtype=np.float64))
    if la.shape != lo.shape or la.ndim != 1 or la.shape[0] == 0:
        raise ValueError("arr_location_variance: two non-empty 1-D arrays of equal length")
    tab = window_table(np.stack([la, lo]), la.shape[0], 1, [Feature("stream", L.F_VAR, (), "var")])
    return float(tab[0, 0, 0] + tab[1, 0, 0])


def _label_stats(cluster_labels, n_clusters=None, want_counts=False):
    torch = require_cuda()
    lib = L.load()
    lab = np.ascontiguousarray(np.asarray(cluster_labels).astype(np.int64, copy=False).ravel())
    if lab.shape[0] == 0:
        raise ValueError("empty label array")
    d = torch.from_numpy(lab).cuda()
    mm = torch.empty(2, dtype=torch.int64, device=d.device)
    L.check(lib.mhb_minmax_i64(d.data_ptr(), d.shape[0], mm.data_ptr(), _stream_ptr(torch)), "label range")
    lo, hi = (int(v) for v in mm.cpu())
    if hi - lo + 1 > _MAX_LABEL_RANGE:
        raise NotImplementedError("cluster labels span %d values; the dense label histogram supports up to %d "
                                  "(labels are small integers from -1 upwards, distribution.py:6-8)"
                                  % (hi - lo + 1, _MAX_LABEL_RANGE))
    counts = torch.empty(hi - lo + 1, dtype=torch.int64, device=d.device)
    out3 = torch.empty(3, dtype=torch.float64, device=d.device)
    L.check(lib.mhb_label_stats(d.data_ptr(), d.shape[0], lo, hi, int(n_clusters) if n_clusters else 0,
                                counts.data_ptr(), out3.data_ptr(), _stream_ptr(torch)), "label_stats")
    res = out3.cpu().numpy()
    if want_counts:
        c = counts.cpu().numpy()
        keep = np.nonzero(c)[0]
        return res, {int(k + lo): int(c[k]) for k in keep}
    return res


def num_clusters(cluster_labels):
    """Number of distinct labels; the noise label -1 counts (distribution.py:58-65)."""
    return int(_label_stats(cluster_labels)[0])


def cluster_totals(cluster_labels):
    """{label: occurrences} (distribution.py:68-76)."""
    return _label_stats(cluster_labels, want_counts=True)[1]


def cluster_entropy(cluster_labels):
    """Shannon entropy of the label counts (distribution.py:79-89)."""
    return float(_label_stats(cluster_labels)[1])


def normalized_cluster_entropy(cluster_labels, n_clusters=None):
    """entropy / ln(n_clusters) (distribution.py:92-102)."""
    return float(_label_stats(cluster_labels, n_clusters)[2])
