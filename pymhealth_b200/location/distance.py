"""Haversine distances -- drop-in for ``mhealth.location.distance``
(reference src/mhealth/location/distance.py).  Degrees in, kilometres out, 2r = 12742.018,
float64 throughout (the reference's gufuncs have float64-only signatures; float32 / integer
inputs are promoted, as numpy does for them)."""

import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr


def _dev(a):
    torch = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(np.array(a, dtype=np.float64, order="C", copy=True)).cuda()


def haversine(lat1, lon1, lat2, lon2):
    """Haversine distance in km between two points given in degrees (distance.py:4-19)."""
    out = haversine_elementwise(np.array([lat1], dtype=np.float64), np.array([lon1], dtype=np.float64),
                                np.array([lat2], dtype=np.float64), np.array([lon2], dtype=np.float64))
    return float(out[0])


def haversine_elementwise(lat1, lon1, lat2, lon2):
    """Elementwise distance between two point vectors, (n),(n),(n),(n)->(n) (distance.py:22-33)."""
    torch = require_cuda()
    a, b, c, d = (_dev(v) for v in (lat1, lon1, lat2, lon2))
    if not (a.shape == b.shape == c.shape == d.shape) or a.dim() != 1:
        raise ValueError("haversine_elementwise: four 1-D arrays of equal length")
    out = torch.empty_like(a)
    L.check(L.load().mhb_haversine_elementwise(a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), a.shape[0],
                                               out.data_ptr(), _stream_ptr(torch)), "haversine_elementwise")
    return out if isinstance(lat1, torch.Tensor) else out.cpu().numpy()


def haversine_vector(lat1, lon1, latcol, loncol):
    """Distance between a fixed point and vectors of points, (),(),(n),(n)->(n) (distance.py:36-44)."""
    torch = require_cuda()
    a, b = _dev(latcol), _dev(loncol)
    if a.shape != b.shape or a.dim() != 1:
        raise ValueError("haversine_vector: two 1-D arrays of equal length")
    out = torch.empty_like(a)
    L.check(L.load().mhb_haversine_vector(float(lat1), float(lon1), a.data_ptr(), b.data_ptr(), a.shape[0],
                                          out.data_ptr(), _stream_ptr(torch)), "haversine_vector")
    return out if isinstance(latcol, torch.Tensor) else out.cpu().numpy()


def haversine_outer_product(lat1, lon1, lat2, lon2):
    """All-pairs distances, (n),(n),(m),(m)->(n,m) (distance.py:47-59)."""
    torch = require_cuda()
    a, b, c, d = (_dev(v) for v in (lat1, lon1, lat2, lon2))
    if a.shape != b.shape or c.shape != d.shape or a.dim() != 1 or c.dim() != 1:
        raise ValueError("haversine_outer_product: (n),(n),(m),(m)")
    out = torch.empty((a.shape[0], c.shape[0]), dtype=torch.float64, device=a.device)
    L.check(L.load().mhb_haversine_outer(a.data_ptr(), b.data_ptr(), a.shape[0], c.data_ptr(), d.data_ptr(),
                                         c.shape[0], out.data_ptr(), _stream_ptr(torch)), "haversine_outer_product")
    return out if isinstance(lat1, torch.Tensor) else out.cpu().numpy()
