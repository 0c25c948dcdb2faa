"""PPG helpers -- drop-in for the window reduction of ``mhealth.heart.ppg`` (reference src/mhealth/heart/ppg.py).

Only ``slope_sum`` (:28-42), a hop-1 sliding-window sum, is a data-parallel window reduction; the Butterworth
band-pass and the sequential adaptive-threshold decision rule of ``pulse_onset_physionet`` (:11-25, 45-93) are out of
scope."""
import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr


def slope_sum(x, w: int):
    """Sum of the derivative of a sliding windowed signal: ``out[i] = sum(diff(x)[i-w:i])`` for ``w <= i < len(x)-1``,
    zero elsewhere; float64 output (ppg.py:28-42)."""
    torch = require_cuda()
    was_torch = isinstance(x, torch.Tensor)
    if was_torch:
        t = x.to(device="cuda")
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        t = t.contiguous().reshape(-1)
    else:
        a = np.asarray(x)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.array(a, order="C", copy=True).reshape(-1)).cuda()
    n = t.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=t.device)
    L.check(L.load().mhb_slope_sum(1 if t.dtype == torch.float64 else 0, t.data_ptr(), n, int(w), out.data_ptr(),
                                   _stream_ptr(torch)), "slope_sum")
    return out if was_torch else out.cpu().numpy()
