"""FFT provider -- drop-in for ``mhealth.fft`` (reference src/mhealth/fft/__init__.py:1-7,
fft/_fft.py:18-48): ``fft(a)`` unnormalised forward DFT, ``ifft(a)`` backward / n, complex128 out,
length == len(a) exactly (no padding).  The reference binds FFTW through CFFI (one plan per call,
fft/_fftw_binder.py:11-17) or falls back to numpy; here the rows are transformed by the
shared-memory mixed-radix FFT of libmhb200 (no cuFFT).  A 2-D input is transformed along its last
axis (one launch for all rows).  Lengths with a prime factor > 31 raise NotImplementedError."""
import numpy as np

from .. import _lib as L
from ..engine import require_cuda, _stream_ptr


def _transform(a, direction):
    torch = require_cuda()
    lib = L.load()
    a = np.asarray(a)
    if a.ndim == 0:
        raise ValueError("fft of a scalar")
    cplx = np.iscomplexobj(a)
    src = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
    n = src.shape[-1]
    rows = int(np.prod(src.shape[:-1])) if src.ndim > 1 else 1
    if n == 0:
        raise ValueError("fft of an empty array")
    flat = src.view(np.float64).reshape(-1) if cplx else src.reshape(-1)
    d = torch.from_numpy(flat).cuda()
    out = torch.empty(rows * n * 2, dtype=torch.float64, device=d.device)
    L.check(lib.mhb_fft_c128(d.data_ptr(), 1 if cplx else 0, rows, n, direction, out.data_ptr(), _stream_ptr(torch)), "fft")
    return out.cpu().numpy().view(np.complex128).reshape(src.shape)


def fft(a):
    return _transform(a, -1)        # FFTW_FORWARD (fft/_fft.py:8)


def ifft(a):
    return _transform(a, +1)        # FFTW_BACKWARD, scaled by 1/n (fft/_fft.py:46-48)
