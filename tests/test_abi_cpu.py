"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/mhb200.h declares, and the product tree never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "mhb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mhb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pymhealth_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libmhb200.so does not export %s" % n
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().mhb_abi_version() == 1
    assert _lib.load().mhb_n_windows(4320000, 500, 250) == 17279
    assert _lib.load().mhb_n_windows(5, 8, 2) == 0


def test_argument_errors_do_not_need_a_gpu():
    from pymhealth_b200 import _lib as L
    lib = L.load()
    geom = L.MhbWindows(1, 100, 100, 0, 1)                      # wsize = 0
    tab = L.MhbTable(None, 0, 0, 0, 0)
    st = lib.mhb_window_stats_f32(None, ctypes.byref(geom), L.i32_array([0]), 1, 0.0, ctypes.byref(tab), None)
    assert st == -1 and b"wsize" in lib.mhb_last_error()
    geom = L.MhbWindows(1, 100, 100, 10, 5)
    st = lib.mhb_window_stats_f32(None, ctypes.byref(geom), L.i32_array([99]), 1, 0.0, ctypes.byref(tab), None)
    assert st == -2                                              # unknown feature id
    st = lib.mhb_fft_c128(None, 0, 1, 2 * 37, -1, None, None)    # prime factor 37 > 31
    assert st in (-1, -3)
    with pytest.raises(ValueError):
        L.check(-1, "x")
    with pytest.raises(NotImplementedError):
        L.check(-3, "x")


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pymhealth_b200 import _lib as L
    from pymhealth_b200.util import rolling_apply
    with pytest.raises(L.MhbError):
        rolling_apply(np.mean)(np.zeros(100, np.float32), 10, 5)
    from pymhealth_b200.location import distance
    with pytest.raises(L.MhbError):
        distance.haversine(0.0, 0.0, 1.0, 1.0)
    # every other public module of the path: the same loud failure, never a numpy substitute
    from pymhealth_b200 import fft, spectral
    from pymhealth_b200.generic import timedom
    from pymhealth_b200.heart import hrv, ppg
    from pymhealth_b200.inertial import accelerometer as acc
    from pymhealth_b200.location import features, distribution
    from pymhealth_b200.util.windows import nonuniform_rolling_apply, get_indices
    x = np.arange(64.0)
    for call in (lambda: acc.magnitude(x, x, x), lambda: acc.rolling_magnitude(np.mean, 8, 4)(x, x, x),
                 lambda: hrv.sdnn(x + 800.0), lambda: hrv.rmssd(x + 800.0), lambda: ppg.slope_sum(x, 5),
                 lambda: timedom.gradient(x), lambda: timedom.zero_crossings(x), lambda: fft.fft(x),
                 lambda: spectral.window_psd(x.astype(np.float32), 16, 8), lambda: get_indices(np.arange(64), 8, 4),
                 lambda: nonuniform_rolling_apply(np.mean)(np.arange(64), x, 8, 4),
                 lambda: features.arr_successive_distance(x, x), lambda: distribution.cluster_entropy(np.arange(8)),
                 lambda: rolling_apply(spectral.spectral_entropy(50.0))(x.astype(np.float32), 16, 8)):
        with pytest.raises(L.MhbError):
            call()


def test_unknown_reducers_are_rejected_not_run_on_cpu():
    from pymhealth_b200.util import rolling_apply
    for bad in (lambda w: w.mean(), np.nanmean, len):
        with pytest.raises(NotImplementedError):
            rolling_apply(bad)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|^\s*from\s+\.\.?\s*oracle", re.M)
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymhealth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if pat.search(txt) or "oracle/_ref" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
