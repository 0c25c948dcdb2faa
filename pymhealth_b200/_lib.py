"""ctypes binding of libmhb200.so (the C ABI declared in include/mhb200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pymhealth_b200/csrc``.
There is NO fallback: if the shared object is missing or a CUDA device is absent, every
compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MHB200_LIB") or os.path.join(_HERE, "libmhb200.so")      # MHB200_LIB: a development build

# ---- feature ids (mirror of include/mhb200.h)
F_MEAN, F_VAR, F_STD, F_MIN, F_MAX, F_DRANGE, F_SKEWNESS, F_KURTOSIS, F_KURTOSIS_EXCESS, \
    F_COEFF_VAR, F_ZERO_CROSSINGS, F_LINE_LENGTH, F_HJORTH_ACTIVITY, F_SUM = range(14)
F_MEDIAN, F_PERCENTILE, F_IQR, F_MODE, F_HJORTH_MOBILITY, F_HJORTH_COMPLEXITY = range(32, 38)
S_TOTAL_POWER, S_BAND_POWER, S_REL_BAND_POWER, S_PEAK_FREQUENCY, S_PEAK_BIN, S_ENTROPY = range(6)
SEG_COLUMNS = ("n_points", "total_distance", "location_variance", "radius_of_gyration",
               "max_home_distance", "home_stay_count", "proportion_home_stay",
               "n_stay_points", "n_labels", "label_entropy", "normalized_label_entropy")

STREAMING = frozenset(range(14))
ORDER = frozenset(range(32, 38))


class MhbWindows(C.Structure):
    _fields_ = [("n_series", C.c_int64), ("series_len", C.c_int64), ("series_stride", C.c_int64),
                ("wsize", C.c_int32), ("wstep", C.c_int32)]


class MhbTable(C.Structure):
    _fields_ = [("out", C.c_void_p), ("out_f32", C.c_int32), ("series_stride", C.c_int64),
                ("window_stride", C.c_int64), ("column_stride", C.c_int64)]


class MhbError(RuntimeError):
    pass


_lib = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/mhb200.h declares
SIGNATURES = {
    "mhb_abi_version": (C.c_int32, []),
    "mhb_last_error": (C.c_char_p, []),
    "mhb_n_windows": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "mhb_window_stats_f32": (C.c_int32, [_vp, C.POINTER(MhbWindows), _i32p, C.c_int32, C.c_double,
                                         C.POINTER(MhbTable), _vp]),
    "mhb_window_stats_f64": (C.c_int32, [_vp, C.POINTER(MhbWindows), _i32p, C.c_int32, C.c_double,
                                         C.POINTER(MhbTable), _vp]),
    "mhb_window_stats_magnitude_f32": (C.c_int32, [_vp, _vp, _vp, C.POINTER(MhbWindows), _i32p, C.c_int32, C.c_double,
                                                   C.POINTER(MhbTable), _vp]),
    "mhb_window_stats_magnitude_f64": (C.c_int32, [_vp, _vp, _vp, C.POINTER(MhbWindows), _i32p, C.c_int32, C.c_double,
                                                   C.POINTER(MhbTable), _vp]),
    "mhb_window_order_f32": (C.c_int32, [_vp, C.POINTER(MhbWindows), _i32p, _f64p, C.c_int32,
                                         C.POINTER(MhbTable), _vp]),
    "mhb_window_order_f64": (C.c_int32, [_vp, C.POINTER(MhbWindows), _i32p, _f64p, C.c_int32,
                                         C.POINTER(MhbTable), _vp]),
    "mhb_get_indices_i64": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp]),
    "mhb_get_indices_f64": (C.c_int32, [_vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int64, _vp, _vp]),
    "mhb_segment_stats_f32": (C.c_int32, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64, _i32p, C.c_int32,
                                          C.c_double, C.POINTER(MhbTable), _vp]),
    "mhb_segment_stats_f64": (C.c_int32, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64, _i32p, C.c_int32,
                                          C.c_double, C.POINTER(MhbTable), _vp]),
    "mhb_segment_order_f32": (C.c_int32, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _i32p, _f64p,
                                          C.c_int32, C.POINTER(MhbTable), _vp]),
    "mhb_segment_order_f64": (C.c_int32, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _i32p, _f64p,
                                          C.c_int32, C.POINTER(MhbTable), _vp]),
    "mhb_window_spectral_f32": (C.c_int32, [_vp, C.POINTER(MhbWindows), C.c_double, _i32p, _f64p, C.c_int32,
                                            C.POINTER(MhbTable), _vp]),
    "mhb_window_features_f32": (C.c_int32, [_vp, C.POINTER(MhbWindows), _i32p, C.c_int32, C.c_double, C.POINTER(MhbTable),
                                            C.c_double, _i32p, _f64p, C.c_int32, C.POINTER(MhbTable), _vp]),
    "mhb_widen_i16_f32": (C.c_int32, [_vp, C.c_int64, C.c_float, _vp, _vp]),
    "mhb_psd_reduce_f64": (C.c_int32, [_vp, _vp, C.c_int64, C.c_int32, _i32p, _f64p, C.c_int32, _vp, _vp]),
    "mhb_fft_c128": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _vp, _vp]),
    "mhb_window_psd_f32": (C.c_int32, [_vp, C.POINTER(MhbWindows), _vp, C.c_int32, _vp]),
    "mhb_accel_elementwise": (C.c_int32, [C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_int64, _vp, _vp]),
    "mhb_accel_sumsq_workspace": (C.c_int64, [C.c_int64]),
    "mhb_accel_magnitude_dot": (C.c_int32, [C.c_int32, _vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp]),
    "mhb_diff_stats_workspace": (C.c_int64, [C.c_int64]),
    "mhb_diff_stats_f64": (C.c_int32, [_vp, C.c_int64, C.c_double, _vp, C.c_int64, _vp, _vp]),
    "mhb_gradient": (C.c_int32, [C.c_int32, _vp, C.c_int64, _vp, _vp]),
    "mhb_zero_crossings": (C.c_int32, [C.c_int32, _vp, C.c_int64, C.c_double, _vp, _vp]),
    "mhb_slope_sum": (C.c_int32, [C.c_int32, _vp, C.c_int64, C.c_int32, _vp, _vp]),
    "mhb_haversine_elementwise": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_int64, _vp, _vp]),
    "mhb_haversine_vector": (C.c_int32, [C.c_double, C.c_double, _vp, _vp, C.c_int64, _vp, _vp]),
    "mhb_haversine_outer": (C.c_int32, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp]),
    "mhb_successive_distance": (C.c_int32, [_vp, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp]),
    "mhb_location_segments": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_int64, _vp, C.c_double, C.c_double,
                                          C.c_int64, _vp, _vp, _vp]),
    "mhb_label_stats": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "mhb_minmax_i64": (C.c_int32, [_vp, C.c_int64, _vp, _vp]),
}


def load():
    """Load libmhb200.so (once).  Raises MhbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MhbError("libmhb200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C pymhealth_b200/csrc` -- there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise MhbError("libmhb200.so lacks symbols declared in include/mhb200.h: %s" % ", ".join(missing))
    if lib.mhb_abi_version() != 1:
        raise MhbError("libmhb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status, what):
    if status == 0:
        return
    msg = load().mhb_last_error().decode("utf-8", "replace")
    if status < 0:
        if status == -3:
            raise NotImplementedError("%s: %s" % (what, msg))
        raise ValueError("%s: %s" % (what, msg))
    raise MhbError("%s: %s" % (what, msg))


def i32_array(vals):
    arr = (C.c_int32 * max(1, len(vals)))(*vals)
    return arr


def f64_array(vals):
    arr = (C.c_double * max(1, len(vals)))(*vals)
    return arr
