/*
 * mhb200.h -- C ABI of libmhb200.so, the B200 (sm_100a) implementation of the
 * sliding-window / spectral / location feature hot path of callumstew/pymhealth.
 *
 * The reference is a pure-Python package; its hot path has no FFI of its own except the
 * CFFI FFTW shim (src/mhealth/fft/_fftw_binder.py:8-29).  The entry points below are what a
 * binding for that path replaces; each cites the reference interface it stands in for
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with ``h_``;
 *   - the caller owns every buffer; the library allocates nothing, keeps no mutable global
 *     state and is re-entrant across streams and host threads;
 *   - all work is enqueued asynchronously on ``stream`` (a cudaStream_t passed as void*);
 *   - return value: 0 = OK, negative = argument error (mhb_status), positive = CUDA error code;
 *     mhb_last_error() returns a thread-local message for the last non-zero status;
 *   - float32 or float64 series in (``*_f32`` / ``*_f64``), features out as float64 (the
 *     reference's rolling_apply always returns float64, util/windows.py:89) or float32
 *     (``out_f32`` != 0) -- accumulation is float64 either way.
 */
#ifndef MHB200_H
#define MHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHB_ABI_VERSION 1

typedef enum mhb_status {
    MHB_OK = 0,
    MHB_E_ARG = -1,          /* null pointer, negative size, wsize < 1, wstep < 1 ... */
    MHB_E_FEATURE = -2,      /* unknown feature id for this entry point */
    MHB_E_UNSUPPORTED = -3,  /* valid request this build cannot serve (e.g. FFT length with a prime factor > 31) */
    MHB_E_WORKSPACE = -4     /* workspace too small (see the *_workspace_bytes query) */
} mhb_status;

/* Reducers of src/mhealth/generic/stats.py and generic/timedom.py, by id.  Column j of the
 * output table holds feature ``features[j]``. */
typedef enum mhb_feature {
    /* streaming family: mhb_window_stats_* */
    MHB_F_MEAN = 0,             /* stats.py:157  np.mean                                   */
    MHB_F_VAR = 1,              /* stats.py:160  np.var (population, ddof = 0)             */
    MHB_F_STD = 2,              /* stats.py:159  np.std                                    */
    MHB_F_MIN = 3,              /* stats.py:161  np.min                                    */
    MHB_F_MAX = 4,              /* stats.py:162  np.max                                    */
    MHB_F_DRANGE = 5,           /* stats.py:34-45                                          */
    MHB_F_SKEWNESS = 6,         /* stats.py:97-110                                         */
    MHB_F_KURTOSIS = 7,         /* stats.py:113-126                                        */
    MHB_F_KURTOSIS_EXCESS = 8,  /* stats.py:129-139                                        */
    MHB_F_COEFF_VAR = 9,        /* stats.py:142-153                                        */
    MHB_F_ZERO_CROSSINGS = 10,  /* timedom.py:52-64 (integer held in a float column)       */
    MHB_F_LINE_LENGTH = 11,     /* timedom.py:67-78                                        */
    MHB_F_HJORTH_ACTIVITY = 12, /* timedom.py:81-95                                        */
    MHB_F_SUM = 13,             /* np.sum (rollable in the reference)           */
    /* order / derivative family: mhb_window_order_* */
    MHB_F_MEDIAN = 32,          /* stats.py:158  np.median                                 */
    MHB_F_PERCENTILE = 33,      /* stats.py:163  np.percentile; q in feature_params[j]     */
    MHB_F_IQR = 34,             /* stats.py:48-59                                          */
    MHB_F_MODE = 35,            /* stats.py:62-94 (jit overload, incl. its first-run quirk) */
    MHB_F_HJORTH_MOBILITY = 36, /* timedom.py:98-114                                       */
    MHB_F_HJORTH_COMPLEXITY = 37 /* timedom.py:135-151                                     */
} mhb_feature;

/* Columns of mhb_window_spectral_*; ``spectral_features[j]`` selects, ``spectral_params``
 * carries (lo, hi) pairs where the column needs them. */
typedef enum mhb_spectral_feature {
    MHB_S_TOTAL_POWER = 0,        /* hrv.power_band(psd, freqs) with default bounds, heart/hrv.py:173-179 */
    MHB_S_BAND_POWER = 1,         /* heart/hrv.py:173-179, lo <= f <= hi (both inclusive)                  */
    MHB_S_REL_BAND_POWER = 2,     /* heart/hrv.py:192-198                                                  */
    MHB_S_PEAK_FREQUENCY = 3,     /* generic/frequency/density.py:18-32, lo <= f < hi, first max          */
    MHB_S_PEAK_BIN = 4,           /* the integer bin behind PEAK_FREQUENCY (bit-exact contract)           */
    MHB_S_ENTROPY = 5             /* generic/information.py:10-20 applied to the PSD row                  */
} mhb_spectral_feature;

typedef struct mhb_windows {
    int64_t n_series;        /* independent series (subject x axis), windows never span series  */
    int64_t series_len;      /* samples per series                                              */
    int64_t series_stride;   /* elements between series starts (>= series_len)                  */
    int32_t wsize;           /* window length W (util/windows.py:54-57)                         */
    int32_t wstep;           /* hop S                                                            */
} mhb_windows;

typedef struct mhb_table {
    void*   out;             /* device table                                                     */
    int32_t out_f32;         /* 0: float64 cells, 1: float32 cells                               */
    int64_t series_stride;   /* element strides of cell (series, window, column)                */
    int64_t window_stride;
    int64_t column_stride;
} mhb_table;

int32_t     mhb_abi_version(void);
const char* mhb_last_error(void);

/* windows per series: max(0, 1 + (series_len - wsize) / wstep)  (util/windows.py:86) */
int64_t mhb_n_windows(int64_t series_len, int32_t wsize, int32_t wstep);

/* ---- kernel 1a: streaming window statistics --------------------------------------------
 * Replaces rolling_apply(f)(arr, wsize, wstep) (util/windows.py:54-95) for every f in the
 * streaming family, all requested reducers in ONE pass over the data.
 * zc_threshold is timedom.zero_crossing_count's ``th`` (timedom.py:52). */
int32_t mhb_window_stats_f32(const float* x, const mhb_windows* geom,
                             const int32_t* h_features, int32_t n_features, double zc_threshold,
                             const mhb_table* table, void* stream);
int32_t mhb_window_stats_f64(const double* x, const mhb_windows* geom,
                             const int32_t* h_features, int32_t n_features, double zc_threshold,
                             const mhb_table* table, void* stream);
/* The same statistics of magnitude(x, y, z) = sqrt(x**2 + y**2 + z**2) (inertial/accelerometer.py:198-225), the step
 * the reference takes immediately before windowing tri-axial data, WITHOUT materialising the magnitude series: the
 * three axes (same geometry) are combined while a stage is copied into shared memory.  Bit-identical to
 * mhb_accel_elementwise(MAGNITUDE) followed by mhb_window_stats_*. */
int32_t mhb_window_stats_magnitude_f32(const float* x, const float* y, const float* z, const mhb_windows* geom,
                                       const int32_t* h_features, int32_t n_features, double zc_threshold,
                                       const mhb_table* table, void* stream);
int32_t mhb_window_stats_magnitude_f64(const double* x, const double* y, const double* z, const mhb_windows* geom,
                                       const int32_t* h_features, int32_t n_features, double zc_threshold,
                                       const mhb_table* table, void* stream);

/* ---- kernel 1b: per-window order statistics / derivative features -----------------------
 * Replaces rolling_apply(np.median | np.percentile | stats.interquartile_range | stats.mode |
 * timedom.hjorth_mobility | timedom.hjorth_complexity).  h_params[j] = q for PERCENTILE. */
int32_t mhb_window_order_f32(const float* x, const mhb_windows* geom,
                             const int32_t* h_features, const double* h_params, int32_t n_features,
                             const mhb_table* table, void* stream);
int32_t mhb_window_order_f64(const double* x, const mhb_windows* geom,
                             const int32_t* h_features, const double* h_params, int32_t n_features,
                             const mhb_table* table, void* stream);

/* ---- non-uniform (index-addressed) windows -----------------------------------------------------
 * get_indices(index, wsize, wstep) (util/windows.py:162-178): starts = arange(index[0], index[-1], wstep),
 * ends = starts + wsize, both located in ``index`` by a left searchsorted.  ``first`` = index[0];
 * ``n_windows`` = len(arange(index[0], index[-1], wstep)) (host arithmetic).  out_indices = int64 [2][n_windows]
 * (row 0 starts, row 1 ends).  _i64 serves integer / datetime64 / timedelta64 indices, _f64 float indices. */
int32_t mhb_get_indices_i64(const int64_t* index, int64_t n, int64_t first, int64_t wsize, int64_t wstep,
                            int64_t n_windows, int64_t* out_indices, void* stream);
int32_t mhb_get_indices_f64(const double* index, int64_t n, double first, double wsize, double wstep,
                            int64_t n_windows, int64_t* out_indices, void* stream);
/* indices_rolling_apply(f, min_window_len)(indices, arr) (util/windows.py:122-159): row i of the table =
 * reducers of x[starts[i] : ends[i]], NaN in every column when ends[i] - starts[i] < min_window_len (or the
 * window is empty).  Streaming family (ids MHB_F_MEAN..MHB_F_SUM), one warp per window, warp-shuffle
 * combination of float64 partial records.  table->series_stride is ignored (one series). */
int32_t mhb_segment_stats_f32(const float* x, int64_t n, const int64_t* starts, const int64_t* ends,
                              int64_t n_windows, int64_t min_window_len, const int32_t* h_features,
                              int32_t n_features, double zc_threshold, const mhb_table* table, void* stream);
int32_t mhb_segment_stats_f64(const double* x, int64_t n, const int64_t* starts, const int64_t* ends,
                              int64_t n_windows, int64_t min_window_len, const int32_t* h_features,
                              int32_t n_features, double zc_threshold, const mhb_table* table, void* stream);
/* the order / derivative family (ids MHB_F_MEDIAN..MHB_F_HJORTH_COMPLEXITY) on the same windows;
 * max_window_len = max(ends - starts) sizes the shared-memory sort buffer. */
int32_t mhb_segment_order_f32(const float* x, int64_t n, const int64_t* starts, const int64_t* ends,
                              int64_t n_windows, int64_t max_window_len, int64_t min_window_len,
                              const int32_t* h_features, const double* h_params, int32_t n_features,
                              const mhb_table* table, void* stream);
int32_t mhb_segment_order_f64(const double* x, int64_t n, const int64_t* starts, const int64_t* ends,
                              int64_t n_windows, int64_t max_window_len, int64_t min_window_len,
                              const int32_t* h_features, const double* h_params, int32_t n_features,
                              const mhb_table* table, void* stream);

/* ---- kernel 2: per-window FFT + PSD reducers ---------------------------------------------
 * Replaces the user-composed chain view -> mhealth.fft.fft -> |F|^2 -> hrv.power_band /
 * density.peak_frequency / information.entropy.  One-sided PSD, bins 0..W/2,
 * freqs = rfftfreq(W, 1/fs).  h_params holds 2 doubles (lo, hi) per column (ignored where the
 * column takes none; NaN = "None", i.e. unbounded). */
int32_t mhb_window_spectral_f32(const float* x, const mhb_windows* geom, double fs,
                                const int32_t* h_features, const double* h_params, int32_t n_features,
                                const mhb_table* table, void* stream);
/* Statistical and spectral columns of the SAME windows in one call -- what a caller of rolling_apply([np.mean, ...,
 * band power, ...]) (util/windows.py:98-107: one full pass per reducer in the reference) wants.  stat_features are
 * streaming-family ids (MHB_F_MEAN..MHB_F_SUM) written to stat_table, spec_features / spec_params as for
 * mhb_window_spectral_f32 written to spec_table (the two tables may be column ranges of one array).  Kernel 1a and
 * kernel 2 are launched back to back on `stream`. */
int32_t mhb_window_features_f32(const float* x, const mhb_windows* geom, const int32_t* stat_features,
                                int32_t n_stat, double zc_threshold, const mhb_table* stat_table, double fs,
                                const int32_t* spec_features, const double* spec_params, int32_t n_spec,
                                const mhb_table* spec_table, void* stream);
/* Batched complex DFT of real or complex rows, replaces fftw_fft(n, in, out, dir)
 * (fft/_fftw_binder.py:11-17) / numpy.fft.fft: in = [n_rows][n] (real: float64; complex:
 * interleaved float64 pairs), out = [n_rows][n] interleaved complex128.  direction -1 forward
 * (FFTW_FORWARD, fft/_fft.py:8), +1 backward scaled by 1/n (fft/_fft.py:46-48). */
int32_t mhb_fft_c128(const double* in, int32_t in_is_complex, int64_t n_rows, int32_t n,
                     int32_t direction, double* out, void* stream);
/* The same reducers on caller-supplied PSD rows (float64 [n_rows][nb]) and frequency vector -- the
 * literal signatures hrv.power_band(psd, freqs, lower, upper) (heart/hrv.py:173-179),
 * relative_power_band (:192-198), density.peak_frequency (density.py:18-32) and
 * information.entropy(x) (information.py:10-20; column MHB_S_ENTROPY, freqs may be NULL).
 * h_params: (lo, hi) per column, NaN = None.  out = float64 [n_rows][n_features]. */
int32_t mhb_psd_reduce_f64(const double* psd, const double* freqs, int64_t n_rows, int32_t nb,
                           const int32_t* h_features, const double* h_params, int32_t n_features,
                           double* out, void* stream);
/* raw one-sided PSD rows, float32 in -> float64|float32 [n_series][nw][W/2+1] */
int32_t mhb_window_psd_f32(const float* x, const mhb_windows* geom, void* psd_out, int32_t out_f32,
                           void* stream);

/* ---- accelerometer pre-stage (src/mhealth/inertial/accelerometer.py) --------------------------------
 * op 0 magnitude(x, y, z) :198-225 -> output of the input type (float32 arithmetic for float32, bit-identical);
 * op 1 roll(y, z) :13-25 and op 2 pitch(x, y, z) :44-56 -> float64 degrees (arctan2 in the input type).
 * x may be NULL for roll. */
int32_t mhb_accel_elementwise(int32_t op, int32_t is_f64, const void* x, const void* y, const void* z, int64_t n,
                              void* out, void* stream);
/* magnitude_dot(x, y, z) :236-259 = sqrt(x.x + y.y + z.z) -> out[0] (float64 accumulation, fixed order).
 * workspace: mhb_accel_sumsq_workspace(n) doubles. */
int64_t mhb_accel_sumsq_workspace(int64_t n);
int32_t mhb_accel_magnitude_dot(int32_t is_f64, const void* x, const void* y, const void* z, int64_t n,
                                double* workspace, int64_t workspace_len, double* out, void* stream);

/* ---- successive-difference statistics (HRV time-domain metrics, src/mhealth/heart/hrv.py:111-170) ---------
 * out7 = {n - 1, sum(diff), mean(diff), var(diff) (population), count(|diff| > abs_threshold), mean(diff^2),
 *         var(x[i+1] + x[i]) (population)}:
 * pnnx = out[4] / out[0], rmssd = sqrt(out[5]), ssd = out[1], sdsd = sqrt(out[3]); the Poincare widths
 * csi_sd1 = factor * sqrt(out[3]), csi_sd2 = factor * sqrt(out[6]) (hrv.py:207-231).
 * workspace: mhb_diff_stats_workspace(n - 1) doubles. */
int64_t mhb_diff_stats_workspace(int64_t n_diff);
int32_t mhb_diff_stats_f64(const double* x, int64_t n, double abs_threshold, double* workspace, int64_t workspace_len,
                           double* out7, void* stream);

/* timedom.gradient(x) (src/mhealth/generic/timedom.py:11-31) -> float64 [n] (differences in the input type);
 * timedom.zero_crossings(x, th) (:34-49) -> n - 1 flags, one byte each. */
int32_t mhb_gradient(int32_t is_f64, const void* x, int64_t n, double* out, void* stream);
int32_t mhb_zero_crossings(int32_t is_f64, const void* x, int64_t n, double threshold, uint8_t* out, void* stream);

/* Raw int16 sensor counts -> float32 samples (out[i] = in[i] * scale, one rounding): the widening rolling_apply does
 * per window on the host (util/windows.py:74-91 accepts any numeric dtype), done once on the device so that a host
 * caller ships 2 bytes per sample. */
int32_t mhb_widen_i16_f32(const int16_t* in, int64_t n, float scale, float* out, void* stream);

/* ppg.slope_sum(x, w) (src/mhealth/heart/ppg.py:28-42): out[i] = sum(diff(x)[i-w : i]) for w <= i < n - 1, else 0;
 * float64 [n] out whatever the input type. */
int32_t mhb_slope_sum(int32_t is_f64, const void* x, int64_t n, int32_t w, double* out, void* stream);

/* ---- kernel 3: location traces -------------------------------------------------------------
 * haversine gufuncs, location/distance.py:22-59 (float64 only, like the reference). */
int32_t mhb_haversine_elementwise(const double* lat1, const double* lon1, const double* lat2,
                                  const double* lon2, int64_t n, double* out, void* stream);
int32_t mhb_haversine_vector(double lat, double lon, const double* latcol, const double* loncol,
                             int64_t n, double* out, void* stream);
int32_t mhb_haversine_outer(const double* lat1, const double* lon1, int64_t n,
                            const double* lat2, const double* lon2, int64_t m, double* out, void* stream);
/* location/features.py:98-113: out[0] = 0 at every segment start, else step distance. */
int32_t mhb_successive_distance(const double* lat, const double* lon, const int64_t* seg_offsets,
                                int64_t n_segments, int64_t n, double* out, void* stream);

/* Per-segment (subject-day) feature rows.  Columns (float64), in this order:
 *   0 n_points  1 total_distance (sum of features.py:98-113)  2 location_variance
 *   (distribution.py:28-39)  3 radius_of_gyration [extension]  4 max_home_distance
 *   5 home_stay_count  6 proportion_home_stay (features.py:71-84)  7 n_stay_points [extension]
 *   8 n_labels (distribution.py:58-65)  9 label_entropy (distribution.py:79-89)
 *   10 normalized_label_entropy (distribution.py:92-102)
 * ``home`` = [n_segments][2] (lat, lon); ``t`` int64 seconds; labels_out (optional, may be
 * NULL) receives the stay-point label of every point (-1 = none). */
#define MHB_SEG_COLUMNS 11
int32_t mhb_location_segments(const double* lat, const double* lon, const int64_t* t,
                              const int64_t* seg_offsets, int64_t n_segments,
                              const double* home, double home_limit_km,
                              double stay_dist_km, int64_t stay_min_seconds,
                              double* out_rows, int64_t* labels_out, void* stream);

/* label statistics, location/distribution.py:58-102.  Dense histogram over
 * [label_min, label_max] (caller passes the range; workspace = (label_max-label_min+1) int64).
 * out3 = {n_labels, entropy, normalized_entropy}; counts_out (optional) gets the histogram. */
int32_t mhb_label_stats(const int64_t* labels, int64_t n, int64_t label_min, int64_t label_max,
                        int64_t n_clusters_override, int64_t* workspace_counts, double* out3,
                        void* stream);
int32_t mhb_minmax_i64(const int64_t* v, int64_t n, int64_t* out2, void* stream);

/* moments of one array as a single window with fp64 data (np.var(lat)+np.var(lon) etc. reuse
 * mhb_window_stats_f64 with wsize = n). */

#ifdef __cplusplus
}
#endif
#endif /* MHB200_H */
