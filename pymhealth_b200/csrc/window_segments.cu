// Non-uniform (index-addressed) windows -- the timestamp path of the reference's window driver.
//
// Replaces (reference src/mhealth/util/windows.py):
//   get_indices(index, wsize, wstep)              :162-178  starts = arange(index[0], index[-1], wstep),
//                                                           ends = starts + wsize, left searchsorted of both
//   indices_rolling_apply(f)(indices, arr)        :122-159  out[i] = f(arr[si:ei]) or NaN when ei - si < min_window_len
//   nonuniform_rolling_apply(f)(index, arr, ...)  :181-249  the two composed
// for every reducer of the streaming family (generic/stats.py, generic/timedom.py; ids MHB_F_MEAN..MHB_F_SUM).
// The order / derivative family on the same windows lives in window_order.cu (mhb_segment_order_*).
//
// Windows are arbitrary [start, end) index pairs (overlapping, unordered, empty), so there is no block
// sharing to exploit: one warp owns one window, lanes stride over its samples with float64 shifted power
// sums (pivot = the window's first sample), and the 32 partial records are combined with warp shuffles in a
// fixed order (deterministic).
#include <math_constants.h>

#include "common.cuh"

namespace mhb {

namespace {

constexpr int kMaxFeatSeg = 32;
constexpr int kSegThreads = 256;

struct SegPlan {
    const void* x;
    int64_t n;
    const int64_t* starts;
    const int64_t* ends;
    int64_t n_windows;
    int64_t min_len;
    double th;
    void* out;
    int64_t o_window, o_col;
    int32_t n_features;
    int32_t feat[kMaxFeatSeg];
    // uniform mode (starts == nullptr): window w of nw per series starts at series * series_stride + wi * S, length W
    int64_t nw, series_stride, o_series;
    int32_t W, S;
};

struct SegAcc {
    double s1, s2, s3, s4, ll;
    double mn, mx;
    int zc;
};

__device__ __forceinline__ void seg_merge(SegAcc& a, const SegAcc& b) {
    a.s1 += b.s1;
    a.s2 += b.s2;
    a.s3 += b.s3;
    a.s4 += b.s4;
    a.ll += b.ll;
    a.mn = fmin(a.mn, b.mn);
    a.mx = fmax(a.mx, b.mx);
    a.zc += b.zc;
}

__device__ __forceinline__ SegAcc seg_shfl_down(const SegAcc& a, int o) {
    SegAcc b;
    b.s1 = __shfl_down_sync(0xffffffffu, a.s1, o);
    b.s2 = __shfl_down_sync(0xffffffffu, a.s2, o);
    b.s3 = __shfl_down_sync(0xffffffffu, a.s3, o);
    b.s4 = __shfl_down_sync(0xffffffffu, a.s4, o);
    b.ll = __shfl_down_sync(0xffffffffu, a.ll, o);
    b.mn = __shfl_down_sync(0xffffffffu, a.mn, o);
    b.mx = __shfl_down_sync(0xffffffffu, a.mx, o);
    b.zc = __shfl_down_sync(0xffffffffu, a.zc, o);
    return b;
}

// one thread's share of a window: samples s + r, s + r + G, ...  (pair terms look one sample back)
template <typename InT>
__device__ __forceinline__ SegAcc seg_scan(const InT* __restrict__ x, int64_t s, int64_t e, int r, int G, double c,
                                           double th) {
    SegAcc a;
    a.s1 = a.s2 = a.s3 = a.s4 = a.ll = 0.0;
    a.mn = CUDART_INF;
    a.mx = -CUDART_INF;
    a.zc = 0;
    for (int64_t i = s + r; i < e; i += G) {
        const double v = static_cast<double>(x[i]);
        const double d = v - c;
        const double d2 = d * d;
        a.s1 += d;
        a.s2 += d2;
        a.s3 = fma(d2, d, a.s3);
        a.s4 = fma(d2, d2, a.s4);
        a.mn = fmin(a.mn, v);
        a.mx = fmax(a.mx, v);
        if (i > s) {
            const double p = static_cast<double>(x[i - 1]);
            a.ll += fabs(v - p);
            // timedom.py:46-49: |x| <= th is zeroed, pos = x > 0  <=>  pos = x > th (th >= 0)
            a.zc += ((v > th) != (p > th)) ? 1 : 0;
        }
    }
    return a;
}

template <typename OutT>
__device__ __forceinline__ void seg_emit(const SegPlan& P, int64_t w, int64_t len, const SegAcc& a, double c,
                                         bool valid) {
    const double n = static_cast<double>(len);
    double mean = CUDART_NAN, var = CUDART_NAN, sd = CUDART_NAN, skew = CUDART_NAN, kurt = CUDART_NAN;
    if (valid) {
        const double inv_n = 1.0 / n;
        const double dl = a.s1 * inv_n;
        mean = c + dl;
        double M2 = a.s2 - a.s1 * dl;
        if (M2 < 0.0 || a.mn == a.mx) M2 = 0.0;
        var = M2 * inv_n;
        sd = sqrt(var);
        skew = 0.0;
        kurt = 0.0;
        if (var > 0.0) {
            const double dl2 = dl * dl;
            const double M3 = a.s3 - 3.0 * dl * a.s2 + 2.0 * n * dl2 * dl;
            const double M4 = a.s4 - 4.0 * dl * a.s3 + 6.0 * dl2 * a.s2 - 3.0 * n * dl2 * dl2;
            const double inv_var = 1.0 / var;
            skew = (M3 * inv_n) * inv_var / sd;
            kurt = (M4 * inv_n) * inv_var * inv_var;
        }
    }
    for (int j = 0; j < P.n_features; ++j) {
        double v = CUDART_NAN;
        if (valid) {
            switch (P.feat[j]) {
                case MHB_F_MEAN: v = mean; break;
                case MHB_F_VAR:
                case MHB_F_HJORTH_ACTIVITY: v = var; break;
                case MHB_F_STD: v = sd; break;
                case MHB_F_MIN: v = a.mn; break;
                case MHB_F_MAX: v = a.mx; break;
                case MHB_F_DRANGE: v = a.mx - a.mn; break;
                case MHB_F_SKEWNESS: v = skew; break;
                case MHB_F_KURTOSIS: v = kurt; break;
                case MHB_F_KURTOSIS_EXCESS: v = kurt - 3.0; break;
                case MHB_F_COEFF_VAR: v = sd / mean; break;
                case MHB_F_ZERO_CROSSINGS: v = static_cast<double>(a.zc); break;
                case MHB_F_LINE_LENGTH: v = a.ll; break;
                case MHB_F_SUM: v = mean * n; break;
                default: break;
            }
        }
        int64_t row = w * P.o_window;
        if (P.starts == nullptr) {
            const int64_t series = w / P.nw;
            row = series * P.o_series + (w - series * P.nw) * P.o_window;
        }
        store_cell<OutT>(P.out, row + j * P.o_col, v);
    }
}

__device__ __forceinline__ void seg_bounds(const SegPlan& P, int64_t w, int64_t& s, int64_t& e) {
    if (P.starts == nullptr) {
        const int64_t series = w / P.nw, wi = w - series * P.nw;
        s = series * P.series_stride + wi * P.S;
        e = s + P.W;
        return;
    }
    s = P.starts[w];
    e = P.ends[w];
    // arr[si:ei] slicing clamps to the array (negative, i.e. from-the-end, indices are not produced by get_indices)
    if (s < 0) s = 0;
    if (e > P.n) e = P.n;
    if (e < s) e = s;
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kSegThreads) segment_stats_kernel(const SegPlan P) {
    const InT* x = reinterpret_cast<const InT*>(P.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPC = kSegThreads / 32;
    // one window per warp, whatever its length: the windows of a call are independent units of very different
    // sizes (a 5-minute RR segment, a day of 1 Hz GPS), and a grid-stride loop over warps balances them
    for (int64_t w = static_cast<int64_t>(blockIdx.x) * WPC + warp; w < P.n_windows;
         w += static_cast<int64_t>(gridDim.x) * WPC) {
        int64_t s, e;
        seg_bounds(P, w, s, e);
        const int64_t len = e - s;
        const bool valid = len >= P.min_len && len > 0;
        SegAcc a;
        double c = 0.0;
        if (valid) {
            c = static_cast<double>(x[s]);
            a = seg_scan<InT>(x, s, e, lane, 32, c, P.th);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const SegAcc b = seg_shfl_down(a, o);
                seg_merge(a, b);
            }
        }
        if (lane == 0) seg_emit<OutT>(P, w, len, a, c, valid);
    }
}

template <typename InT>
int32_t segment_stats_impl(const InT* x, int64_t n, const int64_t* starts, const int64_t* ends, int64_t n_windows,
                           int64_t min_window_len, const int32_t* h_features, int32_t n_features, double zc_threshold,
                           const mhb_table* table, void* stream_v) {
    MHB_REQUIRE(table, MHB_E_ARG, "segment_stats: null table");
    MHB_REQUIRE(n >= 0 && n_windows >= 0, MHB_E_ARG, "segment_stats: negative size");
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxFeatSeg, MHB_E_ARG, "segment_stats: 0..%d features per call",
                kMaxFeatSeg);
    SegPlan P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features ? h_features[j] : -1;
        MHB_REQUIRE(f >= MHB_F_MEAN && f <= MHB_F_SUM, MHB_E_FEATURE,
                    "segment_stats: feature id %d is not in the streaming family", f);
        P.feat[j] = f;
    }
    if (n_windows == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(starts && ends && table->out && (x || n == 0), MHB_E_ARG, "segment_stats: null pointer");
    P.x = x;
    P.n = n;
    P.starts = starts;
    P.ends = ends;
    P.n_windows = n_windows;
    P.min_len = min_window_len;
    P.th = zc_threshold > 0.0 ? zc_threshold : 0.0;
    P.out = table->out;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;
    P.n_features = n_features;
    int64_t ctas = (n_windows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    if (ctas > cap) ctas = cap;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (table->out_f32)
        segment_stats_kernel<InT, float><<<static_cast<unsigned>(ctas), kSegThreads, 0, stream>>>(P);
    else
        segment_stats_kernel<InT, double><<<static_cast<unsigned>(ctas), kSegThreads, 0, stream>>>(P);
    return cuda_status(cudaGetLastError(), "segment_stats launch");
}

}  // namespace

// Uniform windows through the warp-per-window kernel: the fallback of mhb_window_stats_* for geometries whose block
// decomposition degenerates (gcd(W, S) so small that a window is more than ~1000 blocks, e.g. co-prime W and S).
// Every window is evaluated directly from global memory (overlapping windows re-read L2), which is what the
// reference's loop does too; all other geometries take the streaming kernel.
template <typename InT>
int32_t window_stats_direct(const InT* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                            int32_t n_features, double zc_threshold, const mhb_table* table, void* stream_v) {
    SegPlan P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_features; ++j) P.feat[j] = h_features[j];
    P.x = x;
    P.n = (geom->n_series - 1) * geom->series_stride + geom->series_len;
    P.n_windows = nw * geom->n_series;
    P.min_len = 1;
    P.th = zc_threshold > 0.0 ? zc_threshold : 0.0;
    P.out = table->out;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;
    P.o_series = table->series_stride;
    P.n_features = n_features;
    P.nw = nw;
    P.series_stride = geom->series_stride;
    P.W = geom->wsize;
    P.S = geom->wstep;
    int64_t ctas = (P.n_windows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    if (ctas > cap) ctas = cap;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (table->out_f32)
        segment_stats_kernel<InT, float><<<static_cast<unsigned>(ctas), kSegThreads, 0, stream>>>(P);
    else
        segment_stats_kernel<InT, double><<<static_cast<unsigned>(ctas), kSegThreads, 0, stream>>>(P);
    return cuda_status(cudaGetLastError(), "window_stats (direct) launch");
}
template int32_t window_stats_direct<float>(const float*, const mhb_windows*, int64_t, const int32_t*, int32_t, double,
                                            const mhb_table*, void*);
template int32_t window_stats_direct<double>(const double*, const mhb_windows*, int64_t, const int32_t*, int32_t, double,
                                             const mhb_table*, void*);

namespace {

// ---- get_indices: left searchsorted of first + i*step (starts) and first + i*step + size (ends)
template <typename T>
__device__ __forceinline__ T arange_key(T first, T step, int64_t i);
template <>
__device__ __forceinline__ int64_t arange_key<int64_t>(int64_t first, int64_t step, int64_t i) { return first + i * step; }
template <>
__device__ __forceinline__ double arange_key<double>(double first, double step, int64_t i) {
    const double second = __dadd_rn(first, step);
    if (i == 0) return first;
    if (i == 1) return second;
    return __dadd_rn(first, __dmul_rn(static_cast<double>(i), __dsub_rn(second, first)));
}

template <typename T>
__global__ void get_indices_kernel(const T* __restrict__ index, int64_t n, T first, T wsize, T wstep, int64_t nwin,
                                   int64_t* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= 2 * nwin) return;
    const int64_t wi = i < nwin ? i : i - nwin;
    // np.arange(start, stop, step)[i]: exact for integers.  For floats numpy stores start and start + step, then FILLS
    // the rest as start + i * delta with delta = (start + step) - start -- not always the same double as step
    // (0.1, 0.2 -> 0.20000000000000004) -- each operation rounded on its own (no FMA contraction)
    T key = arange_key<T>(first, wstep, wi);
    if (i >= nwin) key = key + wsize;
    int64_t lo = 0, hi = n;                  // first position with index[pos] >= key  (side='left')
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (index[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    out[i] = lo;
}

template <typename T>
int32_t get_indices_impl(const T* index, int64_t n, T first, T wsize, T wstep, int64_t n_windows, int64_t* out,
                         void* stream_v) {
    MHB_REQUIRE(n >= 0 && n_windows >= 0, MHB_E_ARG, "get_indices: negative size");
    if (n_windows == 0) return MHB_OK;
    MHB_REQUIRE(index && out, MHB_E_ARG, "get_indices: null pointer");
    const int64_t total = 2 * n_windows;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    get_indices_kernel<T><<<blocks, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(index, n, first, wsize, wstep,
                                                                                   n_windows, out);
    return cuda_status(cudaGetLastError(), "get_indices launch");
}

}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_get_indices_i64(const int64_t* index, int64_t n, int64_t first, int64_t wsize, int64_t wstep,
                                       int64_t n_windows, int64_t* out_indices, void* stream) {
    return mhb::get_indices_impl<int64_t>(index, n, first, wsize, wstep, n_windows, out_indices, stream);
}

extern "C" int32_t mhb_get_indices_f64(const double* index, int64_t n, double first, double wsize, double wstep,
                                       int64_t n_windows, int64_t* out_indices, void* stream) {
    return mhb::get_indices_impl<double>(index, n, first, wsize, wstep, n_windows, out_indices, stream);
}

extern "C" int32_t mhb_segment_stats_f32(const float* x, int64_t n, const int64_t* starts, const int64_t* ends,
                                         int64_t n_windows, int64_t min_window_len, const int32_t* h_features,
                                         int32_t n_features, double zc_threshold, const mhb_table* table,
                                         void* stream) {
    return mhb::segment_stats_impl<float>(x, n, starts, ends, n_windows, min_window_len, h_features, n_features,
                                          zc_threshold, table, stream);
}

extern "C" int32_t mhb_segment_stats_f64(const double* x, int64_t n, const int64_t* starts, const int64_t* ends,
                                         int64_t n_windows, int64_t min_window_len, const int32_t* h_features,
                                         int32_t n_features, double zc_threshold, const mhb_table* table,
                                         void* stream) {
    return mhb::segment_stats_impl<double>(x, n, starts, ends, n_windows, min_window_len, h_features, n_features,
                                           zc_threshold, table, stream);
}
