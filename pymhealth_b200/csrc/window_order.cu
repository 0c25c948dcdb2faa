// Kernel 1b -- per-window order statistics and derivative features (sm_100a).
//
// Replaces rolling_apply(f) (reference src/mhealth/util/windows.py:68-91) for reducers that need the
// whole window at once:
//   np.median / np.percentile / stats.interquartile_range / stats.mode   (generic/stats.py:48-94,158,163)
//   timedom.hjorth_mobility / hjorth_complexity                           (generic/timedom.py:98-151)
// One thread group (a warp, or the whole CTA for windows that do not fit a warp's share of shared
// memory) owns a window: it stages the window in shared memory, evaluates the derivative features on
// the unsorted copy (two-pass float64 sums), then sorts it in place (bitonic network) and reads the
// order statistics off the sorted array.  All requested columns share the one staging + one sort.
//
// Parity: percentile interpolation is numba's (numba/np/arraymath.py:1696-1701):
//   rank = 1 + (n-1) q/100, f = floor(rank), m = rank - f, val = lower (1-m) + upper m,
// q = 0 / 100 short-circuit to min / max; median of an even window is (a+b)/2; mode reproduces the
// reference's run counter including its first-run quirk (stats.py:81-93).
#include <cstdlib>
#include <math_constants.h>

#include "common.cuh"
#include "sort_regs.cuh"

namespace mhb {

// window_order_blocks.cu: sort every block once + k-way selection (order statistics only); -100 = not covered
// window_order_stream.cu: float32, W = g or 2 g with S = g (warp-private streaming, no block barrier)
int32_t window_order_stream_try(const float* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                                const double* h_params, int32_t n_features, const mhb_table* table, void* stream_v);

template <typename InT>
int32_t window_order_blocks_try(const InT* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                                const double* h_params, int32_t n_features, const mhb_table* table, void* stream_v);

namespace {

constexpr int kMaxFeat = 32;

struct OrderPlan {
    const void* x;
    int64_t series_len, series_stride, nw, total_windows;
    int32_t W, S, P2;            // window, hop, sort length (power of two >= W)
    void* out;
    int64_t o_series, o_window, o_col;
    int32_t n_features;
    int32_t need_sort, need_hjorth;
    int32_t feat[kMaxFeat];
    double param[kMaxFeat];
    // segment mode (non-uniform windows, util/windows.py:122-159): window w = x[starts[w] : ends[w]]
    const int64_t* starts;
    const int64_t* ends;
    int64_t n_total, min_len;
};

template <typename T>
__device__ __forceinline__ T pos_inf();
template <>
__device__ __forceinline__ float pos_inf<float>() { return CUDART_INF_F; }
template <>
__device__ __forceinline__ double pos_inf<double>() { return CUDART_INF; }

// ---- thread-group abstraction: G = 32 (a warp) or G = blockDim (the CTA)
template <int G>
struct Group {
    __device__ static __forceinline__ int rank() { return G == 32 ? (threadIdx.x & 31) : threadIdx.x; }
    __device__ static __forceinline__ void sync() {
        if (G == 32) __syncwarp(); else __syncthreads();
    }
    // sum over the group, result valid in every thread
    __device__ static __forceinline__ double sum(double v, double* scratch) {
        v = warp_sum(v);
        if (G == 32) return v;
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        __syncthreads();
        if (l == 0) scratch[w] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < G / 32; ++i) t += scratch[i];
        return t;
    }
};

// central-difference gradient of the staged window, evaluated on the fly (timedom.py:11-31)
template <typename InT>
__device__ __forceinline__ double grad1(const InT* s, int i, int n) {
    if (i == 0) return static_cast<double>(s[1]) - static_cast<double>(s[0]);
    if (i == n - 1) return static_cast<double>(s[n - 1]) - static_cast<double>(s[n - 2]);
    return (static_cast<double>(s[i + 1]) - static_cast<double>(s[i - 1])) / 2;
}
template <typename InT>
__device__ __forceinline__ double grad2(const InT* s, int i, int n) {
    if (i == 0) return grad1(s, 1, n) - grad1(s, 0, n);
    if (i == n - 1) return grad1(s, n - 1, n) - grad1(s, n - 2, n);
    return (grad1(s, i + 1, n) - grad1(s, i - 1, n)) / 2;
}

template <typename InT>
__device__ __forceinline__ double percentile_sorted(const InT* s, int n, double q) {
    if (n == 1) return static_cast<double>(s[0]);
    if (q == 100.0) return static_cast<double>(s[n - 1]);
    if (q == 0.0) return static_cast<double>(s[0]);
    const double rank = 1 + (n - 1) * (q / 100.0);
    const double f = floor(rank);
    const double m = rank - f;
    const int fi = static_cast<int>(f);
    if (fi >= n) return static_cast<double>(s[n - 1]);      // q just below 100: the rank rounds to n (m = 0); never read s[n]
    const double lower = static_cast<double>(s[fi - 1]);
    const double upper = static_cast<double>(s[fi]);
    return lower * (1 - m) + upper * m;
}

template <typename InT>
__device__ double mode_sorted(const InT* s, int n) {
    // stats.py:81-93: e1 = s[0], c1 = 1, c2 = 0; equal neighbours bump c2, a new value resets it to 1
    InT best = s[0];
    int c1 = 1, c2 = 0;
    for (int i = 1; i < n; ++i) {
        if (s[i] == s[i - 1]) {
            ++c2;
            if (c2 > c1) {
                c1 = c2;
                best = s[i];
            }
        } else {
            c2 = 1;
        }
    }
    return static_cast<double>(best);
}

template <typename InT, typename OutT, int G>
__global__ void __launch_bounds__(256) window_order_kernel(const OrderPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double scratch[8];
    constexpr int groups_per_cta = 256 / G;
    const int grp = G == 32 ? (threadIdx.x >> 5) : 0;
    const int r = Group<G>::rank();
    InT* buf = reinterpret_cast<InT*>(smem_raw) + static_cast<size_t>(grp) * P.P2;
    const InT* xg = reinterpret_cast<const InT*>(P.x);
    int n = P.W;
    const bool segmented = P.starts != nullptr;

    for (int64_t w = static_cast<int64_t>(blockIdx.x) * groups_per_cta + grp; w < P.total_windows;
         w += static_cast<int64_t>(gridDim.x) * groups_per_cta) {
        int64_t series = w / P.nw;
        int64_t wi = w - series * P.nw;
        const InT* src = xg + series * P.series_stride + wi * P.S;
        if (segmented) {
            int64_t s = P.starts[w], e = P.ends[w];
            if (s < 0) s = 0;
            if (e > P.n_total) e = P.n_total;
            if (e < s) e = s;
            series = 0;
            wi = w;
            src = xg + s;
            n = static_cast<int>(e - s);
            if (n < P.min_len || n < 1 || (P.need_hjorth && n < 2)) {     // windows.py:152-155: too short -> NaN
                for (int j = r; j < P.n_features; j += G)
                    store_cell<OutT>(P.out, wi * P.o_window + j * P.o_col, CUDART_NAN);
                continue;                                                // group-uniform branch
            }
        }
        int p2w = P.P2;                        // sort length of this window
        if (segmented) {
            p2w = 1;
            while (p2w < n) p2w <<= 1;
        }
        Group<G>::sync();                      // previous window's readers are done with buf
        for (int i = r; i < p2w; i += G) buf[i] = i < n ? src[i] : pos_inf<InT>();
        Group<G>::sync();

        double mob = 0.0, cpx = 0.0;
        if (P.need_hjorth) {
            // two-pass population variances of x, x', x'' in float64 (np.var, arraymath.py:469-487)
            double sx = 0, s1 = 0, s2 = 0;
            for (int i = r; i < n; i += G) {
                sx += static_cast<double>(buf[i]);
                s1 += grad1(buf, i, n);
                s2 += grad2(buf, i, n);
            }
            const double inv = 1.0 / n;
            const double mx = Group<G>::sum(sx, scratch) * inv;
            const double m1 = Group<G>::sum(s1, scratch) * inv;
            const double m2 = Group<G>::sum(s2, scratch) * inv;
            double vx = 0, v1 = 0, v2 = 0;
            for (int i = r; i < n; i += G) {
                const double a = static_cast<double>(buf[i]) - mx;
                const double b = grad1(buf, i, n) - m1;
                const double cc = grad2(buf, i, n) - m2;
                vx += a * a;
                v1 += b * b;
                v2 += cc * cc;
            }
            vx = Group<G>::sum(vx, scratch) * inv;
            v1 = Group<G>::sum(v1, scratch) * inv;
            v2 = Group<G>::sum(v2, scratch) * inv;
            mob = sqrt(v1 / vx);                                // timedom.py:113-114
            cpx = sqrt(v2 / v1) / mob;                          // timedom.py:149-151
        }

        if (P.need_sort && G == 32 && p2w >= 32 && p2w <= 512) {
            // a warp sorts up to 512 elements in registers (shuffles across lanes), then puts them back
            switch (p2w) {
                case 32: sort_smem_via_regs<InT, 1>(buf, r); break;
                case 64: sort_smem_via_regs<InT, 2>(buf, r); break;
                case 128: sort_smem_via_regs<InT, 4>(buf, r); break;
                case 256: sort_smem_via_regs<InT, 8>(buf, r); break;
                default: sort_smem_via_regs<InT, 16>(buf, r); break;
            }
            Group<G>::sync();
        } else if (P.need_sort) {
            for (int k2 = 2; k2 <= p2w; k2 <<= 1) {
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int t = r; t < (p2w >> 1); t += G) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int l = i + j;
                        const InT a = buf[i], b = buf[l];
                        const bool up = (i & k2) == 0;
                        if ((a > b) == up) {
                            buf[i] = b;
                            buf[l] = a;
                        }
                    }
                    Group<G>::sync();
                }
            }
        }

        const int64_t obase = series * P.o_series + wi * P.o_window;
        for (int j = r; j < P.n_features; j += G) {       // one column per thread
            double v;
            switch (P.feat[j]) {
                case MHB_F_MEDIAN: {
                    const int h = n >> 1;
                    v = (n & 1) ? static_cast<double>(buf[h])
                                : (static_cast<double>(buf[h - 1]) + static_cast<double>(buf[h])) / 2;
                    break;
                }
                case MHB_F_PERCENTILE: v = percentile_sorted(buf, n, P.param[j]); break;
                case MHB_F_IQR: v = percentile_sorted(buf, n, 75.0) - percentile_sorted(buf, n, 25.0); break;
                case MHB_F_MODE: v = mode_sorted(buf, n); break;
                case MHB_F_HJORTH_MOBILITY: v = mob; break;
                case MHB_F_HJORTH_COMPLEXITY: v = cpx; break;
                default: v = 0.0; break;
            }
            store_cell<OutT>(P.out, obase + j * P.o_col, v);
        }
    }
}

template <typename InT>
int32_t window_order_impl(const InT* x, const mhb_windows* geom, const int32_t* h_features, const double* h_params,
                          int32_t n_features, const mhb_table* table, void* stream_v) {
    MHB_REQUIRE(geom && table, MHB_E_ARG, "window_order: null geometry/table");
    MHB_REQUIRE(geom->wsize >= 1 && geom->wstep >= 1, MHB_E_ARG, "window_order: wsize and wstep must be >= 1");
    MHB_REQUIRE(geom->n_series >= 0 && geom->series_len >= 0 && geom->series_stride >= geom->series_len,
                MHB_E_ARG, "window_order: bad series geometry");
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxFeat, MHB_E_ARG, "window_order: 0..%d features per call", kMaxFeat);
    const int64_t nw = n_windows_host(geom->series_len, geom->wsize, geom->wstep);
    if (nw == 0 || geom->n_series == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(x && table->out && h_features, MHB_E_ARG, "window_order: null pointer");

    OrderPlan P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features[j];
        MHB_REQUIRE(f >= MHB_F_MEDIAN && f <= MHB_F_HJORTH_COMPLEXITY, MHB_E_FEATURE,
                    "window_order: feature id %d is not in the order/derivative family", f);
        P.feat[j] = f;
        P.param[j] = h_params ? h_params[j] : 0.0;
        if (f == MHB_F_PERCENTILE)
            MHB_REQUIRE(P.param[j] >= 0.0 && P.param[j] <= 100.0, MHB_E_ARG,
                        "window_order: percentile q=%g outside [0, 100]", P.param[j]);
        if (f == MHB_F_HJORTH_MOBILITY || f == MHB_F_HJORTH_COMPLEXITY) {
            P.need_hjorth = 1;
            MHB_REQUIRE(geom->wsize >= 2, MHB_E_ARG, "window_order: Hjorth features need wsize >= 2");
        } else {
            P.need_sort = 1;
        }
    }
    if constexpr (sizeof(InT) == 4) {
        if (getenv("MHB_ORDER_FULLSORT") == nullptr && getenv("MHB_ORDER_NOSTREAM") == nullptr) {
            const int32_t st = window_order_stream_try(x, geom, nw, h_features, h_params, n_features, table, stream_v);
            if (st != -100) return st;
        }
    }
    if (getenv("MHB_ORDER_FULLSORT") == nullptr) {
        const int32_t st = window_order_blocks_try<InT>(x, geom, nw, h_features, h_params, n_features, table, stream_v);
        if (st != -100) return st;
    }
    P.n_features = n_features;
    P.x = x;
    P.series_len = geom->series_len;
    P.series_stride = geom->series_stride;
    P.nw = nw;
    P.total_windows = nw * geom->n_series;
    P.W = geom->wsize;
    P.S = geom->wstep;
    int64_t p2 = 1;
    while (p2 < geom->wsize) p2 <<= 1;
    MHB_REQUIRE(p2 * sizeof(InT) <= 192 * 1024, MHB_E_UNSUPPORTED,
                "window_order: wsize=%d does not fit one CTA's shared memory", geom->wsize);
    P.P2 = static_cast<int32_t>(p2);
    P.out = table->out;
    P.o_series = table->series_stride;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;

    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const bool warp_mode = p2 * sizeof(InT) * 8 <= 96 * 1024;       // 8 warps, <= 96 KB per CTA
    const size_t smem = static_cast<size_t>(p2) * sizeof(InT) * (warp_mode ? 8 : 1);
    const int64_t per_cta = warp_mode ? 8 : 1;
    int64_t ctas = (P.total_windows + per_cta - 1) / per_cta;
    const int64_t max_ctas = static_cast<int64_t>(kNumSMs) * 16;
    if (ctas > max_ctas) ctas = max_ctas;
    cudaError_t e;
    const bool f32 = table->out_f32 != 0;
#define MHB_GO(OUT, G)                                                                                       \
    {                                                                                                        \
        auto kern = window_order_kernel<InT, OUT, G>;                                                        \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
        if (e == cudaSuccess) {                                                                              \
            kern<<<static_cast<unsigned>(ctas), 256, smem, stream>>>(P);                                     \
            e = cudaGetLastError();                                                                          \
        }                                                                                                    \
    }
    if (warp_mode) {
        if (f32) MHB_GO(float, 32) else MHB_GO(double, 32)
    } else {
        if (f32) MHB_GO(float, 256) else MHB_GO(double, 256)
    }
#undef MHB_GO
    return cuda_status(e, "window_order launch");
}

template <typename InT>
int32_t segment_order_impl(const InT* x, int64_t n, const int64_t* starts, const int64_t* ends, int64_t n_windows,
                           int64_t max_window_len, int64_t min_window_len, const int32_t* h_features,
                           const double* h_params, int32_t n_features, const mhb_table* table, void* stream_v) {
    MHB_REQUIRE(table, MHB_E_ARG, "segment_order: null table");
    MHB_REQUIRE(n >= 0 && n_windows >= 0 && max_window_len >= 0, MHB_E_ARG, "segment_order: negative size");
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxFeat, MHB_E_ARG, "segment_order: 0..%d features per call", kMaxFeat);
    if (n_windows == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(starts && ends && table->out && h_features && (x || n == 0), MHB_E_ARG, "segment_order: null pointer");
    OrderPlan P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features[j];
        MHB_REQUIRE(f >= MHB_F_MEDIAN && f <= MHB_F_HJORTH_COMPLEXITY, MHB_E_FEATURE,
                    "segment_order: feature id %d is not in the order/derivative family", f);
        P.feat[j] = f;
        P.param[j] = h_params ? h_params[j] : 0.0;
        if (f == MHB_F_PERCENTILE)
            MHB_REQUIRE(P.param[j] >= 0.0 && P.param[j] <= 100.0, MHB_E_ARG,
                        "segment_order: percentile q=%g outside [0, 100]", P.param[j]);
        if (f == MHB_F_HJORTH_MOBILITY || f == MHB_F_HJORTH_COMPLEXITY) P.need_hjorth = 1;
        else P.need_sort = 1;
    }
    P.n_features = n_features;
    P.x = x;
    P.nw = n_windows;
    P.total_windows = n_windows;
    P.starts = starts;
    P.ends = ends;
    P.n_total = n;
    P.min_len = min_window_len;
    int64_t p2 = 1;
    while (p2 < max_window_len) p2 <<= 1;
    MHB_REQUIRE(p2 * sizeof(InT) <= 192 * 1024, MHB_E_UNSUPPORTED,
                "segment_order: windows of up to %lld samples do not fit one CTA's shared memory",
                static_cast<long long>(max_window_len));
    P.P2 = static_cast<int32_t>(p2);
    P.W = static_cast<int32_t>(max_window_len);
    P.S = 1;
    P.out = table->out;
    P.o_series = 0;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const bool warp_mode = p2 * sizeof(InT) * 8 <= 96 * 1024;
    const size_t smem = static_cast<size_t>(p2) * sizeof(InT) * (warp_mode ? 8 : 1);
    const int64_t per_cta = warp_mode ? 8 : 1;
    int64_t ctas = (n_windows + per_cta - 1) / per_cta;
    const int64_t max_ctas = static_cast<int64_t>(kNumSMs) * 16;
    if (ctas > max_ctas) ctas = max_ctas;
    cudaError_t e;
    const bool f32 = table->out_f32 != 0;
#define MHB_GO(OUT, G)                                                                                       \
    {                                                                                                        \
        auto kern = window_order_kernel<InT, OUT, G>;                                                        \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
        if (e == cudaSuccess) {                                                                              \
            kern<<<static_cast<unsigned>(ctas), 256, smem, stream>>>(P);                                     \
            e = cudaGetLastError();                                                                          \
        }                                                                                                    \
    }
    if (warp_mode) {
        if (f32) MHB_GO(float, 32) else MHB_GO(double, 32)
    } else {
        if (f32) MHB_GO(float, 256) else MHB_GO(double, 256)
    }
#undef MHB_GO
    return cuda_status(e, "segment_order launch");
}

}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_segment_order_f32(const float* x, int64_t n, const int64_t* starts, const int64_t* ends,
                                         int64_t n_windows, int64_t max_window_len, int64_t min_window_len,
                                         const int32_t* h_features, const double* h_params, int32_t n_features,
                                         const mhb_table* table, void* stream) {
    return mhb::segment_order_impl<float>(x, n, starts, ends, n_windows, max_window_len, min_window_len, h_features,
                                          h_params, n_features, table, stream);
}

extern "C" int32_t mhb_segment_order_f64(const double* x, int64_t n, const int64_t* starts, const int64_t* ends,
                                         int64_t n_windows, int64_t max_window_len, int64_t min_window_len,
                                         const int32_t* h_features, const double* h_params, int32_t n_features,
                                         const mhb_table* table, void* stream) {
    return mhb::segment_order_impl<double>(x, n, starts, ends, n_windows, max_window_len, min_window_len, h_features,
                                           h_params, n_features, table, stream);
}

extern "C" int32_t mhb_window_order_f32(const float* x, const mhb_windows* geom, const int32_t* h_features,
                                        const double* h_params, int32_t n_features, const mhb_table* table,
                                        void* stream) {
    return mhb::window_order_impl<float>(x, geom, h_features, h_params, n_features, table, stream);
}

extern "C" int32_t mhb_window_order_f64(const double* x, const mhb_windows* geom, const int32_t* h_features,
                                        const double* h_params, int32_t n_features, const mhb_table* table,
                                        void* stream) {
    return mhb::window_order_impl<double>(x, geom, h_features, h_params, n_features, table, stream);
}
