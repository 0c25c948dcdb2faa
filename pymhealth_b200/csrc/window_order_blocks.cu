// Kernel 1b (fast path) -- order statistics of overlapping windows WITHOUT sorting every window.
//
// Replaces rolling_apply(np.median | np.percentile | stats.interquartile_range) (reference
// src/mhealth/generic/stats.py:48-59,158,163 on the driver of util/windows.py:68-91; numba's percentile interpolation,
// numba/np/arraymath.py:1696-1701).
//
// g = gcd(W, S): a window is k = W/g consecutive blocks and successive windows start hop = S/g blocks apart, so a block
// belongs to k/hop windows.  Every block is sorted ONCE (one warp per block, bitonic network in shared memory); an order
// statistic of a window is then a k-way SELECTION over its sorted blocks: the blocks are cut into at most 32 sorted
// segments, one per lane, and the warp bisects on the VALUE (order-preserving integer keys of the floats): a probe is
// one binary search per lane plus a warp sum of the counts, at most 32 (64) probes for float32 (float64), fewer when the
// window's minimum and maximum share leading bits.  The smallest key whose count reaches r + 1 IS the element of rank r,
// so selected values are exact; the neighbour of rank r + 1 (the interpolation partner) falls out of one more probe.
// Config 3 (W = 500, S = 250) sorts 256 instead of 512 elements per window; config 4 (W = 1920, S = 64) sorts 64 instead
// of 2048.  Anything this file does not cover (mode, Hjorth features, tiny or co-prime geometries) stays on
// window_order.cu.
#include <math_constants.h>

#include "common.cuh"
#include "sort_regs.cuh"

namespace mhb {

namespace {

constexpr int kMaxFeatB = 32;
constexpr int kThreadsOB = 256;
constexpr int kWarpsOB = kThreadsOB / 32;

struct BlocksPlan {
    const void* x;
    int64_t series_stride, nw, batches_per_series, total_batches;
    int32_t W, g, k, hop, P2g, nwb, nseg, seglen;
    void* out;
    int64_t o_series, o_window, o_col;
    int32_t n_features;
    int32_t feat[kMaxFeatB];
    double param[kMaxFeatB];
};

template <typename T>
struct Key;
template <>
struct Key<float> {
    using type = uint32_t;
    __device__ static __forceinline__ uint32_t enc(float v) {
        const uint32_t b = __float_as_uint(v);
        return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    }
    __device__ static __forceinline__ float dec(uint32_t k) {
        return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
    }
    __device__ static __forceinline__ float inf() { return CUDART_INF_F; }
};
template <>
struct Key<double> {
    using type = unsigned long long;
    __device__ static __forceinline__ unsigned long long enc(double v) {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
        return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    }
    __device__ static __forceinline__ double dec(unsigned long long k) {
        return __longlong_as_double(static_cast<long long>(k ^ ((k >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull)));
    }
    __device__ static __forceinline__ double inf() { return CUDART_INF; }
};

// number of elements <= v in the sorted segment s[0 .. len)
template <typename T>
__device__ __forceinline__ int count_le(const T* __restrict__ s, int len, T v) {
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <typename T>
__device__ __forceinline__ T warp_min_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u < v ? u : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// elements of rank r and r + 1 (0-based; the second clamps to the first at the top) of the union of the warp's segments:
// bisection on the value, one binary search per lane and a warp sum per probe
template <typename T>
__device__ __forceinline__ void select_pair(const T* __restrict__ seg, int len, int r, int n, T wmin, T wmax, T& lower,
                                            T& upper) {
    using K = typename Key<T>::type;
    K lo = Key<T>::enc(wmin), hi = Key<T>::enc(wmax);
    while (lo < hi) {
        const K mid = lo + ((hi - lo) >> 1);
        const int cnt = warp_sum(count_le<T>(seg, len, Key<T>::dec(mid)));
        if (cnt >= r + 1) hi = mid;
        else lo = mid + 1;
    }
    lower = Key<T>::dec(lo);
    upper = lower;
    if (r + 1 < n) {
        const int mine = count_le<T>(seg, len, lower);
        const int cnt = warp_sum(mine);
        if (cnt < r + 2) upper = warp_min_t<T>(mine < len ? seg[mine] : Key<T>::inf());
    }
}

// the same for a window made of ONE or TWO sorted blocks A, B of g elements (k <= 2: non-overlapping or 50 % overlapping
// windows): the merge-path partition, a binary search on how many of the r + 1 smallest elements come from A.  One
// THREAD per (window, column): no warp cooperation needed.
template <typename T>
__device__ __forceinline__ void merge_pair(const T* __restrict__ A, const T* __restrict__ B, int g, int gb, int r, T& lower,
                                           T& upper) {
    int lo = r + 1 - gb > 0 ? r + 1 - gb : 0;
    int hi = r + 1 < g ? r + 1 : g;
    while (lo < hi) {
        const int i = (lo + hi) >> 1;           // i from A, j = r + 1 - i from B (j >= 1 here because i < hi <= r + 1)
        const int j = r + 1 - i;
        if (B[j - 1] > A[i]) lo = i + 1;        // A[i] belongs to the r + 1 smallest: take more from A
        else hi = i;
    }
    const int i = lo, j = r + 1 - lo;
    const T inf = Key<T>::inf();
    const T a_last = i > 0 ? A[i - 1] : -inf, b_last = j > 0 ? B[j - 1] : -inf;
    lower = a_last > b_last ? a_last : b_last;
    const T a_next = i < g ? A[i] : inf, b_next = j < gb ? B[j] : inf;
    upper = a_next < b_next ? a_next : b_next;
    if (upper == inf) upper = lower;            // r is the top rank
}

// MERGE: k <= 2 (merge-path selection, a thread per window and column); otherwise k-way bisection (a warp per window).
// EPL > 0: blocks of <= 32 * EPL samples are sorted in registers (sort_regs.cuh), one warp per block; the samples of
// the warp's NEXT block (of this batch, or its first one of the CTA's next batch -- registers do not depend on the
// shared-memory barriers) are loaded before the current one is sorted, so the ~1 us of an HBM load is never exposed.
// EPL == 0: larger blocks, bitonic network in shared memory.
template <typename InT, typename OutT, bool MERGE, int EPL>
__global__ void __launch_bounds__(kThreadsOB, 4) window_order_blocks_kernel(const BlocksPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    InT* sorted = reinterpret_cast<InT*>(smem_raw);                      // [NB][P2g + 1]
    const InT* xg = reinterpret_cast<const InT*>(P.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P2 = P.P2g, g = P.g, n = P.W;
    const int BS = P2 + 1;       // block stride: odd, so that lanes probing the same position of different blocks do not collide
    constexpr int NR = EPL > 0 ? EPL : 1;

    auto batch_geom = [&](int64_t b, int64_t& series, int64_t& w0, int& nwin, int& nblk, const InT*& src0) {
        series = b / P.batches_per_series;
        const int64_t bi = b - series * P.batches_per_series;
        w0 = bi * P.nwb;
        const int64_t left = P.nw - w0;
        nwin = left < P.nwb ? static_cast<int>(left) : P.nwb;
        nblk = (nwin - 1) * P.hop + P.k;
        src0 = xg + series * P.series_stride + w0 * P.hop * static_cast<int64_t>(g);
    };
    auto load_block = [&](InT* v, const InT* src) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int e = lane * NR + i;
            v[i] = e < g ? src[e] : Key<InT>::inf();
        }
    };
    InT nxt[NR];
    if (EPL > 0 && blockIdx.x < P.total_batches) {
        int64_t series, w0;
        int nwin, nblk;
        const InT* src0;
        batch_geom(blockIdx.x, series, w0, nwin, nblk, src0);
        if (warp < nblk) load_block(nxt, src0 + static_cast<int64_t>(warp) * g);
    }

    for (int64_t b = blockIdx.x; b < P.total_batches; b += gridDim.x) {
        int64_t series, w0;
        int nwin, nblk;
        const InT* src0;
        batch_geom(b, series, w0, nwin, nblk, src0);
        __syncthreads();                                                  // previous batch's selections are done
        // ---- phase 1: stage and sort every block of the batch, one warp per block
        for (int blk = warp; blk < nblk; blk += kWarpsOB) {
            InT* buf = sorted + static_cast<size_t>(blk) * BS;
            const InT* src = src0 + static_cast<int64_t>(blk) * g;
            if constexpr (EPL > 0) {               // register-resident network (<= 16 elements per lane)
                InT v[NR];
#pragma unroll
                for (int i = 0; i < NR; ++i) v[i] = nxt[i];
                if (blk + kWarpsOB < nblk) {
                    load_block(nxt, src + static_cast<int64_t>(kWarpsOB) * g);
                } else if (b + gridDim.x < P.total_batches) {
                    int64_t series2, w02;
                    int nwin2, nblk2;
                    const InT* src2;
                    batch_geom(b + gridDim.x, series2, w02, nwin2, nblk2, src2);
                    if (warp < nblk2) load_block(nxt, src2 + static_cast<int64_t>(warp) * g);
                }
                warp_sort_regs_dispatch<NR>(v, lane);
#pragma unroll
                for (int i = 0; i < NR; ++i) buf[lane * NR + i] = v[i];
            } else {
                for (int i = lane; i < P2; i += 32) buf[i] = i < g ? src[i] : Key<InT>::inf();
                __syncwarp();
                for (int k2 = 2; k2 <= P2; k2 <<= 1) {
                    for (int j = k2 >> 1; j > 0; j >>= 1) {
                        for (int t = lane; t < (P2 >> 1); t += 32) {
                            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int l = i + j;
                            const InT a = buf[i], c = buf[l];
                            const bool up = (i & k2) == 0;
                            if ((a > c) == up) {
                                buf[i] = c;
                                buf[l] = a;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (EPL > 0 && warp >= nblk && b + gridDim.x < P.total_batches) {     // a warp without a block in a short batch
            int64_t series2, w02;
            int nwin2, nblk2;
            const InT* src2;
            batch_geom(b + gridDim.x, series2, w02, nwin2, nblk2, src2);
            if (warp < nblk2) load_block(nxt, src2 + static_cast<int64_t>(warp) * g);
        }
        __syncthreads();
        // ---- phase 2: order statistics of every window of the batch
        if (MERGE) {
            // thread per (window, column)
            for (int it = threadIdx.x; it < nwin * P.n_features; it += kThreadsOB) {
                const int wl = it / P.n_features, j = it - wl * P.n_features;
                const InT* A = sorted + static_cast<size_t>(wl * P.hop) * BS;
                const InT* B = P.k == 2 ? A + BS : A;
                const int gb = P.k == 2 ? g : 0;
                const InT wmin = (gb && B[0] < A[0]) ? B[0] : A[0];
                const InT wmax = (gb && B[g - 1] > A[g - 1]) ? B[g - 1] : A[g - 1];
                const int f = P.feat[j];
                double v;
                if (f == MHB_F_MEDIAN) {
                    InT a, c;
                    merge_pair<InT>(A, B, g, gb, (n & 1) ? (n >> 1) : (n >> 1) - 1, a, c);
                    v = (n & 1) ? static_cast<double>(a) : (static_cast<double>(a) + static_cast<double>(c)) / 2;
                } else {
                    // numba's np.percentile (arraymath.py:1696-1701); IQR = p75 - p25 (stats.py:48-59)
                    const int reps = f == MHB_F_IQR ? 2 : 1;
                    double acc = 0.0;
                    for (int rep = 0; rep < reps; ++rep) {
                        const double q = f == MHB_F_IQR ? (rep == 0 ? 75.0 : 25.0) : P.param[j];
                        double pv;
                        if (n == 1 || q == 0.0) {
                            pv = static_cast<double>(wmin);
                        } else if (q == 100.0) {
                            pv = static_cast<double>(wmax);
                        } else {
                            const double rank = 1 + (n - 1) * (q / 100.0);
                            const double fl = floor(rank);
                            const double m = rank - fl;
                            InT a, c;
                            merge_pair<InT>(A, B, g, gb, static_cast<int>(fl) - 1, a, c);
                            pv = static_cast<double>(a) * (1 - m) + static_cast<double>(c) * m;
                        }
                        acc = rep == 0 ? pv : acc - pv;
                    }
                    v = acc;
                }
                store_cell<OutT>(P.out, series * P.o_series + (w0 + wl) * P.o_window + j * P.o_col, v);
            }
            continue;
        }
        // one warp per window; lane = one sorted segment of one of the window's k blocks
        const int lb = lane / P.nseg, sg = lane - lb * P.nseg;
        for (int wl = warp; wl < nwin; wl += kWarpsOB) {
            const InT* seg = sorted;
            int len = 0;
            InT bmin = Key<InT>::inf(), bmax = -Key<InT>::inf();
            if (lb < P.k) {
                const InT* blk = sorted + static_cast<size_t>(wl * P.hop + lb) * BS;
                const int start = sg * P.seglen;
                len = g - start;
                if (len > P.seglen) len = P.seglen;
                if (len < 0) len = 0;
                seg = blk + start;
                if (sg == 0) {
                    bmin = blk[0];
                    bmax = blk[g - 1];
                }
            }
            const InT wmin = warp_min_t<InT>(bmin), wmax = warp_max_t<InT>(bmax);
            const int64_t obase = series * P.o_series + (w0 + wl) * P.o_window;
            for (int j = 0; j < P.n_features; ++j) {
                double v;
                const int f = P.feat[j];
                if (f == MHB_F_MEDIAN) {
                    InT a, c;
                    select_pair<InT>(seg, len, (n & 1) ? (n >> 1) : (n >> 1) - 1, n, wmin, wmax, a, c);
                    v = (n & 1) ? static_cast<double>(a) : (static_cast<double>(a) + static_cast<double>(c)) / 2;
                } else {
                    double acc = 0.0;
                    const int reps = f == MHB_F_IQR ? 2 : 1;
                    for (int rep = 0; rep < reps; ++rep) {
                        const double q = f == MHB_F_IQR ? (rep == 0 ? 75.0 : 25.0) : P.param[j];
                        double pv;
                        if (n == 1 || q == 0.0) {
                            pv = static_cast<double>(wmin);
                        } else if (q == 100.0) {
                            pv = static_cast<double>(wmax);
                        } else {
                            const double rank = 1 + (n - 1) * (q / 100.0);
                            const double fl = floor(rank);
                            const double m = rank - fl;
                            InT a, c;
                            select_pair<InT>(seg, len, static_cast<int>(fl) - 1, n, wmin, wmax, a, c);
                            pv = static_cast<double>(a) * (1 - m) + static_cast<double>(c) * m;
                        }
                        acc = rep == 0 ? pv : acc - pv;
                    }
                    v = acc;
                }
                if (lane == 0) store_cell<OutT>(P.out, obase + j * P.o_col, v);
            }
        }
    }
}

}  // namespace

// Returns -100 when this geometry / feature set is not covered here (the caller then sorts whole windows).
template <typename InT>
int32_t window_order_blocks_try(const InT* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                                const double* h_params, int32_t n_features, const mhb_table* table, void* stream_v) {
    if (n_features <= 0 || n_features > kMaxFeatB) return -100;
    for (int j = 0; j < n_features; ++j)
        if (h_features[j] != MHB_F_MEDIAN && h_features[j] != MHB_F_PERCENTILE && h_features[j] != MHB_F_IQR) return -100;
    const int64_t W = geom->wsize, S = nw == 1 ? geom->wsize : geom->wstep;
    const int64_t g = gcd64(W, S);
    const int64_t k = W / g, hop = S / g;
    if (g < 16 || g > 2048 || k > 32 || hop > 64) return -100;
    if (k > 2 && k < 8) return -100;            // a few blocks per window: sorting the window is cheaper than bisecting
    if (k > 2) {
        // bisection costs one pass of <= 32 (64) probes per selected rank, the window sort is paid once whatever the
        // number of columns: measured (W = 1920, k = 30) 1.5 ms per rank against 18 ms for the sort
        int n_sel = 0;
        for (int j = 0; j < n_features; ++j) n_sel += h_features[j] == MHB_F_IQR ? 2 : 1;
        if (n_sel > 8) return -100;
    }
    int64_t p2 = 32;
    while (p2 < g) p2 <<= 1;
    const int64_t block_bytes = (p2 + 1) * static_cast<int64_t>(sizeof(InT));
    const int64_t nb_max = (32 * 1024) / block_bytes;
    if (nb_max < k) return -100;
    int64_t nwb = (nb_max - k) / hop + 1;
    if (nwb > nw) nwb = nw;
    if (nwb < 1) return -100;
    const int64_t nb = (nwb - 1) * hop + k;
    BlocksPlan P;
    memset(&P, 0, sizeof(P));
    P.x = x;
    P.series_stride = geom->series_stride;
    P.nw = nw;
    P.W = static_cast<int32_t>(W);
    P.g = static_cast<int32_t>(g);
    P.k = static_cast<int32_t>(k);
    P.hop = static_cast<int32_t>(hop);
    P.P2g = static_cast<int32_t>(p2);
    P.nwb = static_cast<int32_t>(nwb);
    int nseg = 1;
    while (nseg * 2 * k <= 32 && nseg * 2 <= g) nseg *= 2;
    P.nseg = nseg;
    P.seglen = static_cast<int32_t>((g + nseg - 1) / nseg);
    P.batches_per_series = (nw + nwb - 1) / nwb;
    P.total_batches = P.batches_per_series * geom->n_series;
    P.out = table->out;
    P.o_series = table->series_stride;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;
    P.n_features = n_features;
    for (int j = 0; j < n_features; ++j) {
        P.feat[j] = h_features[j];
        P.param[j] = h_params ? h_params[j] : 0.0;
    }
    const size_t smem = static_cast<size_t>(nb) * block_bytes;
    cudaError_t e;
    int64_t ctas = static_cast<int64_t>(kNumSMs) * 4;
    if (ctas > P.total_batches) ctas = P.total_batches;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
#define MHB_GOB_E(OUT, E_)                                                                                    \
    {                                                                                                         \
        auto kern = P.k <= 2 ? window_order_blocks_kernel<InT, OUT, true, E_>                                 \
                             : window_order_blocks_kernel<InT, OUT, false, E_>;                               \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  \
        if (e == cudaSuccess) {                                                                               \
            kern<<<static_cast<unsigned>(ctas), kThreadsOB, smem, stream>>>(P);                               \
            e = cudaGetLastError();                                                                           \
        }                                                                                                     \
    }
#define MHB_GOB(OUT)                                                                                          \
    switch (p2) {                                                                                             \
        case 32: MHB_GOB_E(OUT, 1) break;                                                                     \
        case 64: MHB_GOB_E(OUT, 2) break;                                                                     \
        case 128: MHB_GOB_E(OUT, 4) break;                                                                    \
        case 256: MHB_GOB_E(OUT, 8) break;                                                                    \
        case 512: MHB_GOB_E(OUT, 16) break;                                                                   \
        default: MHB_GOB_E(OUT, 0) break;                                                                     \
    }
    if (table->out_f32) MHB_GOB(float) else MHB_GOB(double)
#undef MHB_GOB_E
#undef MHB_GOB
    return cuda_status(e, "window_order_blocks launch");
}

template int32_t window_order_blocks_try<float>(const float*, const mhb_windows*, int64_t, const int32_t*, const double*,
                                                int32_t, const mhb_table*, void*);
template int32_t window_order_blocks_try<double>(const double*, const mhb_windows*, int64_t, const int32_t*, const double*,
                                                 int32_t, const mhb_table*, void*);

}  // namespace mhb
