// Kernel 2 (fast path) -- batched per-window FFT + PSD reducers for the benchmark geometries.
//
// Same contract as window_spectral.cu (which remains the generic path for arbitrary window lengths);
// this file holds the compile-time-planned version the hot configurations use:
//   * a CTA owns a batch of BW consecutive windows of one series.  The samples they cover are staged
//     ONCE in shared memory by a 1-D bulk TMA copy (cp.async.bulk + mbarrier, double-buffered: the
//     tile of the next batch streams in while this batch is transformed), so overlapping windows do
//     not re-read HBM / L2;
//   * the N = W/2 point complex FFT of every window is three register-resident butterflies
//     (R1 x R2 x R3 = N, e.g. 5 x 5 x 10 for W = 500), in place in shared memory, with all index
//     arithmetic folded at compile time; work items of all BW windows are spread over the whole CTA so
//     lanes stay busy whatever N is;
//   * the window mean (float32 estimate) is removed before the transform and bin 0 is restored in
//     float64, exactly as in the generic kernel;
//   * untangling to the one-sided spectrum, band sums, arg-max and the entropy run on register-held
//     PSD values: the PSD row is never written anywhere.
// Reference chain replaced: view (util/windows.py:20-33) -> mhealth.fft.fft (fft/_fft.py:18-29) ->
// |F|^2 -> hrv.power_band / relative_power_band (heart/hrv.py:173-198), density.peak_frequency
// (generic/frequency/density.py:18-32), information.entropy (generic/information.py:10-20).
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "fft_core.cuh"

namespace mhb {

namespace {

constexpr int kMaxColsB = 32;
constexpr int kThreadsB = 256;

struct BatchedPlan {
    const float* x;
    int64_t series_len, series_stride, total_elems, nw;
    int64_t batches_per_series, total_batches;
    int32_t S;
    double bin_hz;
    void* out;
    int32_t out_f32;
    int64_t o_series, o_window, o_col;
    int32_t n_cols;
    int32_t col[kMaxColsB];
    int32_t lo[kMaxColsB], hi[kMaxColsB];
    int32_t use_tma;
};

using C = Cx<float>;

// sum over the 16 lanes of a half-warp (result in every lane of the half)
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- small in-register DFTs (forward, e^{-2 pi i / R})
__device__ __forceinline__ void dft2(C* a) {
    const C t = a[1];
    a[1] = csub(a[0], t);
    a[0] = cadd(a[0], t);
}
__device__ __forceinline__ void dft3(C* a) {
    const float s = 0.86602540378443864676f;
    const C t1 = cadd(a[1], a[2]);
    const C t2 = {a[0].x - 0.5f * t1.x, a[0].y - 0.5f * t1.y};
    const C t3 = cscale(csub(a[1], a[2]), s);
    a[0] = cadd(a[0], t1);
    a[1] = {t2.x + t3.y, t2.y - t3.x};
    a[2] = {t2.x - t3.y, t2.y + t3.x};
}
__device__ __forceinline__ void dft4(C* a) {
    const C t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    const C t2 = cadd(a[1], a[3]), t3 = mul_neg_i(csub(a[1], a[3]));
    a[0] = cadd(t0, t2);
    a[2] = csub(t0, t2);
    a[1] = cadd(t1, t3);
    a[3] = csub(t1, t3);
}
__device__ __forceinline__ void dft5(C* a) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const C p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
    const C p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
    const C a0 = a[0];
    a[0] = {a0.x + p1.x + p2.x, a0.y + p1.y + p2.y};
    const C u1 = {a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y};
    const C u2 = {a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y};
    const C v1 = mul_neg_i(C{s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y});
    const C v2 = mul_neg_i(C{s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y});
    a[1] = cadd(u1, v1);
    a[4] = csub(u1, v1);
    a[2] = cadd(u2, v2);
    a[3] = csub(u2, v2);
}

// constant twiddles exp(-2 pi i m / R) for the composite butterflies
__device__ constexpr float kC8[8] = {1.f, 0.70710678118654752440f, 0.f, -0.70710678118654752440f, -1.f,
                                     -0.70710678118654752440f, 0.f, 0.70710678118654752440f};
__device__ constexpr float kS8[8] = {0.f, -0.70710678118654752440f, -1.f, -0.70710678118654752440f, 0.f,
                                     0.70710678118654752440f, 1.f, 0.70710678118654752440f};
__device__ constexpr float kC10[10] = {1.f, 0.80901699437494742410f, 0.30901699437494742410f, -0.30901699437494742410f,
                                       -0.80901699437494742410f, -1.f, -0.80901699437494742410f, -0.30901699437494742410f,
                                       0.30901699437494742410f, 0.80901699437494742410f};
__device__ constexpr float kS10[10] = {0.f, -0.58778525229247312917f, -0.95105651629515357212f, -0.95105651629515357212f,
                                       -0.58778525229247312917f, 0.f, 0.58778525229247312917f, 0.95105651629515357212f,
                                       0.95105651629515357212f, 0.58778525229247312917f};
__device__ constexpr float kC12[12] = {1.f, 0.86602540378443864676f, 0.5f, 0.f, -0.5f, -0.86602540378443864676f, -1.f,
                                       -0.86602540378443864676f, -0.5f, 0.f, 0.5f, 0.86602540378443864676f};
__device__ constexpr float kS12[12] = {0.f, -0.5f, -0.86602540378443864676f, -1.f, -0.86602540378443864676f, -0.5f, 0.f,
                                       0.5f, 0.86602540378443864676f, 1.f, 0.86602540378443864676f, 0.5f};

template <int R>
__device__ __forceinline__ C const_tw(int m);
template <>
__device__ __forceinline__ C const_tw<8>(int m) { return {kC8[m & 7], kS8[m & 7]}; }
template <>
__device__ __forceinline__ C const_tw<10>(int m) { return {kC10[m % 10], kS10[m % 10]}; }
template <>
__device__ __forceinline__ C const_tw<12>(int m) { return {kC12[m % 12], kS12[m % 12]}; }

template <int R>
__device__ __forceinline__ void dft_small(C* a);
template <>
__device__ __forceinline__ void dft_small<2>(C* a) { dft2(a); }
template <>
__device__ __forceinline__ void dft_small<3>(C* a) { dft3(a); }
template <>
__device__ __forceinline__ void dft_small<4>(C* a) { dft4(a); }
template <>
__device__ __forceinline__ void dft_small<5>(C* a) { dft5(a); }

// Cooley-Tukey composite R = RA * RB in registers:
//   u = RA u2 + u1, t = t2 + RB t1:  y[t2 + RB t1] = sum_u1 w_RA^{u1 t1} w_R^{u1 t2} DFT_RB(a[u1::RA])[t2]
template <int R, int RA, int RB>
__device__ __forceinline__ void dft_composite(C* a) {
    C f[RA][RB];
#pragma unroll
    for (int u1 = 0; u1 < RA; ++u1) {
#pragma unroll
        for (int u2 = 0; u2 < RB; ++u2) f[u1][u2] = a[RA * u2 + u1];
        dft_small<RB>(f[u1]);
#pragma unroll
        for (int t2 = 1; t2 < RB; ++t2)
            if (u1 > 0) f[u1][t2] = cmul(f[u1][t2], const_tw<R>(u1 * t2));
    }
#pragma unroll
    for (int t2 = 0; t2 < RB; ++t2) {
        C g[RA];
#pragma unroll
        for (int u1 = 0; u1 < RA; ++u1) g[u1] = f[u1][t2];
        dft_small<RA>(g);
#pragma unroll
        for (int t1 = 0; t1 < RA; ++t1) a[t2 + RB * t1] = g[t1];
    }
}
template <>
__device__ __forceinline__ void dft_small<8>(C* a) { dft_composite<8, 2, 4>(a); }
template <>
__device__ __forceinline__ void dft_small<10>(C* a) { dft_composite<10, 2, 5>(a); }
template <>
__device__ __forceinline__ void dft_small<12>(C* a) { dft_composite<12, 3, 4>(a); }

// One in-place Stockham pass over all BW windows of the batch.  NS = product of the radices already
// applied.  SRC_TILE: inputs come from the staged samples (first pass; the window mean is removed on the
// fly), otherwise from the FFT buffer.  The batch is processed in two halves of BW/2 windows so that a
// thread never holds more than ceil(BW/2 * N/R / 256) butterflies in registers:
//     read(h0) | barrier | write(h0), read(h1) | barrier | write(h1) | barrier
// (windows are independent, so writing half 0 while reading half 1 is hazard-free).
template <int N, int R, int NS, int HW, bool SRC_TILE, int S>
struct PassHalf {
    static constexpr int NB = N / R;                 // butterflies per window
    static constexpr int ITEMS = HW * NB;
    static constexpr int ROUNDS = (ITEMS + kThreadsB - 1) / kThreadsB;
    static constexpr int TSTEP = N / (NS * R);
    C a[ROUNDS][R];
    int dst[ROUNDS];

    __device__ __forceinline__ void read(const float* __restrict__ tile, const float* __restrict__ mean,
                                         const C* __restrict__ buf, const C* __restrict__ tw, int wbase, int nwin) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int item = threadIdx.x + r * kThreadsB;
            dst[r] = -1;
            if (item < ITEMS) {
                const int wl = item / NB, j = item - wl * NB;
                const int w = wbase + wl;
                if (w < nwin) {
                    const int k = j % NS;
                    if (SRC_TILE) {
                        const float2* z = reinterpret_cast<const float2*>(tile + w * S);
                        const float m = mean[w];
#pragma unroll
                        for (int t = 0; t < R; ++t) {
                            const float2 v = z[j + t * NB];
                            a[r][t] = {v.x - m, v.y - m};
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < R; ++t) a[r][t] = buf[w * N + j + t * NB];
                    }
                    if (NS > 1) {
#pragma unroll
                        for (int t = 1; t < R; ++t) a[r][t] = cmul(a[r][t], tw[t * k * TSTEP]);
                    }
                    dst[r] = w * N + (j - k) * R + k;
                }
            }
        }
    }
    __device__ __forceinline__ void write(C* __restrict__ buf) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            if (dst[r] >= 0) {
                dft_small<R>(a[r]);
#pragma unroll
                for (int t = 0; t < R; ++t) buf[dst[r] + t * NS] = a[r][t];
            }
        }
    }
};

template <int N, int R, int NS, int BW, bool SRC_TILE, int S>
__device__ __forceinline__ void batched_pass(const float* __restrict__ tile, const float* __restrict__ mean,
                                             C* __restrict__ buf, const C* __restrict__ tw, int nwin) {
    constexpr int HW = BW / 2;
    PassHalf<N, R, NS, HW, SRC_TILE, S> h;
    h.read(tile, mean, buf, tw, 0, nwin);
    if (!SRC_TILE) __syncthreads();                  // half 0: every input has been read
    h.write(buf);
    h.read(tile, mean, buf, tw, HW, nwin);
    if (!SRC_TILE) __syncthreads();                  // half 1: every input has been read
    h.write(buf);
    __syncthreads();
}

template <int W, int S, int BW, int R1, int R2, int R3, int MINB>
__global__ void __launch_bounds__(kThreadsB, MINB) spectral_batched_kernel(const BatchedPlan P) {
    constexpr int N = W / 2;
    static_assert(R1 * R2 * R3 == N, "radix plan must multiply to W/2");
    static_assert(BW == kThreadsB / 16, "one half-warp reduces one window of the batch");
    constexpr int NPAIR = N / 2;                     // pairs (k, N-k), k = 1..N/2
    constexpr int PROUNDS = (NPAIR + 15) / 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // 2 barriers
    float* tiles = reinterpret_cast<float*>(smem_raw + 128);                // 2 x tile_elems
    constexpr int tile_elems = (((BW - 1) * S + W + 3) & ~3) + 4;
    C* buf = reinterpret_cast<C*>(tiles + 2 * tile_elems);                  // BW x N complex
    C* tw = buf + BW * N;                                                   // N
    C* tw2 = tw + N;                                                        // N/2 + 1
    float* mean = reinterpret_cast<float*>(tw2 + (N / 2 + 1));              // BW
    double* colres = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(mean + BW) + 7) & ~uintptr_t(7));   // BW x kMaxColsB
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    fill_twiddles<float>(tw, N, N, tid, kThreadsB);
    fill_twiddles<float>(tw2, W, N / 2 + 1, tid, kThreadsB);
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto batch_geom = [&](int64_t b, int64_t& series, int64_t& w0, int& nwin, int64_t& goff, int& n_valid) {
        const uint32_t bps = static_cast<uint32_t>(P.batches_per_series);      // host guarantees < 2^31 batches
        const uint32_t sr = static_cast<uint32_t>(b) / bps;
        const uint32_t bi = static_cast<uint32_t>(b) - sr * bps;
        series = sr;
        w0 = static_cast<int64_t>(bi) * BW;
        const int64_t left = P.nw - w0;
        nwin = left < BW ? static_cast<int>(left) : BW;
        goff = series * P.series_stride + w0 * S;
        n_valid = (nwin - 1) * S + W;
    };
    auto issue = [&](int64_t b, int slot) {              // thread 0 only
        int64_t series, w0, goff;
        int nwin, n_valid;
        batch_geom(b, series, w0, nwin, goff, n_valid);
        const int n_load = (n_valid + 3) & ~3;
        if (P.use_tma && (goff % 4 == 0) && (goff + n_load <= P.total_elems)) {
            mbar_arrive_expect_tx(&full[slot], n_load * 4);
            bulk_g2s(tiles + slot * tile_elems, P.x + goff, n_load * 4, &full[slot]);
        }
    };

    int slot = 0;
    uint32_t parity0 = 0, parity1 = 0;
    int64_t b = blockIdx.x;
    if (tid == 0 && b < P.total_batches) issue(b, 0);

    for (; b < P.total_batches; b += gridDim.x) {
        int64_t series, w0, goff;
        int nwin, n_valid;
        batch_geom(b, series, w0, nwin, goff, n_valid);
        float* tile = tiles + slot * tile_elems;
        const int n_load = (n_valid + 3) & ~3;
        const bool via_tma = P.use_tma && (goff % 4 == 0) && (goff + n_load <= P.total_elems);
        if (via_tma) {
            if (slot == 0) {
                mbar_wait(&full[0], parity0);
                parity0 ^= 1;
            } else {
                mbar_wait(&full[1], parity1);
                parity1 ^= 1;
            }
        } else {
            __syncthreads();                            // previous readers of this slot are done
            for (int i = tid; i < n_valid; i += kThreadsB) tile[i] = P.x[goff + i];
            __syncthreads();
        }
        // prefetch the next batch's tile into the other slot: its last readers (pass 1 of the previous
        // iteration) are separated from here by that iteration's barriers
        const int64_t bn = b + gridDim.x;
        if (tid == 0 && bn < P.total_batches) issue(bn, slot ^ 1);

        // ---- float32 mean estimate per window (pivot of the transform): half-warp `hw` owns window `hw`
        const int l16 = lane & 15;
        const int hw = warp * 2 + (lane >> 4);           // window of this half-warp, 0..BW-1
        {
            float sm = 0.f;
            if (hw < nwin) {
                const float2* z = reinterpret_cast<const float2*>(tile + hw * S);
                for (int i = l16; i < N; i += 16) {
                    const float2 v = z[i];
                    sm += v.x + v.y;
                }
            }
            sm = half_sum(sm);
            if (l16 == 0 && hw < nwin) mean[hw] = sm * (1.0f / W);
        }
        __syncthreads();

        // ---- FFT: three in-place passes, items of all windows spread over the CTA
        batched_pass<N, R1, 1, BW, true, S>(tile, mean, buf, tw, nwin);
        batched_pass<N, R2, R1, BW, false, S>(tile, mean, buf, tw, nwin);
        if (R3 > 1) batched_pass<N, R3, R1 * R2, BW, false, S>(tile, mean, buf, tw, nwin);

        // ---- epilogue, one half-warp per window: untangle -> PSD (registers + shared row), reductions by shuffles
        {
            const bool act = hw < nwin;
            const int hwc = act ? hw : 0;
            const C* Z = buf + hwc * N;
            float* prow = reinterpret_cast<float*>(buf + hwc * N);         // PSD row overlays the window's FFT buffer
            float psd_lo[PROUNDS], psd_hi[PROUNDS];      // |X[k]|^2, |X[N-k]|^2 for k = 1 + l16 + 16 i
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < PROUNDS; ++i) {
                const int k = 1 + l16 + 16 * i;
                psd_lo[i] = psd_hi[i] = 0.f;
                if (act && k <= NPAIR) {
                    const C zk = Z[k], zn = Z[N - k];
                    const C e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
                    const C o = {0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};
                    const C t = cmul(o, tw2[k]);
                    const float ar = e.x + t.x, ai = e.y + t.y, br = e.x - t.x, bi = e.y - t.y;
                    psd_lo[i] = ar * ar + ai * ai;
                    if (k != N - k) psd_hi[i] = br * br + bi * bi;
                    tot += psd_lo[i] + psd_hi[i];
                }
            }
            const C z0 = Z[0];
            const double x0 = static_cast<double>(z0.x) + static_cast<double>(z0.y) +
                              static_cast<double>(W) * static_cast<double>(mean[hwc]);
            const double dc = x0 * x0;                   // exact DC: FFT(x - m)[0] + W m
            const float xn = z0.x - z0.y;
            const float nyq = xn * xn;                   // bin N
            __syncwarp();                                // every lane has read its Z values
#pragma unroll
            for (int i = 0; i < PROUNDS; ++i) {
                const int k = 1 + l16 + 16 * i;
                if (act && k <= NPAIR) {
                    prow[k] = psd_lo[i];
                    if (k != N - k) prow[N - k] = psd_hi[i];
                }
            }
            if (act && l16 == 0) {
                prow[0] = static_cast<float>(dc);
                prow[N] = nyq;
            }
            const double rest = static_cast<double>(half_sum(tot)) + static_cast<double>(nyq);   // all bins but 0
            const double total = rest + dc;
            const double inv_total = 1.0 / total;
            // entropy from the register-held values; the DC term through log1p (p0 may be within 1e-7 of 1)
            const float inv = static_cast<float>(inv_total);
            float h = 0.f;
#pragma unroll
            for (int i = 0; i < PROUNDS; ++i) {
                const int k = 1 + l16 + 16 * i;
                if (k <= NPAIR) {
                    const float q1 = psd_lo[i] * inv + 1e-30f;
                    h += q1 * __logf(q1);
                    if (k != N - k) {
                        const float q2 = psd_hi[i] * inv + 1e-30f;
                        h += q2 * __logf(q2);
                    }
                }
            }
            if (l16 == 0) {
                const float q = nyq * inv + 1e-30f;
                h += q * __logf(q);
                const float qrest = static_cast<float>(rest * inv_total);      // 1 - p0, formed in float64
                const float p0 = static_cast<float>(dc * inv_total);
                h += p0 * log1pf(-qrest);
            }
            const float hs = half_sum(h);
            __syncwarp();                                // PSD row visible to the half-warp
            double* res = colres + hwc * kMaxColsB;
#pragma unroll 1
            for (int j = 0; j < P.n_cols; ++j) {
                const int kind = P.col[j];
                const int lo = P.lo[j], hi = P.hi[j];
                double v;
                if (kind == MHB_S_TOTAL_POWER) {
                    v = total;
                } else if (kind == MHB_S_ENTROPY) {
                    v = -static_cast<double>(hs);
                } else if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
                    float acc = 0.f;
                    for (int k = (lo > 1 ? lo : 1) + l16; k < hi; k += 16) acc += prow[k];
                    double bsum = static_cast<double>(half_sum(acc));
                    if (lo <= 0 && hi > 0) bsum += dc;
                    v = kind == MHB_S_BAND_POWER ? bsum : bsum * inv_total;
                } else {                                  // peak frequency / bin: first maximum in [lo, hi)
                    float best = -1.f;
                    int arg = 0x7fffffff;
                    for (int k = lo + l16; k < hi; k += 16) {
                        const float pv = prow[k];
                        if (pv > best) {
                            best = pv;
                            arg = k;
                        }
                    }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                        if (ob > best || (ob == best && oa < arg)) {
                            best = ob;
                            arg = oa;
                        }
                    }
                    if (arg == 0x7fffffff) v = CUDART_NAN;
                    else v = kind == MHB_S_PEAK_BIN ? static_cast<double>(arg) : static_cast<double>(arg) * P.bin_hz;
                }
                if (act && l16 == 0) res[j] = v;      // idle half-warps alias window 0: they must not store
            }
            __syncwarp();
            if (act && l16 < P.n_cols) {                 // lane j stores column j: one coalesced row segment
                const int64_t o = series * P.o_series + (w0 + hw) * P.o_window + l16 * P.o_col;
                const double v = res[l16];
                if (P.out_f32) reinterpret_cast<float*>(P.out)[o] = static_cast<float>(v);
                else reinterpret_cast<double*>(P.out)[o] = v;
            }
            if (act && P.n_cols > 16) {
                for (int j = 16 + l16; j < P.n_cols; j += 16) {
                    const int64_t o = series * P.o_series + (w0 + hw) * P.o_window + j * P.o_col;
                    if (P.out_f32) reinterpret_cast<float*>(P.out)[o] = static_cast<float>(res[j]);
                    else reinterpret_cast<double*>(P.out)[o] = res[j];
                }
            }
        }
        // no barrier here: the next batch writes `mean` / `buf` only after its own mean-phase barrier, which
        // every warp reaches after finishing this epilogue; `mean` is rewritten by the warp that just read it
        slot ^= 1;
    }
}

template <int W, int BW>
size_t batched_smem_bytes(int S) {
    constexpr int N = W / 2;
    const size_t tile_elems = (((BW - 1) * S + W + 3) & ~3) + 4;
    return 128 + 2 * tile_elems * 4 + sizeof(C) * (static_cast<size_t>(BW) * N + N + N / 2 + 1) + 4 * BW +
           8 * BW * kMaxColsB + 64;
}

template <int W, int S, int BW, int R1, int R2, int R3, int MINB>
int32_t launch_batched(const BatchedPlan& P, void* stream) {
    auto kern = spectral_batched_kernel<W, S, BW, R1, R2, R3, MINB>;
    const size_t smem = batched_smem_bytes<W, BW>(P.S);
    if (smem > 200 * 1024) return -100;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e, "spectral_batched attr");
    int per_sm = static_cast<int>((224 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > MINB) per_sm = MINB;                            // __launch_bounds__(256, MINB)
    int64_t ctas = static_cast<int64_t>(kNumSMs) * per_sm;
    if (ctas > P.total_batches) ctas = P.total_batches;
    kern<<<static_cast<unsigned>(ctas), kThreadsB, smem, static_cast<cudaStream_t>(stream)>>>(P);
    return cuda_status(cudaGetLastError(), "spectral_batched launch");
}

}  // namespace

// Returns -100 when the geometry has no compile-time plan (the caller then uses the generic kernel).
int32_t spectral_batched_try(const float* x, const mhb_windows* geom, int64_t nw, double bin_hz, const int32_t* cols,
                             const int32_t* lo, const int32_t* hi, int32_t n_cols, void* out, int32_t out_f32,
                             int64_t o_series, int64_t o_window, int64_t o_col, void* stream) {
    if (n_cols <= 0 || n_cols > kMaxColsB) return -100;
    if (geom->wstep % 2 != 0 || geom->wstep > geom->wsize) return -100;
    BatchedPlan P;
    memset(&P, 0, sizeof(P));
    P.x = x;
    P.series_len = geom->series_len;
    P.series_stride = geom->series_stride;
    P.total_elems = (geom->n_series - 1) * geom->series_stride + geom->series_len;
    P.nw = nw;
    P.S = geom->wstep;
    P.bin_hz = bin_hz;
    P.out = out;
    P.out_f32 = out_f32;
    P.o_series = o_series;
    P.o_window = o_window;
    P.o_col = o_col;
    P.n_cols = n_cols;
    for (int j = 0; j < n_cols; ++j) {
        P.col[j] = cols[j];
        P.lo[j] = lo[j];
        P.hi[j] = hi[j];
    }
    P.use_tma = (reinterpret_cast<uintptr_t>(x) % 16 == 0 && geom->series_stride % 4 == 0) ? 1 : 0;
    if (geom->wsize == 500 && geom->wstep == 250) {
        constexpr int BW = 16;
        P.batches_per_series = (nw + BW - 1) / BW;
        P.total_batches = P.batches_per_series * geom->n_series;
        return launch_batched<500, 250, BW, 5, 5, 10, 2>(P, stream);     // 128 regs, 2 CTAs/SM: no spills (3 CTAs spill)
    }
    // W = 1920 (config 4): the in-place passes would need BW * 960 / 256 complex values per thread in
    // registers; that geometry stays on the generic kernel until a ping-pong variant is written.
    return -100;
}

}  // namespace mhb
