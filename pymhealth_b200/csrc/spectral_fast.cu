// Kernel 2 (fast path, W = 500 / S = 250) -- per-window FFT + PSD reducers with thread-resident sub-transforms.
//
// Same contract as window_spectral.cu (the generic path for arbitrary window lengths).  Reference chain
// replaced: view (util/windows.py:20-33) -> mhealth.fft.fft (fft/_fft.py:18-29) -> |F|^2 ->
// hrv.power_band / relative_power_band (heart/hrv.py:173-198), density.peak_frequency
// (generic/frequency/density.py:18-32), information.entropy (generic/information.py:10-20).
//
// A CTA (7 warps) owns a batch of BW = 16 consecutive windows of one series; their samples arrive ONCE by a
// 1-D bulk TMA copy (cp.async.bulk + mbarrier, double buffered).  The real W-point DFT is an N = W/2 = 250
// point complex FFT, factored N = 25 x 10 so that the whole transform is TWO register-resident passes with
// ONE exchange through shared memory:
//   pass A  thread (window w, n2 = 0..9):  25-point DFT (5 x 5 in registers) over z[10 n1 + n2], times
//           w250^(n2 k1), stored to buf[w][n2][k1];
//   pass B  thread (window w, p = 0..12):  the two 10-point DFTs (2 x 5) that give Z[p + 25 k2] and
//           Z[(25 - p) + 25 k2] -- exactly the pairs (k, N - k) the real-input untangling needs -- so the
//           one-sided spectrum, |X|^2, the band sums and the arg-max are formed on REGISTER values; the PSD
//           row is never written anywhere.
// Lanes of a half-warp hold the 16 windows of the batch (same role), so every twiddle / mask load is a
// broadcast and every window-strided access is conflict free (250 and 2*251 words are = 26, 22 mod 32).
// Roles are half-warp uniform: a warp that has no pass-A (or pass-B) work idles as a WARP, it never burns
// issue slots on masked lanes.
// The window mean (float32 estimate) is removed before the transform and bin 0 is restored in float64
// (FFT(x - m)[0] + W m), exactly as in the generic kernel; the samples are also pre-scaled by 1/2 (exact),
// which absorbs the 1/2 of the untangling step.
// Cross-thread reductions (13 partials per window and quantity) go through shared memory in a fixed order:
// results are deterministic.
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "fft_consts.cuh"
#include "fft_core.cuh"

namespace mhb {

namespace {

using C = Cx<float>;

constexpr int kW = 500, kS = 250, kN = 250;
constexpr int kBW = 16;                  // windows per batch
constexpr int kThreadsF = 224;           // 14 half-warps
constexpr int kNA = 10;                  // pass-A threads per window (n2)
constexpr int kNP = 13;                  // pass-B threads per window (p = 0..12)
constexpr int kWSB = 251;                // complex stride between windows in buf (odd: conflict free)
constexpr int kMaxColsF = 32;
constexpr int kMaxSum = 6;               // distinct band ranges per call
constexpr int kMaxArg = 4;               // distinct arg-max ranges per call
constexpr int kTileElems = ((kBW - 1) * kS + kW + 3 + 4) & ~3;     // 4256 floats
constexpr int kMeanParts = 7;

struct FastPlan {
    const float* x;
    int64_t series_stride, total_elems, nw;
    int64_t batches_per_series, total_batches;
    double bin_hz;
    void* out;
    int32_t out_f32;
    int64_t o_series, o_window, o_col;
    int32_t n_cols, n_sum, n_arg;
    int32_t col[kMaxColsF];              // MHB_S_* kind
    int32_t cref[kMaxColsF];             // index into the sum / arg range tables
    int32_t sum_lo[kMaxSum], sum_hi[kMaxSum];
    int32_t arg_lo[kMaxArg], arg_hi[kMaxArg];
    int32_t use_tma;
};

// ---- small in-register DFTs (forward, e^{-2 pi i / R})
__device__ __forceinline__ void dft2(C* a) {
    const C t = a[1];
    a[1] = csub(a[0], t);
    a[0] = cadd(a[0], t);
}
__device__ __forceinline__ void dft5(C* a) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const C p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
    const C p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
    const C a0 = a[0];
    a[0] = {a0.x + p1.x + p2.x, a0.y + p1.y + p2.y};
    const C u1 = {fmaf(c2, p2.x, fmaf(c1, p1.x, a0.x)), fmaf(c2, p2.y, fmaf(c1, p1.y, a0.y))};
    const C u2 = {fmaf(c1, p2.x, fmaf(c2, p1.x, a0.x)), fmaf(c1, p2.y, fmaf(c2, p1.y, a0.y))};
    const C v1 = mul_neg_i(C{fmaf(s2, m2.x, s1 * m1.x), fmaf(s2, m2.y, s1 * m1.y)});
    const C v2 = mul_neg_i(C{fmaf(-s1, m2.x, s2 * m1.x), fmaf(-s1, m2.y, s2 * m1.y)});
    a[1] = cadd(u1, v1);
    a[4] = csub(u1, v1);
    a[2] = cadd(u2, v2);
    a[3] = csub(u2, v2);
}

template <int R>
__device__ __forceinline__ C ctw(int m);
template <>
__device__ __forceinline__ C ctw<10>(int m) { return {kCos10[m % 10], kNSin10[m % 10]}; }
template <>
__device__ __forceinline__ C ctw<25>(int m) { return {kCos25[m % 25], kNSin25[m % 25]}; }

template <int R>
__device__ __forceinline__ void dft_prime(C* a);
template <>
__device__ __forceinline__ void dft_prime<2>(C* a) { dft2(a); }
template <>
__device__ __forceinline__ void dft_prime<5>(C* a) { dft5(a); }

// Cooley-Tukey composite R = RA * RB in registers, natural order in and out:
//   u = RA u2 + u1, t = t2 + RB t1:  y[t] = sum_u1 w_RA^{u1 t1} w_R^{u1 t2} DFT_RB(a[u1::RA])[t2]
template <int R, int RA, int RB>
__device__ __forceinline__ void dft_composite(C* a) {
    C f[RA][RB];
#pragma unroll
    for (int u1 = 0; u1 < RA; ++u1) {
#pragma unroll
        for (int u2 = 0; u2 < RB; ++u2) f[u1][u2] = a[RA * u2 + u1];
        dft_prime<RB>(f[u1]);
#pragma unroll
        for (int t2 = 1; t2 < RB; ++t2)
            if (u1 > 0) f[u1][t2] = cmul(f[u1][t2], ctw<R>(u1 * t2));
    }
#pragma unroll
    for (int t2 = 0; t2 < RB; ++t2) {
        C g[RA];
#pragma unroll
        for (int u1 = 0; u1 < RA; ++u1) g[u1] = f[u1][t2];
        dft_prime<RA>(g);
#pragma unroll
        for (int t1 = 0; t1 < RA; ++t1) a[t2 + RB * t1] = g[t1];
    }
}
__device__ __forceinline__ void dft10(C* a) { dft_composite<10, 2, 5>(a); }
__device__ __forceinline__ void dft25(C* a) { dft_composite<25, 5, 5>(a); }

// shared-memory layout of the reduction scratch that overlays the consumed tile slot
struct Scratch {
    float* ptot;      // [BW][NP]  sum of the thread's bins (bin 0 excluded)
    float* ph;        // [BW][NP]  entropy partial (natural log units)
    float* psum;      // [kMaxSum][BW][NP]
    float* pbest;     // [kMaxArg][BW][NP]
    int* parg;        // [kMaxArg][BW][NP]
    double* dc;       // [BW]     exact bin 0
    __device__ __forceinline__ void carve(float* base) {
        constexpr int Q = kBW * kNP;
        dc = reinterpret_cast<double*>(base);
        float* p = base + 2 * kBW;
        ptot = p; p += Q;
        ph = p; p += Q;
        psum = p; p += kMaxSum * Q;
        pbest = p; p += kMaxArg * Q;
        parg = reinterpret_cast<int*>(p);
    }
};
static_assert(2 * kBW + (2 + kMaxSum + 2 * kMaxArg) * kBW * kNP <= kTileElems, "scratch must fit a tile slot");

// bit k2 of the mask: bin (lowset ? p + 25 k2 : 250 - p - 25 k2) lies in [lo, hi)
__device__ __forceinline__ uint32_t range_mask(int p, bool lowset, int lo, int hi) {
    uint32_t m = 0;
#pragma unroll
    for (int k2 = 0; k2 < 10; ++k2) {
        const int k = lowset ? p + 25 * k2 : kN - p - 25 * k2;
        if (k >= lo && k < hi) m |= 1u << k2;
    }
    return m;
}

__global__ void __launch_bounds__(kThreadsF, 3) spectral_fast_kernel(const FastPlan P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // 2 barriers
    float* tiles = reinterpret_cast<float*>(smem_raw + 128);                // 2 x kTileElems
    C* buf = reinterpret_cast<C*>(tiles + 2 * kTileElems);                  // BW x kWSB complex
    C* twA = buf + kBW * kWSB;                                              // [25][10]  w250^(n2 k1)
    C* twB = twA + 25 * kNA;                                                // [10][13]  w500^(p + 25 k2)
    float* msum = reinterpret_cast<float*>(twB + 10 * kNP);                 // [kMeanParts][BW]
    float* piv = msum + kMeanParts * kBW;                                   // [BW] pivot (mean estimate) per window
    uint32_t* masks = reinterpret_cast<uint32_t*>(piv + kBW);               // [kMaxSum + kMaxArg][2][16]
    const int tid = threadIdx.x;
    const int w = tid & 15;              // window of the batch this thread works for (all phases)
    const int role = tid >> 4;           // half-warp index 0..13

    for (int i = tid; i < 25 * kNA; i += kThreadsF) {
        const int k1 = i / kNA, n2 = i - k1 * kNA;
        double s, c;
        sincospi(-2.0 * static_cast<double>(k1 * n2) / 250.0, &s, &c);
        twA[i] = {static_cast<float>(c), static_cast<float>(s)};
    }
    for (int i = tid; i < 10 * kNP; i += kThreadsF) {
        const int k2 = i / kNP, p = i - k2 * kNP;
        double s, c;
        sincospi(-2.0 * static_cast<double>(p + 25 * k2) / 500.0, &s, &c);
        twB[i] = {static_cast<float>(c), static_cast<float>(s)};
    }
    for (int i = tid; i < (kMaxSum + kMaxArg) * 2 * 16; i += kThreadsF) {
        const int r = i >> 5, lowset = (i >> 4) & 1, p = i & 15;
        const bool is_sum = r < kMaxSum;
        const int ri = is_sum ? r : r - kMaxSum;
        const bool live = is_sum ? ri < P.n_sum : ri < P.n_arg;
        const int lo = is_sum ? P.sum_lo[ri] : P.arg_lo[ri], hi = is_sum ? P.sum_hi[ri] : P.arg_hi[ri];
        masks[i] = (live && p < kNP) ? range_mask(p, lowset != 0, lo, hi) : 0u;
    }
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
        w0 = static_cast<int64_t>(bi) * kBW;
        const int64_t left = P.nw - w0;
        nwin = left < kBW ? static_cast<int>(left) : kBW;
        goff = series * P.series_stride + w0 * kS;
        n_valid = (nwin - 1) * kS + kW;
    };
    auto tma_ok = [&](int64_t goff, int n_load) {
        return P.use_tma && (goff % 4 == 0) && (goff + n_load <= P.total_elems);
    };
    auto issue = [&](int64_t b, int slot) {              // thread 0 only
        int64_t series, w0, goff;
        int nwin, n_valid;
        batch_geom(b, series, w0, nwin, goff, n_valid);
        const int n_load = (n_valid + 3) & ~3;
        if (tma_ok(goff, n_load)) {
            fence_proxy_async();                          // earlier generic-proxy accesses of the slot are done
            mbar_arrive_expect_tx(&full[slot], n_load * 4);
            bulk_g2s(tiles + slot * kTileElems, P.x + goff, n_load * 4, &full[slot]);
        }
    };

    int slot = 0;
    uint32_t parity0 = 0, parity1 = 0;
    int64_t b = blockIdx.x;
    if (tid == 0) {
        if (b < P.total_batches) issue(b, 0);
        if (b + gridDim.x < P.total_batches) issue(b + gridDim.x, 1);
    }
    bool first = true;

    for (; b < P.total_batches; b += gridDim.x) {
        int64_t series, w0, goff;
        int nwin, n_valid;
        batch_geom(b, series, w0, nwin, goff, n_valid);
        float* tile = tiles + slot * kTileElems;
        const int n_load = (n_valid + 3) & ~3;
        if (tma_ok(goff, n_load)) {
            if (slot == 0) {
                mbar_wait(&full[0], parity0);
                parity0 ^= 1;
            } else {
                mbar_wait(&full[1], parity1);
                parity1 ^= 1;
            }
        } else {
            // unaligned base or the last few samples of the buffer: guarded cooperative copy.  The slot's previous
            // readers (the finalize phase two batches ago) are behind at least one barrier of the last iteration.
            for (int i = tid; i < n_valid; i += kThreadsF) tile[i] = P.x[goff + i];
            __syncthreads();
        }
        const bool act = w < nwin;

        // ---- window mean estimate (pivot of the transform): 14 threads per window, every other float2
        {
            float sm = 0.f;
            if (act) {
                const float2* z = reinterpret_cast<const float2*>(tile + w * kS) + role * 18;
                const int cnt = role == 13 ? 8 : 9;       // float2 pairs [36 role, 36 role + 36) within 250
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    if (i < cnt) {
                        const float2 v = z[2 * i];
                        sm += v.x + v.y;
                    }
            }
            sm += __shfl_xor_sync(0xffffffffu, sm, 16);    // roles 2j and 2j + 1 share a warp
            if ((role & 1) == 0) msum[(role >> 1) * kBW + w] = sm;
        }
        __syncthreads();                                   // (B1) also: everyone finished the previous batch
        // refill the slot the previous batch used (its reduction scratch is dead now)
        if (tid == 0 && !first) {
            const int64_t bn = b + gridDim.x;
            if (bn < P.total_batches) issue(bn, slot ^ 1);
        }
        first = false;

        // ---- pass A: 25-point DFTs of the stride-10 subsequences, inter-pass twiddle, exchange
        if (role < kNA && act) {
            const int n2 = role;
            float m = 0.f;
#pragma unroll
            for (int j = 0; j < kMeanParts; ++j) m += msum[j * kBW + w];
            m *= (1.0f / 250.0f);                          // 250 of the 500 samples were summed
            const float mh = -0.5f * m;
            const float2* z = reinterpret_cast<const float2*>(tile + w * kS) + n2;
            C a[25];
#pragma unroll
            for (int n1 = 0; n1 < 25; ++n1) {
                const float2 v = z[kNA * n1];
                a[n1] = {fmaf(v.x, 0.5f, mh), fmaf(v.y, 0.5f, mh)};
            }
            dft25(a);
            C* dst = buf + w * kWSB + n2 * 25;
            dst[0] = a[0];
            if (n2 == 0) {
#pragma unroll
                for (int k1 = 1; k1 < 25; ++k1) dst[k1] = a[k1];
            } else {
                const C* tw = twA + n2;
#pragma unroll
                for (int k1 = 1; k1 < 25; ++k1) dst[k1] = cmul(a[k1], tw[k1 * kNA]);
            }
            if (n2 == 0) piv[w] = m;
        }
        __syncthreads();                                   // (B2)

        // ---- pass B phase 1: 10-point DFTs -> untangle -> |X|^2 in registers -> partial reductions
        Scratch sc;
        sc.carve(tile);                                    // the tile was consumed by pass A
        float psd_l[10], psd_h[10];                        // p >= 1: bins p + 25 k2 / 250 - p - 25 k2
        const bool has_b = role < kNP - 1;                 // p = role + 1 = 1..12
        const bool has_0 = role == kNP - 1;                // p = 0 lives alone in the lower half of warp 6
        const int p = has_b ? role + 1 : 0;
        float pivot = 0.f;
        if ((has_b || has_0) && act) {
            const C* bw = buf + w * kWSB;
            pivot = piv[w];
            C A[10];
#pragma unroll
            for (int n2 = 0; n2 < 10; ++n2) A[n2] = bw[n2 * 25 + p];
            dft10(A);                                      // A[k2] = Z[p + 25 k2] / 2
            float tot = 0.f;
            if (has_b) {
                C B[10];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) B[n2] = bw[n2 * 25 + 25 - p];
                dft10(B);                                  // B[k2] = Z[25 - p + 25 k2] / 2
#pragma unroll
                for (int k2 = 0; k2 < 10; ++k2) {
                    const C zk = A[k2], zn = B[9 - k2];   // (k, N - k), k = p + 25 k2
                    const C t2 = twB[k2 * kNP + p];
                    const C e = {zk.x + zn.x, zk.y - zn.y};
                    const C o = {zk.y + zn.y, zn.x - zk.x};
                    const C t = cmul(o, t2);
                    const float ar = e.x + t.x, ai = e.y + t.y, br = e.x - t.x, bi = e.y - t.y;
                    psd_l[k2] = fmaf(ar, ar, ai * ai);
                    psd_h[k2] = fmaf(br, br, bi * bi);
                    tot += psd_l[k2] + psd_h[k2];
                }
            } else {
                // p = 0: Z[25 k2]; pairs (k2, 10 - k2) for k2 = 1..4, the self pair k2 = 5, and bins 0 / N from Z[0]
#pragma unroll
                for (int k2 = 0; k2 < 10; ++k2) psd_l[k2] = psd_h[k2] = 0.f;
#pragma unroll
                for (int k2 = 1; k2 <= 5; ++k2) {
                    const C zk = A[k2], zn = A[10 - k2];
                    const C t2 = twB[k2 * kNP];
                    const C e = {zk.x + zn.x, zk.y - zn.y};
                    const C o = {zk.y + zn.y, zn.x - zk.x};
                    const C t = cmul(o, t2);
                    const float ar = e.x + t.x, ai = e.y + t.y, br = e.x - t.x, bi = e.y - t.y;
                    psd_l[k2] = fmaf(ar, ar, ai * ai);                    // bin 25 k2
                    if (k2 < 5) psd_h[k2] = fmaf(br, br, bi * bi);        // bin 250 - 25 k2
                    tot += psd_l[k2] + psd_h[k2];
                }
                const float xn = 2.f * (A[0].x - A[0].y);
                psd_h[0] = xn * xn;                                       // bin N = 250 (Nyquist)
                tot += psd_h[0];
                const double x0 = 2.0 * (static_cast<double>(A[0].x) + static_cast<double>(A[0].y)) +
                                  static_cast<double>(kW) * static_cast<double>(pivot);
                sc.dc[w] = x0 * x0;                                       // exact DC: FFT(x - m)[0] + W m
            }
            const int q = w * kNP + p;
            sc.ptot[q] = tot;
            // band sums over the thread's bins (bin 0 is added by the finalize step from sc.dc)
#pragma unroll 1
            for (int r = 0; r < P.n_sum; ++r) {
                const uint32_t ml = masks[(r * 2 + 1) * 16 + p], mh2 = masks[(r * 2) * 16 + p];
                float acc = 0.f;
#pragma unroll
                for (int k2 = 0; k2 < 10; ++k2) {
                    if (ml & (1u << k2)) acc += psd_l[k2];
                    if (mh2 & (1u << k2)) acc += psd_h[k2];
                }
                sc.psum[r * kBW * kNP + q] = acc;
            }
            // first maximum over the thread's bins, visited in ascending bin order:
            //   p, 25 - p, p + 25, 50 - p, ...   (low set k2 = i, high set k2 = 9 - i)
#pragma unroll 1
            for (int r = 0; r < P.n_arg; ++r) {
                const uint32_t ml = masks[((kMaxSum + r) * 2 + 1) * 16 + p], mh2 = masks[((kMaxSum + r) * 2) * 16 + p];
                float best = -1.f;
                int arg = 0x7fffffff;
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    if (has_b || i > 0) {                                 // p = 0: bin 0 is handled by the finalize step
                        if ((ml & (1u << i)) && psd_l[i] > best) {
                            best = psd_l[i];
                            arg = p + 25 * i;
                        }
                    }
                    if ((mh2 & (1u << (9 - i))) && psd_h[9 - i] > best) {
                        best = psd_h[9 - i];
                        arg = kN - p - 25 * (9 - i);
                    }
                }
                sc.pbest[r * kBW * kNP + q] = best;
                sc.parg[r * kBW * kNP + q] = arg;
            }
        }
        __syncthreads();                                   // (B3)

        // ---- pass B phase 2: entropy of the normalised PSD from the register values
        if ((has_b || has_0) && act) {
            float rest = 0.f;
#pragma unroll
            for (int j = 0; j < kNP; ++j) rest += sc.ptot[w * kNP + j];
            const double dc = sc.dc[w];
            const float total = rest + static_cast<float>(dc);
            const float inv = __fdividef(1.0f, total);
            float h = 0.f;                                  // sum q log2 q
#pragma unroll
            for (int k2 = 0; k2 < 10; ++k2) {
                if (has_b || (k2 >= 1 && k2 <= 5)) {
                    const float q1 = fmaf(psd_l[k2], inv, 1e-30f);
                    h = fmaf(q1, __log2f(q1), h);
                }
                if (has_b || k2 <= 4) {                      // p = 0: bins 250 - 25 k2 (k2 = 1..4) and the Nyquist bin
                    const float q2 = fmaf(psd_h[k2], inv, 1e-30f);
                    h = fmaf(q2, __log2f(q2), h);
                }
            }
            h *= 0.69314718055994530942f;
            if (has_0) {
                // the DC term through log1p in float64: p0 may be within 1e-7 of 1 (gravity axis)
                const double tot64 = static_cast<double>(rest) + dc;
                const double p0 = dc / tot64, qrest = static_cast<double>(rest) / tot64;
                h += static_cast<float>((p0 + 1e-30) * (qrest < 0.5 ? log1p(-qrest) : log(p0 + 1e-30)));
            }
            sc.ph[w * kNP + p] = h;
        }
        __syncthreads();                                   // (B4)

        // ---- finalize: thread (window w, column role [+ 14, ...]) merges the 13 partials of its column
        if (act) {
            for (int j = role; j < P.n_cols; j += kThreadsF / 16) {
                const int kind = P.col[j], ref = P.cref[j];
                const int q0 = w * kNP;
                double v;
                if (kind == MHB_S_ENTROPY) {
                    float h = 0.f;
#pragma unroll
                    for (int i = 0; i < kNP; ++i) h += sc.ph[q0 + i];
                    v = -static_cast<double>(h);
                } else if (kind == MHB_S_PEAK_FREQUENCY || kind == MHB_S_PEAK_BIN) {
                    float best = -1.f;
                    int arg = 0x7fffffff;
                    if (P.arg_lo[ref] <= 0 && P.arg_hi[ref] > 0) {
                        best = static_cast<float>(sc.dc[w]);
                        arg = 0;
                    }
                    const float* pb = sc.pbest + ref * kBW * kNP + q0;
                    const int* pa = sc.parg + ref * kBW * kNP + q0;
#pragma unroll
                    for (int i = 0; i < kNP; ++i) {
                        const float ob = pb[i];
                        const int oa = pa[i];
                        if (ob > best || (ob == best && oa < arg)) {
                            best = ob;
                            arg = oa;
                        }
                    }
                    if (arg == 0x7fffffff) v = CUDART_NAN;
                    else v = kind == MHB_S_PEAK_BIN ? static_cast<double>(arg) : static_cast<double>(arg) * P.bin_hz;
                } else {
                    float rest = 0.f;
#pragma unroll
                    for (int i = 0; i < kNP; ++i) rest += sc.ptot[q0 + i];
                    const double dc = sc.dc[w];
                    const double total = static_cast<double>(rest) + dc;
                    if (kind == MHB_S_TOTAL_POWER) {
                        v = total;
                    } else {
                        float acc = 0.f;
                        const float* ps = sc.psum + ref * kBW * kNP + q0;
#pragma unroll
                        for (int i = 0; i < kNP; ++i) acc += ps[i];
                        double bsum = static_cast<double>(acc);
                        if (P.sum_lo[ref] <= 0 && P.sum_hi[ref] > 0) bsum += dc;
                        v = kind == MHB_S_BAND_POWER ? bsum : bsum / total;
                    }
                }
                const int64_t o = series * P.o_series + (w0 + w) * P.o_window + j * P.o_col;
                if (P.out_f32) reinterpret_cast<float*>(P.out)[o] = static_cast<float>(v);
                else reinterpret_cast<double*>(P.out)[o] = v;
            }
        }
        // no barrier here: the next batch touches msum / buf / the other tile slot only, and its TMA refill of
        // THIS slot is issued after its barrier (B1), which every thread reaches after finishing this step
        slot ^= 1;
    }
}

size_t fast_smem_bytes() {
    return 128 + 2 * sizeof(float) * kTileElems + sizeof(C) * (kBW * kWSB + 25 * kNA + 10 * kNP) +
           sizeof(float) * (kMeanParts + 1) * kBW + sizeof(uint32_t) * (kMaxSum + kMaxArg) * 2 * 16 + 64;
}

}  // namespace

// Returns -100 when the geometry / column set has no fast plan (the caller then uses the generic kernel).
int32_t spectral_fast_try(const float* x, const mhb_windows* geom, int64_t nw, double bin_hz, const int32_t* cols,
                          const int32_t* lo, const int32_t* hi, int32_t n_cols, void* out, int32_t out_f32,
                          int64_t o_series, int64_t o_window, int64_t o_col, void* stream) {
    if (n_cols <= 0 || n_cols > kMaxColsF) return -100;
    if (geom->wsize != kW || geom->wstep != kS) return -100;
    FastPlan P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_cols; ++j) {
        const int kind = cols[j];
        P.col[j] = kind;
        P.cref[j] = 0;
        if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
            int r = 0;
            while (r < P.n_sum && !(P.sum_lo[r] == lo[j] && P.sum_hi[r] == hi[j])) ++r;
            if (r == P.n_sum) {
                if (P.n_sum == kMaxSum) return -100;
                P.sum_lo[r] = lo[j];
                P.sum_hi[r] = hi[j];
                ++P.n_sum;
            }
            P.cref[j] = r;
        } else if (kind == MHB_S_PEAK_FREQUENCY || kind == MHB_S_PEAK_BIN) {
            int r = 0;
            while (r < P.n_arg && !(P.arg_lo[r] == lo[j] && P.arg_hi[r] == hi[j])) ++r;
            if (r == P.n_arg) {
                if (P.n_arg == kMaxArg) return -100;
                P.arg_lo[r] = lo[j];
                P.arg_hi[r] = hi[j];
                ++P.n_arg;
            }
            P.cref[j] = r;
        }
    }
    P.x = x;
    P.series_stride = geom->series_stride;
    P.total_elems = (geom->n_series - 1) * geom->series_stride + geom->series_len;
    P.nw = nw;
    P.bin_hz = bin_hz;
    P.out = out;
    P.out_f32 = out_f32;
    P.o_series = o_series;
    P.o_window = o_window;
    P.o_col = o_col;
    P.n_cols = n_cols;
    P.use_tma = (reinterpret_cast<uintptr_t>(x) % 16 == 0 && geom->series_stride % 4 == 0) ? 1 : 0;
    P.batches_per_series = (nw + kBW - 1) / kBW;
    P.total_batches = P.batches_per_series * geom->n_series;
    if (P.total_batches >= (1LL << 31)) return -100;
    const size_t smem = fast_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(spectral_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e, "spectral_fast attr");
    int64_t ctas = static_cast<int64_t>(kNumSMs) * 3;
    if (ctas > P.total_batches) ctas = P.total_batches;
    spectral_fast_kernel<<<static_cast<unsigned>(ctas), kThreadsF, smem, static_cast<cudaStream_t>(stream)>>>(P);
    return cuda_status(cudaGetLastError(), "spectral_fast launch");
}

}  // namespace mhb
