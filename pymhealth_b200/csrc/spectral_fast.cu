// Kernel 2 (fast path, W = 500 / S = 250) -- per-window FFT + PSD reducers with thread-resident sub-transforms.
//
// Same contract as window_spectral.cu (the generic path for arbitrary window lengths).  Reference chain
// replaced: view (util/windows.py:20-33) -> mhealth.fft.fft (fft/_fft.py:18-29) -> |F|^2 ->
// hrv.power_band / relative_power_band (heart/hrv.py:173-198), density.peak_frequency
// (generic/frequency/density.py:18-32), information.entropy (generic/information.py:10-20).
//
// A CTA (8 warps) owns a batch of BW = 16 consecutive windows of one series; their samples arrive ONCE by a
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
#include "psd_entropy.cuh"

namespace mhb {

namespace {

using C = Cx<float>;

constexpr int kW = 500, kS = 250, kN = 250;
constexpr int kBW = 16;                  // windows per batch
constexpr int kThreadsF = 256;           // 16 half-warps: 10 pass-A roles / 13 pass-B roles / 3 reducer warps
constexpr int kNA = 10;                  // pass-A threads per window (n2)
constexpr int kNP = 13;                  // pass-B threads per window (p = 0..12)
constexpr int kWSB = 251;                // complex stride between windows in buf (odd: conflict free)
constexpr int kRowStride = 254;          // float stride between PSD rows: 2 x odd, so the 16 windows of the two roles of a warp (bins k and k + 1) cover all 32 banks
constexpr int kMaxColsF = 32;
constexpr int kTileElems = ((kBW - 1) * kS + kW + 3 + 4) & ~3;     // 4256 floats
static_assert(2 * kBW + kBW * kRowStride <= kTileElems, "PSD rows + DC cells must fit a consumed tile slot");

struct FastPlan {
    const float* x;
    int64_t series_stride, total_elems, nw;
    int64_t batches_per_series, total_batches;
    double bin_hz;
    void* out;
    int32_t out_f32;
    int64_t o_series, o_window, o_col;
    int32_t n_cols;
    int32_t col[kMaxColsF];              // MHB_S_* kind
    int32_t lo[kMaxColsF], hi[kMaxColsF];   // bin range [lo, hi) of the column
    int32_t rn[3];                       // columns handled by each of the three reducer warps ...
    int32_t rcol[3][kMaxColsF];          // ... and which (cost-balanced on the host)
    int32_t use_tma;
};

// ---- small in-register DFTs (forward, e^{-2 pi i / R}): the packed butterflies of fft_core.cuh
__device__ __forceinline__ void dft2(C* a) { rdft2(a); }
__device__ __forceinline__ void dft5(C* a) { rdft5(a); }

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
// Same transform, but every output is handed to `emit(t, value)` as soon as its last butterfly is done, so the caller's
// twiddle loads / stores interleave with the remaining butterflies instead of queueing behind them.
template <int R, int RA, int RB, typename Emit>
__device__ __forceinline__ void dft_composite_emit(C* a, Emit emit) {
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
        for (int t1 = 0; t1 < RA; ++t1) emit(t2 + RB * t1, g[t1]);
    }
}
__device__ __forceinline__ void dft10(C* a) { dft_composite<10, 2, 5>(a); }
__device__ __forceinline__ void dft25(C* a) { dft_composite<25, 5, 5>(a); }

#ifdef MHB_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8][8];   // [phase][warp]
__device__ unsigned long long g_reduce_cycles[16];    // reducer sub-phases (summed over the reducer warps)
#define MHB_RTICK(i)                                                                                         \
    do {                                                                                                     \
        const long long now_ = clock64();                                                                    \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_reduce_cycles[i], static_cast<unsigned long long>(now_ - rl)); \
        rl = clock64();                                                                                      \
    } while (0)
#else
#define MHB_RTICK(i) do {} while (0)
#endif

// Deferred PSD reducers of one finished batch.  Three warps (the ones with no pass-A work) share the COLUMNS: warp
// rw takes columns c = rw, rw + 3, ...; inside a warp a pair of lanes owns a window.  Totals and the entropy come
// from the 13 per-thread records pass B left behind (sum, sum y log2 y, exponent); band sums and arg-max read only
// the bins of their range from the window's PSD row in shared memory.  Short instruction streams on purpose: on a
// saturated SM a warp issues about once every five cycles, so the longest per-warp stream of a phase sets its length.
struct Records {
    const float* ptot;   // [BW][NP]  sum of the thread's bins (bin 0 excluded)
    const float* ph;     // [BW][NP]  sum y log2 y over the thread's bins, y = psd 2^-pe
    const int* pe;       // [BW][NP]  binary exponent of the thread's ptot
};

__device__ __forceinline__ float pair_sum(float v) { return v + __shfl_xor_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double pair_sum(double v) { return v + __shfl_xor_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ void reduce_rows(const FastPlan& P, const float* __restrict__ slot_base, const Records R,
                                            const int4* __restrict__ desc, int n_mine, int lane, uint32_t series,
                                            int64_t w0, int nwin) {
#ifdef MHB_PHASE_TIMING
    long long rl = clock64();
#endif
    const int w = lane >> 1, j = lane & 1;
    const bool act = w < nwin;
    const int wc = act ? w : 0;                       // idle pairs alias window 0: they never store
    const double* dcs = reinterpret_cast<const double*>(slot_base);
    const float* prow = slot_base + 2 * kBW + wc * kRowStride;      // prow[k] = |X[k]|^2, k = 1..250
    const double dc = dcs[wc];
    const int q0 = wc * kNP;
    // lane j folds records j, j + 2, ..., in a fixed order; the pair sum gives both lanes the same bits
    double rest = 0.0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int idx = j + 2 * i;
        if (idx < kNP) rest += static_cast<double>(R.ptot[q0 + idx]);
    }
    rest = pair_sum(rest);                            // all bins but 0
    const double total = rest + dc;
    MHB_RTICK(0);
#pragma unroll 1
    for (int ci = 0; ci < n_mine; ++ci) {
        const int4 dsc = desc[ci];                    // (column, kind, lo, hi) from shared memory
        const int c = dsc.x, kind = dsc.y, lo = dsc.z, hi = dsc.w;
        double v;
        if (kind == MHB_S_TOTAL_POWER) {
            v = total;
        } else if (kind == MHB_S_ENTROPY) {
            // information.py:10-20: H = -p0 ln p0 - sum_{k>=1} p_k ln p_k.  With E = binary exponent of the total and
            // f = total 2^-E in [1, 2):
            //   -(sum_{k>=1} p_k log2 p_k) f = (rest 2^-E) log2 f + sum_t [ tot_t 2^-E (E - e_t) - S_t 2^(e_t - E) ]
            // every term is O(1): a float log2 of f (absolute error 2^-22) and a float64 sum of float records suffice
            const float tf = static_cast<float>(total);
            const int E = ((__float_as_int(tf) >> 23) & 0xff) - 127;
            const float down = __int_as_float((127 - E) << 23);              // 2^-E
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int idx = j + 2 * i;
                if (idx < kNP) {
                    const int e = R.pe[q0 + idx];
                    const int de = E - e;
                    const float u = R.ptot[q0 + idx] * down;
                    const float rel = de < 60 ? __int_as_float((127 - de) << 23) : 0.f;      // 2^(e - E)
                    acc += static_cast<double>(u) * static_cast<double>(de) - static_cast<double>(R.ph[q0 + idx] * rel);
                }
            }
            acc = pair_sum(acc);
            const float f = tf * down;
            const float hrest2 = static_cast<float>(rest * static_cast<double>(down) * static_cast<double>(__log2f(f)) + acc) *
                                 __fdividef(1.0f, f);
            const float inv_t = __fdividef(1.0f, tf);
            const float p0 = static_cast<float>(dc) * inv_t, qrest = static_cast<float>(rest) * inv_t;
            const float h0 = p0 > 0.f ? -p0 * (qrest < 0.5f ? log1pf(-qrest) : __logf(p0)) : 0.f;   // log1p near p0 = 1
            const float hf = fmaf(0.69314718055994530942f, hrest2, h0);
            v = total > 0.0 ? static_cast<double>(hf) : CUDART_NAN;
            if (total > 0.0 && hf < kToneEntropy && p0 < 0.5f) {      // noiseless tone: psd_entropy.cuh (both lanes of
                const unsigned pm = 3u << (lane & 30);                // the pair hold the same bits and branch together)
                v = entropy_dominant_bin(
                    prow, kN + 1, dc, total, j, 2,
                    [pm](float& b, int& a) {
                        const float ob = __shfl_xor_sync(pm, b, 1);
                        const int oa = __shfl_xor_sync(pm, a, 1);
                        if (ob > b || (ob == b && oa < a)) {
                            b = ob;
                            a = oa;
                        }
                    },
                    [pm](double s) { return s + __shfl_xor_sync(pm, s, 1); });
            }
        } else if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
            float a0 = 0.f, a1 = 0.f;
            int k = (lo > 1 ? lo : 1) + j;
            for (; k + 2 < hi; k += 4) {
                a0 += prow[k];
                a1 += prow[k + 2];
            }
            if (k < hi) a0 += prow[k];
            double bsum = static_cast<double>(pair_sum(a0 + a1));
            if (lo <= 0 && hi > 0) bsum += dc;
            v = kind == MHB_S_BAND_POWER ? bsum : bsum / total;
        } else {                                      // peak frequency / bin: first maximum in [lo, hi)
            // four independent running maxima over interleaved bins (the compare-select chain is the latency), merged
            // with "lower bin wins a tie" so the result is the first maximum
            float b0 = -1.f, b1 = -1.f, b2 = -1.f, b3 = -1.f;
            int g0 = 0x7fffffff, g1 = 0x7fffffff, g2 = 0x7fffffff, g3 = 0x7fffffff;
            int k = (lo > 1 ? lo : 1) + j;
            for (; k + 6 < hi; k += 8) {
                const float v0 = prow[k], v1 = prow[k + 2], v2 = prow[k + 4], v3 = prow[k + 6];
                if (v0 > b0) { b0 = v0; g0 = k; }
                if (v1 > b1) { b1 = v1; g1 = k + 2; }
                if (v2 > b2) { b2 = v2; g2 = k + 4; }
                if (v3 > b3) { b3 = v3; g3 = k + 6; }
            }
            for (; k < hi; k += 2) {
                const float v0 = prow[k];
                if (v0 > b0) { b0 = v0; g0 = k; }
            }
            if (b1 > b0 || (b1 == b0 && g1 < g0)) { b0 = b1; g0 = g1; }
            if (b3 > b2 || (b3 == b2 && g3 < g2)) { b2 = b3; g2 = g3; }
            if (b2 > b0 || (b2 == b0 && g2 < g0)) { b0 = b2; g0 = g2; }
            float best = b0;
            int arg = g0;
            if (j == 0 && lo <= 0 && hi > 0) {        // bin 0 competes too; ties go to the lower bin
                const float d = static_cast<float>(dc);
                if (d >= best) {
                    best = d;
                    arg = 0;
                }
            }
            const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, 1);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
            if (arg == 0x7fffffff) v = CUDART_NAN;
            else v = kind == MHB_S_PEAK_BIN ? static_cast<double>(arg) : static_cast<double>(arg) * P.bin_hz;
        }
        MHB_RTICK(1 + c);
        if (act && j == 0) {
            const int64_t o = static_cast<int64_t>(series) * P.o_series + (w0 + w) * P.o_window + c * P.o_col;
            if (P.out_f32) reinterpret_cast<float*>(P.out)[o] = static_cast<float>(v);
            else reinterpret_cast<double*>(P.out)[o] = v;
        }
    }
}

// ---- pass B of one (window, p): 10-point DFTs -> untangle -> |X|^2 -> PSD row + records.
// Thread (w, p) transforms columns p and 25 - p of buf: A[k2] = Z[p + 25 k2] / 2, B[k2] = Z[25 - p + 25 k2] / 2, the
// partner of bin k = p + 25 k2 being N - k = (25 - p) + 25 (9 - k2).  P0 (p = 0): the partner of 25 k2 is 25 (10 - k2), a
// value of the SAME column, so one transform serves both sets; its k2 = 0 pair yields bins 0 and N, and the high-set
// values k2 >= 1 duplicate the low set (they are kept out of the records).
template <bool P0>
__device__ __forceinline__ void pass_b(const C* __restrict__ bw, const C* __restrict__ twB, float* __restrict__ prow, int p,
                                       int w, double* __restrict__ dcs, float pivot, const Records R) {
    C A[10], B[10];
#pragma unroll
    for (int n2 = 0; n2 < 10; ++n2) {
        A[n2] = bw[n2 * 25 + p];
        if (!P0) B[n2] = bw[n2 * 25 + 25 - p];
    }
    dft10(A);
    if (!P0) dft10(B);
    float2 psd[10];                                // this thread's bins: .x low set (bin p + 25 k2), .y high set
    float* lo_ptr = prow + p;                      // bin p + 25 k2
    float* hi_ptr = prow + kN - p;                 // bin 250 - p - 25 k2
#pragma unroll
    for (int k2 = 0; k2 < 10; ++k2) {
        const C zk = A[k2], zn = P0 ? A[(10 - k2) % 10] : B[9 - k2];       // (k, N - k), k = p + 25 k2
        const C t2 = twB[k2 * kNP + p];
        // e = zk + conj(zn), o = -i (zk - conj(zn)), X[k] = e + t2 o, conj(X[N - k]) = e - t2 o: eight packed
        // instructions per pair of bins
        const float2 e = __fadd2_rn(f2(zk), make_float2(zn.x, -zn.y));
        const C o = cx(__fadd2_rn(make_float2(zk.y, -zk.x), make_float2(zn.y, zn.x)));
        const C t = cmul(o, t2);
        // (Re X[k], Re X[N - k]) = e.x +- t.x as ONE packed FMA (1, -1) * t.x + e.x with both scalars as broadcast operands
        // (same roundings as the two additions); a packed add of (e.x, e.x) and (t.x, -t.x) needs the second pair built
        // with two MOVs first
        const float2 pm = make_float2(1.f, -1.f);
        const float2 re = __ffma2_rn(pm, make_float2(t.x, t.x), make_float2(e.x, e.x));
        const float2 im = __ffma2_rn(pm, make_float2(t.y, t.y), make_float2(e.y, e.y));
        psd[k2] = __ffma2_rn(re, re, __fmul2_rn(im, im));
        if (!P0 || k2 == 0) hi_ptr[-25 * k2] = psd[k2].y;      // P0: bin N; the other high-set bins belong to the low set
        if (!P0 || k2 > 0) lo_ptr[25 * k2] = psd[k2].x;
    }
    if (P0) {
        // exact bin 0: FFT(x - m)[0] + W m in float64; the records must count every bin once
        const double x0 = 2.0 * (static_cast<double>(A[0].x) + static_cast<double>(A[0].y)) +
                          static_cast<double>(kW) * static_cast<double>(pivot);
        dcs[w] = x0 * x0;
        psd[0].x = 0.f;
#pragma unroll
        for (int k2 = 1; k2 < 10; ++k2) psd[k2].y = 0.f;
    }
    // records for the deferred reducers: total of the thread's bins, and the entropy partial in ONE pass: with
    // y = psd 2^-e (e = binary exponent of this thread's total -- an exact scaling that keeps |log2 y| small for
    // the bins that matter)  sum psd log2 psd = 2^e sum y log2 y + e tot
    float2 tt = psd[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) tt = __fadd2_rn(tt, psd[i]);
    const float tot = tt.x + tt.y;
    const int eb = (__float_as_int(tot) >> 23) & 0xff;
    const float scale = __int_as_float((254 - eb) << 23);               // 2^-(eb - 127)
    float2 hh = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        // + 1e-37: y log2 y -> 0 for an empty bin without a compare (the offset is 2^-123 against y = O(1) for every bin
        // that matters, far below float32 resolution)
        const float2 y = __ffma2_rn(psd[i], make_float2(scale, scale), make_float2(1e-37f, 1e-37f));
        hh = __ffma2_rn(y, make_float2(log2_normal(y.x), log2_normal(y.y)), hh);
    }
    const int q = w * kNP + p;
    const_cast<float*>(R.ptot)[q] = tot;
    const_cast<float*>(R.ph)[q] = hh.x + hh.y;
    const_cast<int*>(R.pe)[q] = eb - 127;
}

#ifdef MHB_PHASE_TIMING
#define MHB_TICK(ph)                                                             \
    do {                                                                         \
        const long long now_ = clock64();                                        \
        if ((threadIdx.x & 31) == 0) tacc[ph] += now_ - tlast;                   \
        tlast = clock64();                                                       \
    } while (0)
#else
#define MHB_TICK(ph) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreadsF, 3) spectral_fast_kernel(const FastPlan P) {
#ifdef MHB_PHASE_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // 2 barriers
    float* tiles = reinterpret_cast<float*>(smem_raw + 128);                // 2 x kTileElems
    C* buf = reinterpret_cast<C*>(tiles + 2 * kTileElems);                  // BW x kWSB complex
    C* twA = buf + kBW * kWSB;                                              // [25][10]  w250^(n2 k1)
    C* twB = twA + 25 * kNA;                                                // [10][13]  w500^(p + 25 k2)
    float* piv = reinterpret_cast<float*>(twB + 10 * kNP);                  // [BW] pivot (mean estimate) per window
    float* rec_tot = piv + kBW;                                             // [BW][NP] pass-B records (see Records)
    float* rec_h = rec_tot + kBW * kNP;
    int* rec_e = reinterpret_cast<int*>(rec_h + kBW * kNP);
    const Records R = {rec_tot, rec_h, rec_e};
    int4* desc = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(rec_e + kBW * kNP) + 15) & ~uintptr_t(15));   // [3][kMaxColsF]
    for (int i = threadIdx.x; i < 3 * kMaxColsF; i += kThreadsF) {
        const int rwi = i / kMaxColsF, ci = i - rwi * kMaxColsF;
        if (ci < P.rn[rwi]) {
            const int c = P.rcol[rwi][ci];
            desc[i] = make_int4(c, P.col[c], P.lo[c], P.hi[c]);
        }
    }
    const int tid = threadIdx.x;
    const int w = tid & 15;              // window of the batch this thread transforms
    const int role = tid >> 4;           // half-warp index 0..13

    for (int i = tid; i < 25 * kNA; i += kThreadsF) twA[i] = {kTw250[i].x, kTw250[i].y};
    for (int i = tid; i < 10 * kNP; i += kThreadsF) twB[i] = {kTw500[i].x, kTw500[i].y};
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // batch b = (series, bi): advanced incrementally (no division in the loop)
    const uint32_t bps = static_cast<uint32_t>(P.batches_per_series);          // host guarantees < 2^31 batches
    uint32_t series = static_cast<uint32_t>(blockIdx.x) / bps;
    uint32_t bi = static_cast<uint32_t>(blockIdx.x) - series * bps;
    const uint32_t step_s = gridDim.x / bps, step_b = gridDim.x - step_s * bps;
    auto geom_of = [&](uint32_t sr, uint32_t bix, int64_t& w0, int& nwin, int64_t& goff, int& n_valid) {
        w0 = static_cast<int64_t>(bix) * kBW;
        const int64_t left = P.nw - w0;
        nwin = left < kBW ? static_cast<int>(left) : kBW;
        goff = static_cast<int64_t>(sr) * P.series_stride + w0 * kS;
        n_valid = (nwin - 1) * kS + kW;
    };
    auto advance = [&](uint32_t& sr, uint32_t& bix) {
        sr += step_s;
        bix += step_b;
        if (bix >= bps) {
            bix -= bps;
            ++sr;
        }
    };
    auto tma_ok = [&](int64_t goff, int n_load) {
        return P.use_tma && (goff % 4 == 0) && (goff + n_load <= P.total_elems);
    };
    auto issue = [&](uint32_t sr, uint32_t bix, int slot) {              // thread 0 only
        int64_t w0, goff;
        int nwin, n_valid;
        geom_of(sr, bix, w0, nwin, goff, n_valid);
        const int n_load = (n_valid + 3) & ~3;
        if (tma_ok(goff, n_load)) {
            fence_proxy_async();                          // earlier generic-proxy accesses of the slot are done
            mbar_arrive_expect_tx(&full[slot], n_load * 4);
            bulk_g2s(tiles + slot * kTileElems, P.x + goff, n_load * 4, &full[slot]);
        }
    };

    int slot = 0;
    uint32_t parity0 = 0, parity1 = 0;
    const uint32_t n_series = static_cast<uint32_t>(P.total_batches / bps);
    if (tid == 0) {
        if (series < n_series) issue(series, bi, 0);
        uint32_t s2 = series, b2 = bi;
        advance(s2, b2);
        if (s2 < n_series) issue(s2, b2, 1);
    }
    bool first = true;
    uint32_t prev_series = 0;            // the batch whose PSD rows wait in the other slot
    int64_t prev_w0 = 0;
    int prev_nwin = 0;

    // one iteration per batch plus a last one that only reduces the final batch's rows (`live` false)
    for (;; advance(series, bi)) {
        const bool live = series < n_series;
        if (!live && first) break;                         // this CTA had no batch at all
        int64_t w0 = 0, goff = 0;
        int nwin = 0, n_valid = 0;
        float* tile = tiles + slot * kTileElems;
        if (live) {
            geom_of(series, bi, w0, nwin, goff, n_valid);
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
                // unaligned base or the last few samples of the buffer: guarded cooperative copy.  The slot held the
                // PSD rows of the batch before the previous one; they were reduced before the barriers of the last
                // iteration.
                for (int i = tid; i < n_valid; i += kThreadsF) tile[i] = P.x[goff + i];
                __syncthreads();
            }
        }
        const bool act = w < nwin;
        MHB_TICK(0);                                       // tile wait

        if (role < kNA) {
            // ---- pass A (warps 0-4): 25-point DFTs of the stride-10 subsequences, inter-pass twiddle, exchange
            if (act && live) {
                const int n2 = role;
                // pivot of the transform: a 4-sample estimate of the window mean (any value near the mean works:
                // bin 0 is restored exactly in pass B; the pivot only keeps a large DC out of the float32 dynamic
                // range).  Every thread of the window reads the same 4 samples, so they agree bit for bit.
                const float* xw = tile + w * kS;
                const float m = ((xw[62] + xw[187]) + (xw[312] + xw[437])) * 0.25f;
                const float mh = -0.5f * m;
                const float2* z = reinterpret_cast<const float2*>(xw) + n2;
                C a[25];
#pragma unroll
                for (int n1 = 0; n1 < 25; ++n1) {
                    a[n1] = cx(__ffma2_rn(z[kNA * n1], make_float2(0.5f, 0.5f), make_float2(mh, mh)));
                }
                dft25(a);
                C* dst = buf + w * kWSB + n2 * 25;
                auto st = [](C* q, C v) {      // plain STS.64 from the register pair the product was formed in
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(smem_u32(q)), "f"(v.x), "f"(v.y) : "memory");
                };
                st(dst, a[0]);
                // w250^(n2 k1); the n2 = 0 threads multiply by the table's exact (1, 0) (they share their warps with
                // n2 = 1: a separate copy loop would only add a divergent path)
                const C* tw = twA + n2;
#pragma unroll
                for (int k0 = 1; k0 < 25; k0 += 6) {       // six broadcast loads in flight, then their products
                    C t[6];
#pragma unroll
                    for (int u = 0; u < 6; ++u) t[u] = tw[(k0 + u) * kNA];
#pragma unroll
                    for (int u = 0; u < 6; ++u) st(dst + k0 + u, cmul(a[k0 + u], t[u]));
                }
                if (n2 == 0) piv[w] = m;
            }
        } else if (!first) {
            // ---- warps 5-7: PSD reducers of the PREVIOUS batch, whose rows sit in the other (consumed) tile slot
            reduce_rows(P, tiles + (slot ^ 1) * kTileElems, R, desc + ((tid >> 5) - 5) * kMaxColsF, P.rn[(tid >> 5) - 5], tid & 31,
                        prev_series, prev_w0, prev_nwin);
        }
        if (!live) break;
        MHB_TICK(1);                                       // pass A / deferred reducers
        __syncthreads();                                   // (B1) pass A done, previous batch fully reduced
        MHB_TICK(2);                                       // wait at B1
        // refill the slot the previous batch used (its PSD rows are dead now)
        if (tid == 0 && !first) {
            uint32_t s2 = series, b2 = bi;
            advance(s2, b2);
            if (s2 < n_series) issue(s2, b2, slot ^ 1);
        }
        first = false;

        // ---- pass B: 10-point DFTs -> untangle -> |X|^2 -> PSD row in the consumed tile slot.
        // Thread (w, p) transforms columns p and 25 - p of buf: A[k2] = Z[p + 25 k2] / 2, B[k2] = Z[25 - p + 25 k2] / 2,
        // the partner of bin k = p + 25 k2 being N - k = (25 - p) + 25 (9 - k2).  p = 0 runs the SAME code: it loads
        // column 0 twice and rotates, B[m] = A[(m + 1) mod 10], because the partner of 25 k2 is 25 (10 - k2); its
        // k2 = 0 pair then yields bins 0 and N, and its high-set values k2 >= 1 are duplicates of its low set.
        // Pass-B roles are handed out from the LAST half-warp down (role_b = 15 - role): pass A keeps warps 0-4 busy, i.e.
        // two warps of sub-partition 0 (warp w runs on sub-partition w mod 4) and one of each other; counting from the
        // top gives sub-partition 0 one pass-B warp (warp 4) and the others two, so the FMA-pipe cycles of a batch are
        // spread 1830 / 1480 / 1710 / 1710 over the four schedulers instead of 2360 / 1710 / 1710 / 1180.
        const int role_b = 15 - role;
        if (role_b < kNP && act) {
            float* prow = tile + 2 * kBW + w * kRowStride;
            if (role_b < kNP - 1) pass_b<false>(buf + w * kWSB, twB, prow, role_b + 1, w, nullptr, 0.f, R);
            else pass_b<true>(buf + w * kWSB, twB, prow, 0, w, reinterpret_cast<double*>(tile), piv[w], R);   // p = 0: upper half of warp 1
        }
        MHB_TICK(3);                                       // pass B
        __syncthreads();                                   // (B2) PSD rows complete; buf free for the next pass A
        MHB_TICK(4);                                       // wait at B2
        prev_series = series;
        prev_w0 = w0;
        prev_nwin = nwin;
        slot ^= 1;
    }
#ifdef MHB_PHASE_TIMING
    if ((threadIdx.x & 31) == 0)
        for (int ph = 0; ph < 8; ++ph) atomicAdd(&g_phase_cycles[ph][threadIdx.x >> 5], static_cast<unsigned long long>(tacc[ph]));
#endif
}

size_t fast_smem_bytes() {
    return 128 + 2 * sizeof(float) * kTileElems + sizeof(C) * (kBW * kWSB + 25 * kNA + 10 * kNP) + sizeof(float) * (kBW + 3 * kBW * kNP) + 16 * 3 * kMaxColsF + 64 + 16;
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
        P.col[j] = cols[j];
        P.lo[j] = lo[j];
        P.hi[j] = hi[j];
    }
    {   // longest-processing-time-first assignment of the columns to the three reducer warps
        int cost[kMaxColsF], order[kMaxColsF], load[3] = {0, 0, 0};
        for (int j = 0; j < n_cols; ++j) {
            const int span = hi[j] > lo[j] ? hi[j] - lo[j] : 0;
            const int kind = cols[j];
            cost[j] = kind == MHB_S_TOTAL_POWER ? 40 : kind == MHB_S_ENTROPY ? 160
                      : (kind == MHB_S_BAND_POWER ? 80 + span : kind == MHB_S_REL_BAND_POWER ? 120 + span : 100 + 3 * span);
            order[j] = j;
        }
        for (int a = 1; a < n_cols; ++a)
            for (int b2 = a; b2 > 0 && cost[order[b2]] > cost[order[b2 - 1]]; --b2) {
                const int tmp = order[b2];
                order[b2] = order[b2 - 1];
                order[b2 - 1] = tmp;
            }
        for (int a = 0; a < n_cols; ++a) {
            int best = 0;
            for (int r = 1; r < 3; ++r)
                if (load[r] < load[best]) best = r;
            P.rcol[best][P.rn[best]++] = order[a];
            load[best] += cost[order[a]];
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
