// Kernel 2 (fast path, W = 1920 / S = 64: BASELINE config 4, 30 s PPG windows with a 1 s hop) -- per-window FFT + PSD
// reducers with thread-resident sub-transforms.  Same contract as window_spectral.cu (the generic path).
//
// Reference chain replaced: view (util/windows.py:20-33) -> mhealth.fft.fft (fft/_fft.py:18-29) -> |F|^2 ->
// hrv.power_band / relative_power_band (heart/hrv.py:173-198), density.peak_frequency
// (generic/frequency/density.py:18-32), information.entropy (generic/information.py:10-20).
//
// The real 1920-point DFT is an N = 960 point complex FFT, N = 8 x 12 x 10, three register-resident passes:
//   A1  thread (w, n2, b):  8-point DFT over z[120 a + 10 b + n2], a = 0..7, times w96^(b c)
//   A2  thread (w, n2, c):  12-point DFT over b -> Y[n2][k1 = c + 8 d], times w960^(n2 k1)   (in place: every thread
//                           loads its inputs, barrier, then stores)
//   B   thread (w, p):      the two 10-point DFTs over n2 that give Z[p + 96 k2] and Z[(96 - p) + 96 k2] -- exactly the
//                           (k, N - k) pairs of the real-input untangling -- so |X|^2, the per-thread totals, the
//                           entropy records and the masked band / arg-max partials are formed on REGISTER values
//   C   warp w:             merges the 49 partial records of window w with shuffles and stores the columns.
// A CTA (320 threads, two per SM) owns a batch of 6 consecutive windows of one series (2240 samples, one bulk TMA copy,
// double buffered); 97 % of a tile overlaps the next batch's, so the re-reads are L2 hits.  Pivot removal / exact bin 0 / 1/2 prescale as in
// spectral_fast.cu.  Deterministic: every reduction has a fixed order.
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "fft_core.cuh"
#include "psd_entropy.cuh"

namespace mhb {

namespace {

using C = Cx<float>;

constexpr int kW = 1920, kS = 64, kN = 960;
// Batch size / CTA shape: items per batch are 120 (A1), 80 (A2) and 49 (B) per window, so 6 windows on 320 threads give
// 3 / 2 / 1 rounds at 75 % / 75 % / 92 % occupancy of the thread slots, two CTAs per SM.  Measured on config 4 (16 series):
// 4 windows x 256 threads x 3 CTAs 8.83 ms, 5 x 320 x 2 8.32, 6 x 320 x 2 8.12, 7 x 384 x 2 8.51, 8 x 384 x 2 9.28.
#ifndef MHB_W1920_BW
#define MHB_W1920_BW 6
#define MHB_W1920_T 320
#define MHB_W1920_CTAS 2
#endif
constexpr int kBW = MHB_W1920_BW;        // windows per batch
constexpr int kT = MHB_W1920_T;
constexpr int kCtasPerSM = MHB_W1920_CTAS;
static_assert(kBW * 49 <= kT && kBW * 80 <= 2 * kT && kBW * 32 <= kT,
              "pass B is one round, pass A2 two rounds, pass C one warp per window");
constexpr int kNS2 = 99;                 // complex stride between the n2 blocks of a window (96 used)
constexpr int kWSTR = 10 * kNS2 + 1;     // complex stride between windows
constexpr int kCS = 122;                 // complex stride between the c planes of the A1 -> A2 exchange layout (8 x 122 <= kWSTR)
static_assert(8 * kCS <= kWSTR && kCS >= 120 && kCS % 16 == 10, "A1 -> A2 exchange layout");
constexpr int kNP = 49;                  // pass-B threads per window: p = 0..48
constexpr int kMaxCols = 32, kMaxSum = 4, kMaxArg = 2;
constexpr int kTile = (kBW - 1) * kS + kW;      // 2240 floats for 6 windows
constexpr int kMaskStride = 52;

struct Plan1920 {
    const float* x;
    int64_t series_stride, total_elems, nw;
    int64_t batches_per_series, total_batches;
    double bin_hz;
    void* out;
    int32_t out_f32;
    int64_t o_series, o_window, o_col;
    int32_t n_cols, n_sum, n_arg;
    int32_t col[kMaxCols], cref[kMaxCols];
    int32_t sum_lo[kMaxSum], sum_hi[kMaxSum], arg_lo[kMaxArg], arg_hi[kMaxArg];
    int32_t use_tma;
};

// bit k2 (low set: bin p + 96 k2) / bit 10 + k2 (high set: bin 960 - p - 96 k2) of the range mask of thread p
__device__ __forceinline__ uint32_t range_mask(int p, int lo, int hi) {
    uint32_t m = 0;
    for (int k2 = 0; k2 < 10; ++k2) {
        const int kl = p + 96 * k2, kh = kN - p - 96 * k2;
        if (kl >= lo && kl < hi && !(p == 0 && k2 == 0)) m |= 1u << k2;                  // bin 0: finalize step
        if (kh >= lo && kh < hi && !((p == 0 || p == 48) && k2 >= (p == 0 ? 1 : 0))) m |= 1u << (10 + k2);
    }
    return m;
}

__global__ void __launch_bounds__(kT, kCtasPerSM) spectral_w1920_kernel(const Plan1920 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    float* tiles = reinterpret_cast<float*>(smem_raw + 128);                // 2 x kTile
    C* buf = reinterpret_cast<C*>(tiles + 2 * kTile);                       // kBW x kWSTR
    C* tw96 = buf + kBW * kWSTR;                                            // [b * 8 + c]      w96^(b c)
    C* tw960 = tw96 + 96;                                                   // [k1 * 10 + n2]   w960^(n2 k1)
    C* twB = tw960 + 960;                                                   // [k2 * 49 + p]    w1920^(p + 96 k2)
    double* dcs = reinterpret_cast<double*>(twB + 10 * kNP);               // [kBW] exact bin 0
    float* piv = reinterpret_cast<float*>(dcs + kBW);                       // [kBW]
    float* rec_tot = piv + kBW;                                             // [kBW][kNP]
    float* rec_h = rec_tot + kBW * kNP;
    int* rec_e = reinterpret_cast<int*>(rec_h + kBW * kNP);
    float* psum = reinterpret_cast<float*>(rec_e + kBW * kNP);              // [kMaxSum][kBW][kNP]
    float* pbest = psum + kMaxSum * kBW * kNP;                              // [kMaxArg][kBW][kNP]
    int* parg = reinterpret_cast<int*>(pbest + kMaxArg * kBW * kNP);
    uint32_t* masks = reinterpret_cast<uint32_t*>(parg + kMaxArg * kBW * kNP);   // [kMaxSum + kMaxArg][kMaskStride]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 96; i += kT) {
        double sn, cs;
        sincospi(-2.0 * static_cast<double>((i >> 3) * (i & 7)) / 96.0, &sn, &cs);
        tw96[i] = {static_cast<float>(cs), static_cast<float>(sn)};
    }
    for (int i = tid; i < 960; i += kT) {
        const int k1 = i / 10, n2 = i - k1 * 10;
        double sn, cs;
        sincospi(-2.0 * static_cast<double>(k1 * n2) / 960.0, &sn, &cs);
        tw960[i] = {static_cast<float>(cs), static_cast<float>(sn)};
    }
    for (int i = tid; i < 10 * kNP; i += kT) {
        const int k2 = i / kNP, p = i - k2 * kNP;
        double sn, cs;
        sincospi(-2.0 * static_cast<double>(p + 96 * k2) / 1920.0, &sn, &cs);
        twB[i] = {static_cast<float>(cs), static_cast<float>(sn)};
    }
    for (int i = tid; i < (kMaxSum + kMaxArg) * kMaskStride; i += kT) {
        const int r = i / kMaskStride, p = i - r * kMaskStride;
        const bool is_sum = r < kMaxSum;
        const int ri = is_sum ? r : r - kMaxSum;
        const bool live = is_sum ? ri < P.n_sum : ri < P.n_arg;
        const int lo = is_sum ? P.sum_lo[ri] : P.arg_lo[ri], hi = is_sum ? P.sum_hi[ri] : P.arg_hi[ri];
        masks[i] = (live && p < kNP) ? range_mask(p, lo, hi) : 0u;
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t bps = static_cast<uint32_t>(P.batches_per_series);
    uint32_t series = static_cast<uint32_t>(blockIdx.x) / bps;
    uint32_t bi = static_cast<uint32_t>(blockIdx.x) - series * bps;
    const uint32_t step_s = gridDim.x / bps, step_b = gridDim.x - step_s * bps;
    const uint32_t n_series = static_cast<uint32_t>(P.total_batches / bps);
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
            fence_proxy_async();
            mbar_arrive_expect_tx(&full[slot], n_load * 4);
            bulk_g2s(tiles + slot * kTile, P.x + goff, n_load * 4, &full[slot]);
        }
    };

    int slot = 0;
    uint32_t parity0 = 0, parity1 = 0;
    if (tid == 0 && series < n_series) issue(series, bi, 0);

    for (; series < n_series; advance(series, bi)) {
        int64_t w0, goff;
        int nwin, n_valid;
        geom_of(series, bi, w0, nwin, goff, n_valid);
        float* tile = tiles + slot * kTile;
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
            for (int i = tid; i < n_valid; i += kT) tile[i] = P.x[goff + i];
        }
        // prefetch the next batch's tile into the other slot: its last readers (pass A1 of the previous batch) are
        // several barriers behind
        if (tid == 0) {
            uint32_t s2 = series, b2 = bi;
            advance(s2, b2);
            if (s2 < n_series) issue(s2, b2, slot ^ 1);
        }
        if (tma_ok(goff, n_load) == false) __syncthreads();       // cooperative copy visible

        // ---- pass A1: 8-point DFTs; item (w, b, n2), n2 fastest so that a warp reads consecutive float2
#pragma unroll 1
        for (int it = tid; it < kBW * 120; it += kT) {
            const int w = it / 120, r = it - w * 120;
            if (w >= nwin) continue;
            const int b = r / 10, n2 = r - b * 10;
            // pivot of the window: the mean of 4 samples from its two halves.  Every thread of the window reads the same
            // addresses (broadcasts), so all agree bit for bit; any value near the mean works (it only keeps a large DC
            // out of the float32 dynamic range), bin 0 is restored exactly in pass B.
            const float2* zw = reinterpret_cast<const float2*>(tile + w * kS);
            const float2 pa = zw[240], pb2 = zw[720];
            const float m = ((pa.x + pa.y) + (pb2.x + pb2.y)) * 0.25f;
            if (r == 0) piv[w] = m;
            const float mh = -0.5f * m;
            const float2* z = zw + 10 * b + n2;
            C a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 v = z[120 * q];
                a[q] = {fmaf(v.x, 0.5f, mh), fmaf(v.y, 0.5f, mh)};
            }
            rdft<8>(a);
            // A1 -> A2 exchange layout [c][b][n2] with a c stride of 122 = 10 (mod 16): consecutive lanes (n2 fastest, then
            // b) store consecutive complex words, and pass A2's lanes (n fastest, then c) read words congruent to their
            // lane-linear index modulo the 16 64-bit bank pairs -- both conflict free (the [n2][k1] layout A2 writes
            // and pass B reads is a different view of the same per-window region)
            C* dst = buf + w * kWSTR + 10 * b + n2;
            const C* tw = tw96 + b * 8;
            dst[0] = a[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) dst[c * kCS] = cmul(a[c], tw[c]);
        }
        __syncthreads();

        // ---- pass A2: 12-point DFTs, in place: load both of this thread's items, barrier, transform, store
        {
            C in0[12], in1[12];
            const int it0 = tid, it1 = tid + kT;
            const int wA = it0 / 80, rA = it0 - wA * 80, cA = rA / 10, nA = rA - cA * 10;
            const int wB = it1 / 80, rB = it1 - wB * 80, cB = rB / 10, nB = rB - cB * 10;
            const bool has0 = wA < nwin, has1 = it1 < kBW * 80 && wB < nwin;
            C* blk0 = buf + wA * kWSTR + nA * kNS2;
            C* blk1 = buf + wB * kWSTR + nB * kNS2;
            if (has0) {
                const C* src = buf + wA * kWSTR + cA * kCS + nA;
#pragma unroll
                for (int b = 0; b < 12; ++b) in0[b] = src[10 * b];
            }
            if (has1) {
                const C* src = buf + wB * kWSTR + cB * kCS + nB;
#pragma unroll
                for (int b = 0; b < 12; ++b) in1[b] = src[10 * b];
            }
            __syncthreads();
            if (has0) {
                rdft<12>(in0);
                const C* tw = tw960 + nA;
#pragma unroll
                for (int d = 0; d < 12; ++d) {
                    const int k1 = cA + 8 * d;
                    blk0[k1] = cmul(in0[d], tw[k1 * 10]);
                }
            }
            if (has1) {
                rdft<12>(in1);
                const C* tw = tw960 + nB;
#pragma unroll
                for (int d = 0; d < 12; ++d) {
                    const int k1 = cB + 8 * d;
                    blk1[k1] = cmul(in1[d], tw[k1 * 10]);
                }
            }
        }
        __syncthreads();

        // ---- pass B: thread (w, p): columns p and 96 - p -> (k, N - k) pairs -> |X|^2 in registers -> partials
        if (tid < kBW * kNP) {
            const int w = tid / kNP, p = tid - w * kNP;
            if (w < nwin) {
                const bool p0 = p == 0, pself = p == 48;
                const C* bw = buf + w * kWSTR;
                const int pb = p0 ? 0 : 96 - p;
                C A[10], B[10];
#pragma unroll
                for (int n2 = 0; n2 < 10; ++n2) {
                    A[n2] = bw[n2 * kNS2 + p];
                    B[n2] = bw[n2 * kNS2 + pb];
                }
                rdft<10>(A);                                   // A[k2] = Z[p + 96 k2] / 2
                rdft<10>(B);                                   // B[k2] = Z[96 - p + 96 k2] / 2
                if (p0) {                                      // partner of 96 k2 is 96 (10 - k2): rotate
                    const C b0 = B[0];
#pragma unroll
                    for (int m = 0; m < 9; ++m) B[m] = B[m + 1];
                    B[9] = b0;
                }
                float psd[20];                                 // [2 k2] low set (bin p + 96 k2), [2 k2 + 1] high set
#pragma unroll
                for (int k2 = 0; k2 < 10; ++k2) {
                    const C zk = A[k2], zn = B[9 - k2];
                    const C t2 = twB[k2 * kNP + p];
                    const C e = {zk.x + zn.x, zk.y - zn.y};
                    const C o = {zk.y + zn.y, zn.x - zk.x};
                    const C t = cmul(o, t2);
                    const float ar = e.x + t.x, ai = e.y + t.y, br = e.x - t.x, bi2 = e.y - t.y;
                    psd[2 * k2] = fmaf(ar, ar, ai * ai);
                    psd[2 * k2 + 1] = fmaf(br, br, bi2 * bi2);
                }
                if (p0) {
                    const double x0 = 2.0 * (static_cast<double>(A[0].x) + static_cast<double>(A[0].y)) +
                                      static_cast<double>(kW) * static_cast<double>(piv[w]);
                    dcs[w] = x0 * x0;                          // exact DC: FFT(x - m)[0] + W m
                    psd[0] = 0.f;                              // bin 0 is carried in float64
#pragma unroll
                    for (int k2 = 1; k2 < 10; ++k2) psd[2 * k2 + 1] = 0.f;      // duplicates of the low set
                }
                if (pself) {
#pragma unroll
                    for (int k2 = 0; k2 < 10; ++k2) psd[2 * k2 + 1] = 0.f;      // column 48 pairs with itself
                }
                float ta = 0.f, tb = 0.f;
#pragma unroll
                for (int i = 0; i < 20; i += 2) {
                    ta += psd[i];
                    tb += psd[i + 1];
                }
                const float tot = ta + tb;
                const int eb = (__float_as_int(tot) >> 23) & 0xff;
                const float scale = __int_as_float((254 - eb) << 23);
                float ha = 0.f, hb = 0.f;
#pragma unroll
                for (int i = 0; i < 20; i += 2) {
                    const float y1 = psd[i] * scale, y2 = psd[i + 1] * scale;
                    ha = fmaf(y1, log2_normal(fmaxf(y1, 1e-37f)), ha);
                    hb = fmaf(y2, log2_normal(fmaxf(y2, 1e-37f)), hb);
                }
                const int q = w * kNP + p;
                rec_tot[q] = tot;
                rec_h[q] = ha + hb;
                rec_e[q] = eb - 127;
#pragma unroll 1
                for (int r = 0; r < P.n_sum; ++r) {
                    const uint32_t m = masks[r * kMaskStride + p];
                    float acc = 0.f;
#pragma unroll
                    for (int k2 = 0; k2 < 10; ++k2) {
                        if (m & (1u << k2)) acc += psd[2 * k2];
                        if (m & (1u << (10 + k2))) acc += psd[2 * k2 + 1];
                    }
                    psum[r * kBW * kNP + q] = acc;
                }
                // first maximum over the thread's bins in ascending bin order: low k2 = i, then high k2 = 9 - i
#pragma unroll 1
                for (int r = 0; r < P.n_arg; ++r) {
                    const uint32_t m = masks[(kMaxSum + r) * kMaskStride + p];
                    float best = -1.f;
                    int arg = 0x7fffffff;
#pragma unroll
                    for (int i = 0; i < 10; ++i) {
                        if ((m & (1u << i)) && psd[2 * i] > best) {
                            best = psd[2 * i];
                            arg = p + 96 * i;
                        }
                        if ((m & (1u << (19 - i))) && psd[2 * (9 - i) + 1] > best) {
                            best = psd[2 * (9 - i) + 1];
                            arg = kN - p - 96 * (9 - i);
                        }
                    }
                    pbest[r * kBW * kNP + q] = best;
                    parg[r * kBW * kNP + q] = arg;
                }
            }
        }
        __syncthreads();

        // ---- pass C: warp w merges the 49 partial records of window w (lanes l and l + 32) and stores the columns
        if (warp < nwin) {
            const int w = warp;
            const int q0 = w * kNP;
            const bool two = lane + 32 < kNP;
            const double dc = dcs[w];
            const float tA = rec_tot[q0 + lane], tB = two ? rec_tot[q0 + lane + 32] : 0.f;
            const double rest = warp_sum(static_cast<double>(tA) + static_cast<double>(tB));
            const double total = rest + dc;
#pragma unroll 1
            for (int c = 0; c < P.n_cols; ++c) {
                const int kind = P.col[c], ref = P.cref[c];
                double v;
                bool redo = false;
                if (kind == MHB_S_TOTAL_POWER) {
                    v = total;
                } else if (kind == MHB_S_ENTROPY) {
                    // see spectral_fast.cu: E = exponent of the total, every term O(1)
                    const float tf = static_cast<float>(total);
                    const int E = ((__float_as_int(tf) >> 23) & 0xff) - 127;
                    const float down = __int_as_float((127 - E) << 23);
                    double acc = 0.0;
                    {
                        const int e = rec_e[q0 + lane], de = E - e;
                        const float rel = de < 60 ? __int_as_float((127 - de) << 23) : 0.f;
                        acc += static_cast<double>(tA * down) * static_cast<double>(de) -
                               static_cast<double>(rec_h[q0 + lane] * rel);
                    }
                    if (two) {
                        const int e = rec_e[q0 + lane + 32], de = E - e;
                        const float rel = de < 60 ? __int_as_float((127 - de) << 23) : 0.f;
                        acc += static_cast<double>(tB * down) * static_cast<double>(de) -
                               static_cast<double>(rec_h[q0 + lane + 32] * rel);
                    }
                    acc = warp_sum(acc);
                    const float f = tf * down;
                    const float hrest2 = static_cast<float>(rest * static_cast<double>(down) *
                                                            static_cast<double>(__log2f(f)) + acc) * __fdividef(1.0f, f);
                    const float inv_t = __fdividef(1.0f, tf);
                    const float pz = static_cast<float>(dc) * inv_t, qrest = static_cast<float>(rest) * inv_t;
                    const float h0 = pz > 0.f ? -pz * (qrest < 0.5f ? log1pf(-qrest) : __logf(pz)) : 0.f;
                    const float hf = fmaf(0.69314718055994530942f, hrest2, h0);
                    v = total > 0.0 ? static_cast<double>(hf) : CUDART_NAN;
                    redo = total > 0.0 && hf < kToneEntropy && pz < 0.5f;        // noiseless tone: psd_entropy.cuh
                } else if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
                    const float* ps = psum + ref * kBW * kNP + q0;
                    float acc = ps[lane] + (two ? ps[lane + 32] : 0.f);
                    double bsum = static_cast<double>(warp_sum(acc));
                    if (P.sum_lo[ref] <= 0 && P.sum_hi[ref] > 0) bsum += dc;
                    v = kind == MHB_S_BAND_POWER ? bsum : bsum / total;
                } else {
                    const float* pb = pbest + ref * kBW * kNP + q0;
                    const int* pa = parg + ref * kBW * kNP + q0;
                    float best = pb[lane];
                    int arg = pa[lane];
                    if (two) {
                        const float ob = pb[lane + 32];
                        const int oa = pa[lane + 32];
                        if (ob > best || (ob == best && oa < arg)) {
                            best = ob;
                            arg = oa;
                        }
                    }
                    if (lane == 0 && P.arg_lo[ref] <= 0 && P.arg_hi[ref] > 0) {
                        const float d = static_cast<float>(dc);
                        if (d >= best) {
                            best = d;
                            arg = 0;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
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
                if (lane == 0) {
                    const int64_t o = static_cast<int64_t>(series) * P.o_series + (w0 + w) * P.o_window + c * P.o_col;
                    if (redo) {
                        if (P.out_f32) reinterpret_cast<uint32_t*>(P.out)[o] = kRedoMarkF32;
                        else reinterpret_cast<unsigned long long*>(P.out)[o] = kRedoMarkF64;
                    } else if (P.out_f32) {
                        reinterpret_cast<float*>(P.out)[o] = static_cast<float>(v);
                    } else {
                        reinterpret_cast<double*>(P.out)[o] = v;
                    }
                }
            }
        }
        // no barrier here: the next batch's pivots / pass A1 write piv and buf (pass B's reads are behind the barrier
        // above); the records are rewritten by its pass B, three barriers from here
        slot ^= 1;
    }
}

size_t smem_bytes_1920() {
    return 128 + 2 * sizeof(float) * kTile + sizeof(C) * (kBW * kWSTR + 96 + 960 + 10 * kNP) + sizeof(double) * kBW +
           sizeof(float) * kBW + sizeof(float) * (3 + kMaxSum + 2 * kMaxArg) * kBW * kNP +
           sizeof(uint32_t) * (kMaxSum + kMaxArg) * kMaskStride + 64;
}

}  // namespace

// Returns -100 when the geometry / column set has no plan here (the caller then uses the generic kernel).
int32_t spectral_w1920_try(const float* x, const mhb_windows* geom, int64_t nw, double bin_hz, const int32_t* cols,
                           const int32_t* lo, const int32_t* hi, int32_t n_cols, void* out, int32_t out_f32,
                           int64_t o_series, int64_t o_window, int64_t o_col, void* stream) {
    if (n_cols <= 0 || n_cols > kMaxCols) return -100;
    if (geom->wsize != kW || geom->wstep != kS) return -100;
    Plan1920 P;
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < n_cols; ++j) {
        const int kind = cols[j];
        P.col[j] = kind;
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
    const size_t smem = smem_bytes_1920();
    cudaError_t e = cudaFuncSetAttribute(spectral_w1920_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e, "spectral_w1920 attr");
    int64_t ctas = static_cast<int64_t>(kNumSMs) * kCtasPerSM;
    if (ctas > P.total_batches) ctas = P.total_batches;
    spectral_w1920_kernel<<<static_cast<unsigned>(ctas), kT, smem, static_cast<cudaStream_t>(stream)>>>(P);
    return cuda_status(cudaGetLastError(), "spectral_w1920 launch");
}

}  // namespace mhb
