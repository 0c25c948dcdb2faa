// In-shared-memory mixed-radix Stockham FFT used by the spectral kernels (no cuFFT).
//
// The reference's FFT provider is FFTW through a 6-line CFFI shim, one plan per call
// (src/mhealth/fft/_fftw_binder.py:8-20), or numpy's pocketfft; here a thread group (a warp or a
// CTA) transforms one row that lives in shared memory: radix 4 / 2 / 3 / 5 butterflies plus a
// generic O(p^2) butterfly for other primes <= 31, autosort (no bit reversal), twiddles from a
// shared table that the CTA fills once with sincospi in float64.
#pragma once
#include "common.cuh"
#include "fft_consts.cuh"

namespace mhb {

constexpr int kMaxRadices = 24;
constexpr int kMaxPrime = 31;

struct FftPlan {
    int32_t n;                    // transform length
    int32_t n_radices;
    int32_t radix[kMaxRadices];
};

// host: factor n into supported radices; false if a prime factor > kMaxPrime remains
// `composite`: also use the register-resident composite butterflies 16 / 12 / 10 / 8 / 6 (float32 transforms): fewer
// passes, i.e. fewer round trips through shared memory (960 = 16 x 12 x 5 instead of 4 x 4 x 4 x 3 x 5)
static inline bool fft_plan(int32_t n, FftPlan* p, bool composite = false) {
    p->n = n;
    p->n_radices = 0;
    int32_t r = n;
    auto push = [&](int f) { p->radix[p->n_radices++] = f; };
    if (composite) {
        const int big[5] = {16, 12, 10, 8, 6};
        for (int i = 0; i < 5; ++i)
            while (r % big[i] == 0 && r / big[i] != 2 && r / big[i] != 3) { push(big[i]); r /= big[i]; }
    }
    while (r % 4 == 0) { push(4); r /= 4; }
    while (r % 2 == 0) { push(2); r /= 2; }
    for (int f = 3; f <= kMaxPrime; f += 2)
        while (r % f == 0) {
            if (p->n_radices >= kMaxRadices) return false;
            push(f);
            r /= f;
        }
    return r == 1;
}

template <typename T>
struct Cx {
    T x, y;
};
template <>
struct __align__(8) Cx<float> {      // a register pair: the operand of the packed FP32 instructions, one LDS.64 / STS.64
    float x, y;
};
template <typename T>
__device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T>
__device__ __forceinline__ Cx<T> cscale(Cx<T> a, T s) { return {a.x * s, a.y * s}; }
template <typename T>
__device__ __forceinline__ Cx<T> mul_neg_i(Cx<T> a) { return {a.y, -a.x}; }   // a * (-i)
// a + s b, a + (-i) b, a - (-i) b  (s real)
template <typename T>
__device__ __forceinline__ Cx<T> caxpy(T s, Cx<T> b, Cx<T> a) { return {a.x + s * b.x, a.y + s * b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cadd_negi(Cx<T> a, Cx<T> b) { return {a.x + b.y, a.y - b.x}; }
template <typename T>
__device__ __forceinline__ Cx<T> csub_negi(Cx<T> a, Cx<T> b) { return {a.x - b.y, a.y + b.x}; }

// ---- float32: Blackwell packed FP32 (crt/sm_100_rt.h: FADD2 / FMUL2 / FFMA2 on an aligned register pair).  A complex
// value IS a pair, so every complex add is one instruction and a complex product two: the swapped / sign-flipped
// operands these need (a.y, -a.x ...) are operand modifiers of the packed instructions (SASS `.F32x2.LO_HI.NP`), and a
// real factor is a broadcast operand (`R.F32` or an immediate).  The FMA pipe retires a packed instruction in two
// cycles, so the arithmetic throughput is that of the scalar form; what halves is the number of ISSUE SLOTS, which is
// what bounds the FFT kernels (tools/ubench/packed_pipes.cu).
__device__ __forceinline__ float2 f2(Cx<float> a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ Cx<float> cx(float2 a) { return {a.x, a.y}; }
__device__ __forceinline__ Cx<float> cadd(Cx<float> a, Cx<float> b) { return cx(__fadd2_rn(f2(a), f2(b))); }
__device__ __forceinline__ Cx<float> csub(Cx<float> a, Cx<float> b) { return cx(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
__device__ __forceinline__ Cx<float> cmul(Cx<float> a, Cx<float> b) {
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y));
    return cx(__ffma2_rn(f2(a), make_float2(b.x, b.x), t));
}
__device__ __forceinline__ Cx<float> cscale(Cx<float> a, float s) { return cx(__fmul2_rn(f2(a), make_float2(s, s))); }
__device__ __forceinline__ Cx<float> caxpy(float s, Cx<float> b, Cx<float> a) { return cx(__ffma2_rn(f2(b), make_float2(s, s), f2(a))); }
__device__ __forceinline__ Cx<float> cadd_negi(Cx<float> a, Cx<float> b) { return cx(__fadd2_rn(f2(a), make_float2(b.y, -b.x))); }
__device__ __forceinline__ Cx<float> csub_negi(Cx<float> a, Cx<float> b) { return cx(__fadd2_rn(f2(a), make_float2(-b.y, b.x))); }

// fill tw[j] = exp(-2 pi i j / n), j = 0..count-1, cooperatively by `nthreads` threads
template <typename T>
__device__ __forceinline__ void fill_twiddles(Cx<T>* tw, int n, int count, int rank, int nthreads) {
    for (int j = rank; j < count; j += nthreads) {
        double s, c;
        sincospi(-2.0 * static_cast<double>(j) / static_cast<double>(n), &s, &c);
        tw[j] = {static_cast<T>(c), static_cast<T>(s)};
    }
}

// Shared-memory index swizzle for the Stockham buffers: one slot of padding per 16 elements, so that the radix-R strided
// stores of the early passes (stride R elements = a multiple of 16 for the big radices: every lane on the same bank)
// spread over the banks.  PAD = 0 keeps the plain index (float64 transforms, exact-size buffers).
template <int PAD>
__device__ __forceinline__ int sidx(int i) { return PAD ? i + (i >> 4) : i; }
static inline size_t padded_len(size_t n) { return n + (n >> 4) + 1; }

// ---- register-resident butterflies (float32): primes 2 / 3 / 4 / 5 and Cooley-Tukey composites of them
using Cf = Cx<float>;
__device__ __forceinline__ void rdft2(Cf* a) {
    const Cf t = a[1];
    a[1] = csub(a[0], t);
    a[0] = cadd(a[0], t);
}
__device__ __forceinline__ void rdft3(Cf* a) {
    const float s = 0.86602540378443864676f;
    const Cf t1 = cadd(a[1], a[2]);
    const Cf t2 = caxpy(-0.5f, t1, a[0]);
    const Cf t3 = cscale(csub(a[1], a[2]), s);
    a[0] = cadd(a[0], t1);
    a[1] = cadd_negi(t2, t3);
    a[2] = csub_negi(t2, t3);
}
__device__ __forceinline__ void rdft4(Cf* a) {
    const Cf t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    const Cf t2 = cadd(a[1], a[3]), t3 = csub(a[1], a[3]);
    a[0] = cadd(t0, t2);
    a[2] = csub(t0, t2);
    a[1] = cadd_negi(t1, t3);
    a[3] = csub_negi(t1, t3);
}
__device__ __forceinline__ void rdft5(Cf* a) {      // 18 packed instructions
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const Cf p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
    const Cf p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
    const Cf a0 = a[0];
    a[0] = cadd(cadd(a0, p1), p2);
    const Cf u1 = caxpy(c2, p2, caxpy(c1, p1, a0));
    const Cf u2 = caxpy(c1, p2, caxpy(c2, p1, a0));
    const Cf t1 = caxpy(s2, m2, cscale(m1, s1));
    const Cf t2 = caxpy(-s1, m2, cscale(m1, s2));
    a[1] = cadd_negi(u1, t1);
    a[4] = csub_negi(u1, t1);
    a[2] = cadd_negi(u2, t2);
    a[3] = csub_negi(u2, t2);
}
template <int R>
__device__ __forceinline__ Cf rtw(int m);
template <>
__device__ __forceinline__ Cf rtw<6>(int m) { return {kCos6[m % 6], kNSin6[m % 6]}; }
template <>
__device__ __forceinline__ Cf rtw<8>(int m) { return {kCos8[m % 8], kNSin8[m % 8]}; }
template <>
__device__ __forceinline__ Cf rtw<10>(int m) { return {kCos10[m % 10], kNSin10[m % 10]}; }
template <>
__device__ __forceinline__ Cf rtw<12>(int m) { return {kCos12[m % 12], kNSin12[m % 12]}; }
template <>
__device__ __forceinline__ Cf rtw<16>(int m) { return {kCos16[m % 16], kNSin16[m % 16]}; }
template <int R>
__device__ __forceinline__ void rdft_prime(Cf* a);
template <>
__device__ __forceinline__ void rdft_prime<2>(Cf* a) { rdft2(a); }
template <>
__device__ __forceinline__ void rdft_prime<3>(Cf* a) { rdft3(a); }
template <>
__device__ __forceinline__ void rdft_prime<4>(Cf* a) { rdft4(a); }
template <>
__device__ __forceinline__ void rdft_prime<5>(Cf* a) { rdft5(a); }
// R = RA * RB, natural order in and out:
//   u = RA u2 + u1, t = t2 + RB t1:  y[t] = sum_u1 w_RA^{u1 t1} w_R^{u1 t2} DFT_RB(a[u1::RA])[t2]
template <int R, int RA, int RB>
__device__ __forceinline__ void rdft_composite(Cf* a) {
    Cf f[RA][RB];
#pragma unroll
    for (int u1 = 0; u1 < RA; ++u1) {
#pragma unroll
        for (int u2 = 0; u2 < RB; ++u2) f[u1][u2] = a[RA * u2 + u1];
        rdft_prime<RB>(f[u1]);
#pragma unroll
        for (int t2 = 1; t2 < RB; ++t2)
            if (u1 > 0) f[u1][t2] = cmul(f[u1][t2], rtw<R>(u1 * t2));
    }
#pragma unroll
    for (int t2 = 0; t2 < RB; ++t2) {
        Cf g[RA];
#pragma unroll
        for (int u1 = 0; u1 < RA; ++u1) g[u1] = f[u1][t2];
        rdft_prime<RA>(g);
#pragma unroll
        for (int t1 = 0; t1 < RA; ++t1) a[t2 + RB * t1] = g[t1];
    }
}
template <int R>
__device__ __forceinline__ void rdft(Cf* a);
template <>
__device__ __forceinline__ void rdft<6>(Cf* a) { rdft_composite<6, 2, 3>(a); }
template <>
__device__ __forceinline__ void rdft<8>(Cf* a) { rdft_composite<8, 2, 4>(a); }
template <>
__device__ __forceinline__ void rdft<10>(Cf* a) { rdft_composite<10, 2, 5>(a); }
template <>
__device__ __forceinline__ void rdft<12>(Cf* a) { rdft_composite<12, 3, 4>(a); }
template <>
__device__ __forceinline__ void rdft<16>(Cf* a) { rdft_composite<16, 4, 4>(a); }

// One Stockham pass of a composite radix (float32): same indexing as stockham_pass below
template <int R, int PAD>
__device__ __forceinline__ void stockham_pass_composite(const Cf* __restrict__ in, Cf* __restrict__ out, int n, int ns,
                                                        const Cf* __restrict__ tw, int r, int G) {
    const int nb = n / R;
    const int tstep = n / (ns * R);
    for (int j = r; j < nb; j += G) {
        const int k = j % ns;
        Cf a[R];
#pragma unroll
        for (int t = 0; t < R; ++t) a[t] = in[sidx<PAD>(j + t * nb)];
        if (ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) a[t] = cmul(a[t], tw[t * k * tstep]);
        }
        rdft<R>(a);
        const int o = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) out[sidx<PAD>(o + t * ns)] = a[t];
    }
}

// One Stockham pass of radix R: `in` -> `out`, both n complex values in shared memory.
// ns = product of the radices already applied.  Executed by G threads with rank r.
template <typename T, int R, int PAD>
__device__ __forceinline__ void stockham_pass(const Cx<T>* __restrict__ in, Cx<T>* __restrict__ out, int n, int ns,
                                              const Cx<T>* __restrict__ tw, int r, int G) {
    const int nb = n / R;                 // butterflies
    const int tstep = n / (ns * R);       // twiddle table stride for angle 2 pi k / (ns R)
    for (int j = r; j < nb; j += G) {
        const int k = j % ns;
        Cx<T> a[R];
#pragma unroll
        for (int t = 0; t < R; ++t) a[t] = in[sidx<PAD>(j + t * nb)];
        if (ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) a[t] = cmul(a[t], tw[t * k * tstep]);
        }
        Cx<T> y[R];
        if (R == 2) {
            y[0] = cadd(a[0], a[1]);
            y[1] = csub(a[0], a[1]);
        } else if (R == 3) {
            const T s = static_cast<T>(0.86602540378443864676);
            const Cx<T> t1 = cadd(a[1], a[2]);
            const Cx<T> t2 = {a[0].x - t1.x * static_cast<T>(0.5), a[0].y - t1.y * static_cast<T>(0.5)};
            const Cx<T> t3 = cscale(csub(a[1], a[2]), s);
            y[0] = cadd(a[0], t1);
            y[1] = {t2.x + t3.y, t2.y - t3.x};
            y[2] = {t2.x - t3.y, t2.y + t3.x};
        } else if (R == 4) {
            const Cx<T> t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
            const Cx<T> t2 = cadd(a[1], a[3]), t3 = mul_neg_i(csub(a[1], a[3]));
            y[0] = cadd(t0, t2);
            y[2] = csub(t0, t2);
            y[1] = cadd(t1, t3);
            y[3] = csub(t1, t3);
        } else if (R == 5) {
            const T c1 = static_cast<T>(0.30901699437494742410), c2 = static_cast<T>(-0.80901699437494742410);
            const T s1 = static_cast<T>(0.95105651629515357212), s2 = static_cast<T>(0.58778525229247312917);
            const Cx<T> p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
            const Cx<T> p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
            y[0] = {a[0].x + p1.x + p2.x, a[0].y + p1.y + p2.y};
            const Cx<T> u1 = {a[0].x + c1 * p1.x + c2 * p2.x, a[0].y + c1 * p1.y + c2 * p2.y};
            const Cx<T> u2 = {a[0].x + c2 * p1.x + c1 * p2.x, a[0].y + c2 * p1.y + c1 * p2.y};
            const Cx<T> v1 = mul_neg_i(Cx<T>{s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y});
            const Cx<T> v2 = mul_neg_i(Cx<T>{s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y});
            y[1] = cadd(u1, v1);
            y[4] = csub(u1, v1);
            y[2] = cadd(u2, v2);
            y[3] = csub(u2, v2);
        }
        const int o = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) out[sidx<PAD>(o + t * ns)] = y[t];
    }
}

// generic prime radix p (7..31): O(p^2) butterfly, DFT matrix entries from the twiddle table
template <typename T, int PAD>
__device__ __forceinline__ void stockham_pass_generic(const Cx<T>* __restrict__ in, Cx<T>* __restrict__ out, int n,
                                                      int ns, int p, const Cx<T>* __restrict__ tw, int r, int G) {
    const int nb = n / p;
    const int tstep = n / (ns * p);
    const int pstep = n / p;              // exp(-2 pi i u / p) = tw[u * pstep]
    for (int j = r; j < nb; j += G) {
        const int k = j % ns;
        const int o = (j - k) * p + k;
        for (int t = 0; t < p; ++t) {
            Cx<T> acc = in[sidx<PAD>(j)];
            for (int u = 1; u < p; ++u) {
                Cx<T> a = in[sidx<PAD>(j + u * nb)];
                if (ns > 1) a = cmul(a, tw[u * k * tstep]);
                acc = cadd(acc, cmul(a, tw[((t * u) % p) * pstep]));
            }
            out[sidx<PAD>(o + t * ns)] = acc;
        }
    }
}

// Full transform.  Returns the buffer that holds the result (a or b).  sync() separates passes.
template <typename T, int PAD = 0, typename SyncFn>
__device__ __forceinline__ Cx<T>* stockham_fft(Cx<T>* a, Cx<T>* b, const FftPlan& plan, const Cx<T>* tw, int r, int G,
                                               SyncFn sync) {
    int ns = 1;
    Cx<T>* in = a;
    Cx<T>* out = b;
    for (int i = 0; i < plan.n_radices; ++i) {
        const int R = plan.radix[i];
        switch (R) {
            case 2: stockham_pass<T, 2, PAD>(in, out, plan.n, ns, tw, r, G); break;
            case 3: stockham_pass<T, 3, PAD>(in, out, plan.n, ns, tw, r, G); break;
            case 4: stockham_pass<T, 4, PAD>(in, out, plan.n, ns, tw, r, G); break;
            case 5: stockham_pass<T, 5, PAD>(in, out, plan.n, ns, tw, r, G); break;
            case 6:
            case 8:
            case 10:
            case 12:
            case 16:
                if constexpr (sizeof(T) == 4) {            // composite radices exist for float32 plans only
                    if (R == 6) stockham_pass_composite<6, PAD>(in, out, plan.n, ns, tw, r, G);
                    else if (R == 8) stockham_pass_composite<8, PAD>(in, out, plan.n, ns, tw, r, G);
                    else if (R == 10) stockham_pass_composite<10, PAD>(in, out, plan.n, ns, tw, r, G);
                    else if (R == 12) stockham_pass_composite<12, PAD>(in, out, plan.n, ns, tw, r, G);
                    else stockham_pass_composite<16, PAD>(in, out, plan.n, ns, tw, r, G);
                }
                break;
            default: stockham_pass_generic<T, PAD>(in, out, plan.n, ns, R, tw, r, G); break;
        }
        ns *= R;
        sync();
        Cx<T>* t = in;
        in = out;
        out = t;
    }
    return in;
}

}  // namespace mhb
