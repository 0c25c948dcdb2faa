// Kernel 2 -- per-window FFT + PSD reducers (sm_100a, no cuFFT).
//
// Replaces the user-composed spectral chain of the reference:
//   view(x, W, S)                         src/mhealth/util/windows.py:20-33
//   mhealth.fft.fft(window)               src/mhealth/fft/__init__.py:3-7, fft/_fft.py:18-29
//   psd = |F|^2, one-sided bins 0..W/2, freqs = rfftfreq(W, 1/fs)
//   hrv.power_band / relative_power_band  src/mhealth/heart/hrv.py:173-198   (lo <= f <= hi)
//   density.peak_frequency                src/mhealth/generic/frequency/density.py:18-32 (lo <= f < hi, first max)
//   information.entropy(psd)              src/mhealth/generic/information.py:10-20
//
// One warp owns a window.  Even W: the real window is packed into W/2 complex points, transformed
// with the shared-memory Stockham FFT (fft_core.cuh) and untangled into the one-sided spectrum; odd
// W: a W-point complex transform with zero imaginary parts.  The window mean is removed before the
// transform (it only changes bin 0, which is restored exactly from the float64 sum) so that a large
// DC component -- gravity on an accelerometer axis -- does not eat the float32 FFT's dynamic range.
// The epilogue reduces the PSD row to the requested columns with warp shuffles; band and peak ranges
// arrive as integer bin ranges computed on the host with numpy's rfftfreq arithmetic, so the masks
// are bit-exact.
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "fft_core.cuh"
#include "psd_entropy.cuh"

namespace mhb {

int32_t spectral_fast_try(const float* x, const mhb_windows* geom, int64_t nw, double bin_hz, const int32_t* cols,
                          const int32_t* lo, const int32_t* hi, int32_t n_cols, void* out, int32_t out_f32,
                          int64_t o_series, int64_t o_window, int64_t o_col, void* stream);
int32_t spectral_w1920_try(const float* x, const mhb_windows* geom, int64_t nw, double bin_hz, const int32_t* cols,
                           const int32_t* lo, const int32_t* hi, int32_t n_cols, void* out, int32_t out_f32,
                           int64_t o_series, int64_t o_window, int64_t o_col, void* stream);
namespace {

constexpr int kMaxCols = 32;
constexpr int kWarps = 8;          // at most; large windows run fewer warps per CTA so that several CTAs share an SM

struct SpectralPlan {
    const float* x;
    int64_t series_len, series_stride, nw, total_windows;
    int32_t W, S, N, nb, even;        // N = transform length (W/2 or W), nb = W/2 + 1 bins
    FftPlan fft;
    double fs, bin_hz;                // bin_hz = 1 / (W * (1 / fs)), numpy's rfftfreq step
    void* out;                        // feature table (or raw PSD rows when n_cols == 0)
    int32_t out_f32;
    int64_t o_series, o_window, o_col;
    int32_t n_cols;
    int32_t col[kMaxCols];
    int32_t lo[kMaxCols], hi[kMaxCols];   // bin range [lo, hi) of the column
    int32_t redo_col;                     // >= 0: second pass -- only windows whose cell in this column holds the
                                          // redo marker (psd_entropy.cuh) are evaluated; -1: every window
};

__device__ __forceinline__ void put(const SpectralPlan& P, int64_t idx, double v) {
    if (P.out_f32) reinterpret_cast<float*>(P.out)[idx] = static_cast<float>(v);
    else reinterpret_cast<double*>(P.out)[idx] = v;
}

__global__ void __launch_bounds__(kWarps * 32) window_spectral_kernel(const SpectralPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = Cx<float>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = P.N;
    // layout: tw[N] | tw2[nb] (even W only) | per warp: bufA[N] bufB[N]
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* tw2 = tw + N;
    C* bufs = tw2 + (P.even ? P.nb : 0);
    const int NP = N + (N >> 4) + 1;                 // padded buffer length (fft_core.cuh sidx<1>)
    C* A = bufs + static_cast<size_t>(warp) * 2 * NP;
    C* B = A + NP;
    auto pd = [](int i) { return sidx<1>(i); };

    fill_twiddles<float>(tw, N, N, threadIdx.x, blockDim.x);
    if (P.even) fill_twiddles<float>(tw2, P.W, P.nb, threadIdx.x, blockDim.x);
    __syncthreads();

    const int n_warps = blockDim.x >> 5;
    auto do_window = [&](int64_t w) {
        const int64_t series = w / P.nw;
        const int64_t wi = w - series * P.nw;
        const float* src = P.x + series * P.series_stride + wi * P.S;

        // ---- ONE pass over global memory: stage the raw samples (packed) and take the float64 sum (exact DC);
        // the mean is then removed in shared memory
        double sum = 0.0;
        if (P.even) {
            if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
                const float2* s2 = reinterpret_cast<const float2*>(src);
#pragma unroll 4
                for (int i = lane; i < N; i += 32) {
                    const float2 v = s2[i];
                    A[pd(i)] = {v.x, v.y};
                    sum += static_cast<double>(v.x) + static_cast<double>(v.y);
                }
            } else {
#pragma unroll 4
                for (int i = lane; i < N; i += 32) {
                    const float a0 = src[2 * i], a1 = src[2 * i + 1];
                    A[pd(i)] = {a0, a1};
                    sum += static_cast<double>(a0) + static_cast<double>(a1);
                }
            }
        } else {
#pragma unroll 4
            for (int i = lane; i < N; i += 32) {
                const float a0 = src[i];
                A[pd(i)] = {a0, 0.f};
                sum += static_cast<double>(a0);
            }
        }
        sum = warp_sum(sum);
        const float mean = static_cast<float>(sum / P.W);
        __syncwarp();
        if (P.even) {
#pragma unroll 4
            for (int i = lane; i < N; i += 32) A[pd(i)] = {A[pd(i)].x - mean, A[pd(i)].y - mean};
        } else {
#pragma unroll 4
            for (int i = lane; i < N; i += 32) A[pd(i)].x -= mean;
        }
        __syncwarp();

        C* Z = stockham_fft<float, 1>(A, B, P.fft, tw, lane, 32, [] { __syncwarp(); });
        float* psd = reinterpret_cast<float*>(Z == A ? B : A);        // the free buffer, nb floats <= 2N floats

        // ---- one-sided power spectrum
        if (P.even) {
            for (int k = lane; k <= N; k += 32) {
                const C zk = Z[pd(k == N ? 0 : k)];
                const C zn = Z[pd(k == 0 ? 0 : N - k)];
                const C e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};          // (Z[k] + conj Z[N-k]) / 2
                const C o = {0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};         // (Z[k] - conj Z[N-k]) / (2i)
                const C t = cmul(o, tw2[k]);
                const float re = e.x + t.x, im = e.y + t.y;
                psd[k] = re * re + im * im;
            }
        } else {
            for (int k = lane; k < P.nb; k += 32) psd[k] = Z[pd(k)].x * Z[pd(k)].x + Z[pd(k)].y * Z[pd(k)].y;
        }
        __syncwarp();
        const double dc = sum * sum;              // bin 0 of the un-centred window, exact
        if (lane == 0) psd[0] = static_cast<float>(dc);
        __syncwarp();

        const int64_t obase = series * P.o_series + wi * P.o_window;
        if (P.n_cols == 0) {                      // raw PSD rows
            for (int k = lane; k < P.nb; k += 32) put(P, obase + k * P.o_col, k == 0 ? dc : static_cast<double>(psd[k]));
            __syncwarp();
            return;
        }

        // ---- reductions: total power first (relative power / entropy need it)
        double tot = 0.0;
        for (int k = lane; k < P.nb; k += 32) tot += (k == 0) ? dc : static_cast<double>(psd[k]);
        tot = warp_sum(tot);
        for (int j = 0; j < P.n_cols; ++j) {
            const int kind = P.col[j];
            double v = 0.0;
            if (kind == MHB_S_TOTAL_POWER) {
                v = tot;
            } else if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
                double b = 0.0;
                for (int k = P.lo[j] + lane; k < P.hi[j]; k += 32) b += (k == 0) ? dc : static_cast<double>(psd[k]);
                b = warp_sum(b);
                v = kind == MHB_S_BAND_POWER ? b : b / tot;
            } else if (kind == MHB_S_PEAK_FREQUENCY || kind == MHB_S_PEAK_BIN) {
                float best = -1.f;
                int arg = 0x7fffffff;
                for (int k = P.lo[j] + lane; k < P.hi[j]; k += 32) {
                    const float pv = psd[k];
                    if (pv > best) {            // strict: the first maximum of this lane's stride wins
                        best = pv;
                        arg = k;
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
                if (arg == 0x7fffffff) v = CUDART_NAN;                       // empty range
                else v = kind == MHB_S_PEAK_BIN ? static_cast<double>(arg) : static_cast<double>(arg) * P.bin_hz;
            } else if (kind == MHB_S_ENTROPY) {
                // bins 1.. in float32 with the fast log2 (absolute error 2^-22 on O(1) arguments); the DC term, which may
                // be within 1e-7 of 1 on a gravity axis, through log1p in float64
                const float inv = static_cast<float>(1.0 / tot);
                float h0 = 0.f, h1 = 0.f;
                int k = lane == 0 ? 32 : lane;
                for (; k + 32 < P.nb; k += 64) {
                    const float p = fmaf(psd[k], inv, 1e-30f), p2 = fmaf(psd[k + 32], inv, 1e-30f);
                    h0 = fmaf(p, log2_normal(p), h0);
                    h1 = fmaf(p2, log2_normal(p2), h1);
                }
                if (k < P.nb) {
                    const float p = fmaf(psd[k], inv, 1e-30f);
                    h0 = fmaf(p, log2_normal(p), h0);
                }
                double hs = warp_sum(static_cast<double>((h0 + h1) * 0.69314718055994530942f));
                const double p0 = dc / tot + 1e-30, qrest = (tot - dc) / tot;
                hs += p0 * (qrest < 0.5 ? log1p(-qrest) : log(p0));
                v = -hs;
                if (tot > 0.0 && v < static_cast<double>(kToneEntropy) && p0 < 0.5) {    // noiseless tone: psd_entropy.cuh
                    v = entropy_dominant_bin(
                        psd, P.nb, dc, tot, lane, 32,
                        [](float& b, int& a) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const float ob = __shfl_xor_sync(0xffffffffu, b, o);
                                const int oa = __shfl_xor_sync(0xffffffffu, a, o);
                                if (ob > b || (ob == b && oa < a)) {
                                    b = ob;
                                    a = oa;
                                }
                            }
                        },
                        [](double s) { return warp_sum(s); });
                }
            }
            if (lane == 0) put(P, obase + j * P.o_col, v);
        }
        __syncwarp();
    };

    const int64_t first = static_cast<int64_t>(blockIdx.x) * n_warps + warp, stride = static_cast<int64_t>(gridDim.x) * n_warps;
    if (P.redo_col < 0) {
        for (int64_t w = first; w < P.total_windows; w += stride) do_window(w);
    } else {
        // second pass behind a kernel that left redo markers: a warp scans 32 windows at a time (one cell per lane, so
        // the scan is not a chain of dependent loads) and evaluates the marked ones
        for (int64_t g = first * 32; g < P.total_windows; g += stride * 32) {
            const int64_t w = g + lane;
            bool marked = false;
            if (w < P.total_windows) {
                const int64_t series = w / P.nw;
                const int64_t cell = series * P.o_series + (w - series * P.nw) * P.o_window + P.redo_col * P.o_col;
                marked = P.out_f32 ? reinterpret_cast<const uint32_t*>(P.out)[cell] == kRedoMarkF32
                                   : reinterpret_cast<const unsigned long long*>(P.out)[cell] == kRedoMarkF64;
            }
            unsigned m = __ballot_sync(0xffffffffu, marked);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                do_window(g + b);
            }
        }
    }
}

// numpy.fft.rfftfreq(W, 1/fs)[k] = k * (1.0 / (W * d)), d = 1.0 / fs   (float64, same operation order)
inline double rfft_bin_hz(int W, double fs) {
    const double d = 1.0 / fs;
    return 1.0 / (static_cast<double>(W) * d);
}

// first bin with freq >= x (generic/frequency/density.py:9-14), nb if none
inline int first_bin_ge(double x, int nb, double bin_hz) {
    for (int k = 0; k < nb; ++k)
        if (x <= static_cast<double>(k) * bin_hz) return k;
    return nb;
}

int32_t spectral_launch(const float* x, const mhb_windows* geom, double fs, const int32_t* h_features,
                        const double* h_params, int32_t n_features, void* out, int32_t out_f32, int64_t o_series,
                        int64_t o_window, int64_t o_col, void* stream_v, const char* who) {
    MHB_REQUIRE(geom, MHB_E_ARG, "%s: null geometry", who);
    MHB_REQUIRE(geom->wsize >= 2 && geom->wstep >= 1, MHB_E_ARG, "%s: wsize must be >= 2 and wstep >= 1", who);
    MHB_REQUIRE(geom->n_series >= 0 && geom->series_len >= 0 && geom->series_stride >= geom->series_len,
                MHB_E_ARG, "%s: bad series geometry", who);
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxCols, MHB_E_ARG, "%s: 0..%d columns per call", who, kMaxCols);
    MHB_REQUIRE(fs > 0.0, MHB_E_ARG, "%s: sampling rate must be positive", who);
    const int64_t nw = n_windows_host(geom->series_len, geom->wsize, geom->wstep);
    if (nw == 0 || geom->n_series == 0) return MHB_OK;
    MHB_REQUIRE(x && out, MHB_E_ARG, "%s: null data/output pointer", who);

    SpectralPlan P;
    memset(&P, 0, sizeof(P));
    P.x = x;
    P.series_len = geom->series_len;
    P.series_stride = geom->series_stride;
    P.nw = nw;
    P.total_windows = nw * geom->n_series;
    P.W = geom->wsize;
    P.S = geom->wstep;
    P.even = (P.W % 2 == 0) ? 1 : 0;
    P.N = P.even ? P.W / 2 : P.W;
    P.nb = P.W / 2 + 1;
    MHB_REQUIRE(fft_plan(P.N, &P.fft, getenv("MHB_FFT_SMALL_RADIX") == nullptr), MHB_E_UNSUPPORTED,
                "%s: FFT length %d has a prime factor > %d", who, P.N, kMaxPrime);
    P.fs = fs;
    P.bin_hz = rfft_bin_hz(P.W, fs);
    P.out = out;
    P.out_f32 = out_f32;
    P.o_series = o_series;
    P.o_window = o_window;
    P.o_col = o_col;
    P.n_cols = n_features;
    P.redo_col = -1;
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features[j];
        MHB_REQUIRE(f >= MHB_S_TOTAL_POWER && f <= MHB_S_ENTROPY, MHB_E_FEATURE, "%s: unknown spectral column %d", who, f);
        P.col[j] = f;
        const double lo = h_params ? h_params[2 * j] : NAN, hi = h_params ? h_params[2 * j + 1] : NAN;
        if (f == MHB_S_BAND_POWER || f == MHB_S_REL_BAND_POWER) {
            // lo <= f <= hi, both inclusive (hrv.py:178-179); None -> min / max frequency
            P.lo[j] = std::isnan(lo) ? 0 : first_bin_ge(lo, P.nb, P.bin_hz);
            int h = P.nb;
            if (!std::isnan(hi)) {
                h = 0;
                for (int k = 0; k < P.nb; ++k)
                    if (static_cast<double>(k) * P.bin_hz <= hi) h = k + 1;
            }
            P.hi[j] = h;
        } else if (f == MHB_S_PEAK_FREQUENCY || f == MHB_S_PEAK_BIN) {
            // lidx = first f >= lower, uidx = first f >= upper, upper EXCLUSIVE (density.py:30-32)
            P.lo[j] = std::isnan(lo) ? 0 : first_bin_ge(lo, P.nb, P.bin_hz);
            P.hi[j] = std::isnan(hi) ? P.nb : first_bin_ge(hi, P.nb, P.bin_hz);
        } else {
            P.lo[j] = 0;
            P.hi[j] = P.nb;
        }
    }
    if (n_features > 0 && getenv("MHB_SPECTRAL_GENERIC") == nullptr) {
        // thread-resident fast paths: W = 500 / S = 250 (spectral_fast.cu) and W = 1920 / S = 64 (spectral_w1920.cu);
        // MHB_SPECTRAL_GENERIC forces the generic kernel below (tests compare the two)
        const int32_t sf = spectral_fast_try(x, geom, nw, P.bin_hz, P.col, P.lo, P.hi, n_features, out, out_f32,
                                             o_series, o_window, o_col, stream_v);
        if (sf != -100) return sf;
        const int32_t s9 = spectral_w1920_try(x, geom, nw, P.bin_hz, P.col, P.lo, P.hi, n_features, out, out_f32,
                                              o_series, o_window, o_col, stream_v);
        if (s9 != -100) {
            // that kernel keeps no PSD rows, so it leaves the redo marker in the entropy cells of noiseless-tone windows
            // (psd_entropy.cuh); the generic kernel below re-evaluates exactly those (a scan of one cell per window)
            for (int j = 0; j < n_features && P.redo_col < 0; ++j)
                if (P.col[j] == MHB_S_ENTROPY) P.redo_col = j;
            if (s9 != MHB_OK || P.redo_col < 0) return s9;
        }
    }
    // one warp per window, two N-point buffers per warp: pick the warps per CTA that pack the most warps on an SM
    const size_t shared_tab = sizeof(Cx<float>) * (static_cast<size_t>(P.N) + (P.even ? P.nb : 0));
    const size_t per_warp = sizeof(Cx<float>) * 2 * padded_len(static_cast<size_t>(P.N));
    MHB_REQUIRE(shared_tab + per_warp <= 200 * 1024, MHB_E_UNSUPPORTED, "%s: wsize=%d needs %zu bytes of shared memory", who,
                P.W, shared_tab + per_warp);
    int best_warps = 1, best_total = 0;
    for (int wv = 1; wv <= kWarps; ++wv) {
        const size_t sm = shared_tab + per_warp * wv;
        if (sm > 200 * 1024) break;
        int per_sm = static_cast<int>((227 * 1024) / (sm + 1024));
        if (per_sm > 16) per_sm = 16;
        if (per_sm * wv > 64) per_sm = 64 / wv;
        if (per_sm * wv >= best_total) {
            best_total = per_sm * wv;
            best_warps = wv;
        }
    }
    const size_t smem = shared_tab + per_warp * best_warps;
    cudaError_t e = cudaFuncSetAttribute(window_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e, who);
    int64_t ctas = (P.total_windows + best_warps - 1) / best_warps;
    int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * best_warps > 64) per_sm = 64 / best_warps;
    const int64_t max_ctas = static_cast<int64_t>(kNumSMs) * per_sm;
    if (ctas > max_ctas) ctas = max_ctas;
    window_spectral_kernel<<<static_cast<unsigned>(ctas), best_warps * 32, smem, static_cast<cudaStream_t>(stream_v)>>>(P);
    return cuda_status(cudaGetLastError(), who);
}

// ---------------------------------------------------------------------------------------------
// Batched complex128 DFT: the drop-in for fftw_fft(n, in, out, dir) / numpy.fft.fft on rows.
struct FftRowsPlan {
    const double* in;
    double* out;
    int64_t n_rows;
    int32_t n, in_is_complex, direction;
    FftPlan fft;
};

__global__ void __launch_bounds__(256) fft_rows_c128_kernel(const FftRowsPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = Cx<double>;
    const int n = P.n;
    C* tw = reinterpret_cast<C*>(smem_raw);
    C* A = tw + n;
    C* B = A + n;
    const int r = threadIdx.x, G = blockDim.x;
    fill_twiddles<double>(tw, n, n, r, G);
    __syncthreads();
    const bool inv = P.direction > 0;
    for (int64_t row = blockIdx.x; row < P.n_rows; row += gridDim.x) {
        const double* src = P.in + row * static_cast<int64_t>(n) * (P.in_is_complex ? 2 : 1);
        for (int i = r; i < n; i += G) {
            C v;
            if (P.in_is_complex) v = {src[2 * i], src[2 * i + 1]};
            else v = {src[i], 0.0};
            if (inv) v.y = -v.y;                 // ifft(x) = conj(fft(conj(x))) / n
            A[i] = v;
        }
        __syncthreads();
        C* Z = stockham_fft<double>(A, B, P.fft, tw, r, G, [] { __syncthreads(); });
        double* dst = P.out + row * static_cast<int64_t>(n) * 2;
        const double sc = inv ? 1.0 / n : 1.0;
        for (int i = r; i < n; i += G) {
            dst[2 * i] = Z[i].x * sc;
            dst[2 * i + 1] = (inv ? -Z[i].y : Z[i].y) * sc;
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// PSD reducers on caller-supplied rows (float64), the literal drop-in for hrv.power_band /
// relative_power_band (heart/hrv.py:173-198), density.peak_frequency (density.py:18-32) and
// information.entropy (information.py:10-20).  One warp per row; masks are evaluated on the
// caller's `freqs` values exactly as the reference does.
struct PsdReducePlan {
    const double* psd;
    const double* freqs;      // may be null when no column needs it
    int64_t n_rows;
    int32_t nb;
    double* out;              // [n_rows][n_cols]
    int32_t n_cols;
    int32_t col[kMaxCols];
    double lo[kMaxCols], hi[kMaxCols];      // NaN = None
};

__global__ void __launch_bounds__(256) psd_reduce_kernel(const PsdReducePlan P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp0; row < P.n_rows; row += nwarps) {
        const double* x = P.psd + row * P.nb;
        double tot_abs = 0.0, tot = 0.0;
        for (int k = lane; k < P.nb; k += 32) {
            tot_abs += fabs(x[k]);
            tot += x[k];
        }
        tot_abs = warp_sum(tot_abs);
        tot = warp_sum(tot);
        for (int j = 0; j < P.n_cols; ++j) {
            const int kind = P.col[j];
            double v = 0.0;
            if (kind == MHB_S_TOTAL_POWER) {
                v = tot_abs;
            } else if (kind == MHB_S_BAND_POWER || kind == MHB_S_REL_BAND_POWER) {
                const bool all_lo = isnan(P.lo[j]), all_hi = isnan(P.hi[j]);
                double b = 0.0;
                for (int k = lane; k < P.nb; k += 32) {
                    const bool keep = (all_lo || P.freqs[k] >= P.lo[j]) && (all_hi || P.freqs[k] <= P.hi[j]);
                    if (keep) b += fabs(x[k]);
                }
                b = warp_sum(b);
                v = kind == MHB_S_BAND_POWER ? b : b / tot_abs;
            } else if (kind == MHB_S_PEAK_FREQUENCY || kind == MHB_S_PEAK_BIN) {
                // first_index: first i with bound <= freqs[i] (freqs assumed ordered, density.py:9-14)
                int lidx = 0, uidx = P.nb;
                if (!isnan(P.lo[j])) {
                    int f = P.nb;
                    for (int k = lane; k < P.nb; k += 32)
                        if (P.lo[j] <= P.freqs[k]) { f = k; break; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
                    lidx = f;
                }
                if (!isnan(P.hi[j])) {
                    int f = P.nb;
                    for (int k = lane; k < P.nb; k += 32)
                        if (P.hi[j] <= P.freqs[k]) { f = k; break; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
                    uidx = f;
                }
                double best = -CUDART_INF;
                int arg = 0x7fffffff;
                for (int k = lidx + lane; k < uidx; k += 32)
                    if (x[k] > best || arg == 0x7fffffff) {
                        if (arg == 0x7fffffff || x[k] > best) { best = x[k]; arg = k; }
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) {
                        best = ob;
                        arg = oa;
                    }
                }
                if (arg == 0x7fffffff) v = CUDART_NAN;
                else v = kind == MHB_S_PEAK_BIN ? static_cast<double>(arg) : P.freqs[arg];
            } else if (kind == MHB_S_ENTROPY) {
                double h = 0.0;
                for (int k = lane; k < P.nb; k += 32) {
                    const double p = x[k] / tot + 1e-30;
                    h += p * log(p);
                }
                v = -warp_sum(h);
            }
            if (lane == 0) P.out[row * P.n_cols + j] = v;
        }
    }
}

}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_window_spectral_f32(const float* x, const mhb_windows* geom, double fs,
                                           const int32_t* h_features, const double* h_params, int32_t n_features,
                                           const mhb_table* table, void* stream) {
    MHB_REQUIRE(table, MHB_E_ARG, "window_spectral: null table");
    if (n_features == 0) return MHB_OK;
    MHB_REQUIRE(h_features, MHB_E_ARG, "window_spectral: null column list");
    return mhb::spectral_launch(x, geom, fs, h_features, h_params, n_features, table->out, table->out_f32,
                                table->series_stride, table->window_stride, table->column_stride, stream,
                                "window_spectral");
}

// Statistical (kernel 1a family) and spectral columns of the same windows in ONE call: kernel 1a and kernel 2 launched back
// to back on the caller's stream (see DESIGN.md section 4 for why the two are not one kernel: both are bound by
// instruction issue / latency, not by HBM, and the fused, warp-specialised variant measured 2x slower).
extern "C" int32_t mhb_window_features_f32(const float* x, const mhb_windows* geom, const int32_t* stat_features,
                                           int32_t n_stat, double zc_threshold, const mhb_table* stat_table, double fs,
                                           const int32_t* spec_features, const double* spec_params, int32_t n_spec,
                                           const mhb_table* spec_table, void* stream) {
    MHB_REQUIRE(n_stat >= 0 && n_spec >= 0, MHB_E_ARG, "window_features: negative column count");
    MHB_REQUIRE(n_stat == 0 || (stat_features && stat_table), MHB_E_ARG, "window_features: null statistics list / table");
    MHB_REQUIRE(n_spec == 0 || (spec_features && spec_table), MHB_E_ARG, "window_features: null spectral list / table");
    if (n_stat > 0) {
        const int32_t st = mhb_window_stats_f32(x, geom, stat_features, n_stat, zc_threshold, stat_table, stream);
        if (st != MHB_OK) return st;
    }
    if (n_spec > 0) return mhb_window_spectral_f32(x, geom, fs, spec_features, spec_params, n_spec, spec_table, stream);
    return MHB_OK;
}

extern "C" int32_t mhb_window_psd_f32(const float* x, const mhb_windows* geom, void* psd_out, int32_t out_f32,
                                      void* stream) {
    MHB_REQUIRE(geom, MHB_E_ARG, "window_psd: null geometry");
    const int64_t nw = mhb::n_windows_host(geom->series_len, geom->wsize, geom->wstep);
    const int64_t nb = geom->wsize / 2 + 1;
    return mhb::spectral_launch(x, geom, 1.0, nullptr, nullptr, 0, psd_out, out_f32, nw * nb, nb, 1, stream,
                                "window_psd");
}

extern "C" int32_t mhb_fft_c128(const double* in, int32_t in_is_complex, int64_t n_rows, int32_t n, int32_t direction,
                                double* out, void* stream) {
    MHB_REQUIRE(n >= 1 && n_rows >= 0, MHB_E_ARG, "fft_c128: bad sizes");
    MHB_REQUIRE(direction == -1 || direction == 1, MHB_E_ARG, "fft_c128: direction must be -1 (forward) or +1 (backward)");
    if (n_rows == 0) return MHB_OK;
    MHB_REQUIRE(in && out, MHB_E_ARG, "fft_c128: null pointer");
    mhb::FftRowsPlan P;
    memset(&P, 0, sizeof(P));
    P.in = in;
    P.out = out;
    P.n_rows = n_rows;
    P.n = n;
    P.in_is_complex = in_is_complex;
    P.direction = direction;
    MHB_REQUIRE(mhb::fft_plan(n, &P.fft), MHB_E_UNSUPPORTED, "fft_c128: length %d has a prime factor > %d", n,
                mhb::kMaxPrime);
    const size_t smem = sizeof(double) * 2 * 3 * static_cast<size_t>(n);
    MHB_REQUIRE(smem <= 200 * 1024, MHB_E_UNSUPPORTED, "fft_c128: length %d does not fit shared memory", n);
    cudaError_t e = cudaFuncSetAttribute(mhb::fft_rows_c128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return mhb::cuda_status(e, "fft_c128");
    int64_t ctas = n_rows < mhb::kNumSMs * 4 ? n_rows : mhb::kNumSMs * 4;
    mhb::fft_rows_c128_kernel<<<static_cast<unsigned>(ctas), 256, smem, static_cast<cudaStream_t>(stream)>>>(P);
    return mhb::cuda_status(cudaGetLastError(), "fft_c128");
}

extern "C" int32_t mhb_psd_reduce_f64(const double* psd, const double* freqs, int64_t n_rows, int32_t nb,
                                      const int32_t* h_features, const double* h_params, int32_t n_features,
                                      double* out, void* stream) {
    MHB_REQUIRE(n_rows >= 0 && nb >= 1, MHB_E_ARG, "psd_reduce: bad sizes");
    MHB_REQUIRE(n_features >= 0 && n_features <= mhb::kMaxCols, MHB_E_ARG, "psd_reduce: 0..%d columns per call", mhb::kMaxCols);
    if (n_rows == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(psd && out && h_features, MHB_E_ARG, "psd_reduce: null pointer");
    mhb::PsdReducePlan P;
    memset(&P, 0, sizeof(P));
    P.psd = psd;
    P.freqs = freqs;
    P.n_rows = n_rows;
    P.nb = nb;
    P.out = out;
    P.n_cols = n_features;
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features[j];
        MHB_REQUIRE(f >= MHB_S_TOTAL_POWER && f <= MHB_S_ENTROPY, MHB_E_FEATURE, "psd_reduce: unknown column %d", f);
        P.col[j] = f;
        P.lo[j] = h_params ? h_params[2 * j] : NAN;
        P.hi[j] = h_params ? h_params[2 * j + 1] : NAN;
        const bool needs_freqs = (f == MHB_S_PEAK_FREQUENCY) ||
                                 ((f != MHB_S_TOTAL_POWER && f != MHB_S_ENTROPY) && !(std::isnan(P.lo[j]) && std::isnan(P.hi[j])));
        MHB_REQUIRE(!needs_freqs || freqs, MHB_E_ARG, "psd_reduce: column %d needs the frequency vector", j);
    }
    int64_t ctas = (n_rows + 7) / 8;
    if (ctas > mhb::kNumSMs * 8) ctas = mhb::kNumSMs * 8;
    mhb::psd_reduce_kernel<<<static_cast<unsigned>(ctas), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
    return mhb::cuda_status(cudaGetLastError(), "psd_reduce");
}
