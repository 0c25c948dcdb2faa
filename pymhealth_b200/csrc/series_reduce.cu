// Successive-difference statistics of one series (sm_100a) -- the reductions behind the HRV time-domain metrics of
// the reference (src/mhealth/heart/hrv.py): pnn50 / pnnx :111-136 (share of |diff| above a threshold), rmssd :139-147
// (sqrt(mean(diff^2))), ssd :150-158 (sum(diff)), sdsd :161-170 (std(diff)).  np.diff is never materialised: every
// thread forms x[i+1] - x[i] in float64 on the fly and accumulates shifted power sums (pivot = the first difference);
// per-CTA partial records are folded in a fixed order by the last stage, so results are deterministic.
#include "common.cuh"

namespace mhb {

namespace {

struct DiffRec {
    double s1, s2;      // shifted power sums of the differences x[i+1] - x[i]
    double t1, t2;      // ... and of the pair sums x[i+1] + x[i] (Poincare SD2, hrv.py:219-231)
    long long cnt;
};

__global__ void __launch_bounds__(256) diff_partial_kernel(const double* __restrict__ x, int64_t nd, double thr,
                                                           DiffRec* __restrict__ part) {
    __shared__ DiffRec sh[8];
    const double c = x[1] - x[0], c2 = x[1] + x[0];
    double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
    long long cnt = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nd; i += stride) {
        const double a = x[i], b = x[i + 1];
        const double d = b - a;
        const double e = d - c;
        s1 += e;
        s2 = fma(e, e, s2);
        const double f = (b + a) - c2;
        t1 += f;
        t2 = fma(f, f, t2);
        cnt += fabs(d) > thr ? 1 : 0;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = {s1, s2, t1, t2, cnt};
    __syncthreads();
    if (threadIdx.x == 0) {
        DiffRec t = {0.0, 0.0, 0.0, 0.0, 0};
        for (int i = 0; i < 8; ++i) {
            t.s1 += sh[i].s1;
            t.s2 += sh[i].s2;
            t.t1 += sh[i].t1;
            t.t2 += sh[i].t2;
            t.cnt += sh[i].cnt;
        }
        part[blockIdx.x] = t;
    }
}

__global__ void diff_final_kernel(const double* __restrict__ x, int64_t nd, const DiffRec* __restrict__ part, int n_part,
                                  double* __restrict__ out) {
    if (threadIdx.x != 0) return;
    DiffRec t = {0.0, 0.0, 0.0, 0.0, 0};
    for (int i = 0; i < n_part; ++i) {
        t.s1 += part[i].s1;
        t.s2 += part[i].s2;
        t.t1 += part[i].t1;
        t.t2 += part[i].t2;
        t.cnt += part[i].cnt;
    }
    const double c = x[1] - x[0];
    const double n = static_cast<double>(nd);
    const double dl = t.s1 / n;
    const double mean = c + dl;
    double m2 = t.s2 - t.s1 * dl;
    if (m2 < 0.0) m2 = 0.0;
    out[0] = n;
    out[1] = mean * n;                 // sum(diff)
    out[2] = mean;
    out[3] = m2 / n;                   // population variance of diff
    out[4] = static_cast<double>(t.cnt);
    out[5] = m2 / n + mean * mean;     // mean(diff^2)
    double p2 = t.t2 - t.t1 * (t.t1 / n);
    if (p2 < 0.0) p2 = 0.0;
    out[6] = p2 / n;                   // population variance of the pair sums x[i+1] + x[i]
}

// ppg.slope_sum (src/mhealth/heart/ppg.py:28-42): out[i] = sum(dx[i-w : i]) for w <= i < n - 1, 0 elsewhere,
// dx = np.diff(x).  One thread per output sample adds its w differences in index order (float64), so every sample
// is read from L1/L2 w times but from HBM once; w is ~0.15 s of signal (9 samples at 64 Hz).
template <typename T>
__global__ void __launch_bounds__(256) slope_sum_kernel(const T* __restrict__ x, int64_t n, int32_t w,
                                                        double* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double acc = 0.0;
        if (i >= w && i < n - 1) {
            double prev = static_cast<double>(x[i - w]);
            for (int64_t j = i - w + 1; j <= i; ++j) {
                const double cur = static_cast<double>(x[j]);
                acc += cur - prev;
                prev = cur;
            }
        }
        out[i] = acc;
    }
}

// timedom.gradient (src/mhealth/generic/timedom.py:11-31): out[0] = x[1] - x[0], out[n-1] = x[n-1] - x[n-2],
// out[i] = (x[i+1] - x[i-1]) / 2; the difference is formed in the input type (float32 input: float32 subtraction), the
// result is float64.
template <typename T>
__global__ void __launch_bounds__(256) gradient_kernel(const T* __restrict__ x, int64_t n, double* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v;
        if (i == 0) v = static_cast<double>(x[1] - x[0]);
        else if (i == n - 1) v = static_cast<double>(x[n - 1] - x[n - 2]);
        else v = static_cast<double>(x[i + 1] - x[i - 1]) / 2;
        out[i] = v;
    }
}

// timedom.zero_crossings (timedom.py:34-49): samples with |x| <= th count as zero, pos = x > 0, out[i] = pos[i] xor
// pos[i+1] (n - 1 flags, one byte each).
template <typename T>
__global__ void __launch_bounds__(256) zero_crossings_kernel(const T* __restrict__ x, int64_t n, double th,
                                                             uint8_t* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n - 1; i += stride) {
        const double a = static_cast<double>(x[i]), b = static_cast<double>(x[i + 1]);
        const bool pa = a > 0.0 && !(fabs(a) <= th), pb = b > 0.0 && !(fabs(b) <= th);
        out[i] = pa != pb ? 1 : 0;
    }
}

static int64_t grid_1d(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : blocks;
}

}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_gradient(int32_t is_f64, const void* x, int64_t n, double* out, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 2, MHB_E_ARG, "gradient: at least two samples are needed (n = %lld)", static_cast<long long>(n));
    MHB_REQUIRE(x && out, MHB_E_ARG, "gradient: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>(grid_1d(n));
    if (is_f64) gradient_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(x), n, out);
    else gradient_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(x), n, out);
    return cuda_status(cudaGetLastError(), "gradient launch");
}

extern "C" int32_t mhb_zero_crossings(int32_t is_f64, const void* x, int64_t n, double threshold, uint8_t* out,
                                      void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "zero_crossings: negative size");
    if (n < 2) return MHB_OK;
    MHB_REQUIRE(x && out, MHB_E_ARG, "zero_crossings: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>(grid_1d(n - 1));
    if (is_f64) zero_crossings_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(x), n, threshold, out);
    else zero_crossings_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(x), n, threshold, out);
    return cuda_status(cudaGetLastError(), "zero_crossings launch");
}

extern "C" int32_t mhb_slope_sum(int32_t is_f64, const void* x, int64_t n, int32_t w, double* out, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 0 && w >= 0, MHB_E_ARG, "slope_sum: negative size");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(x && out, MHB_E_ARG, "slope_sum: null pointer");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_f64) slope_sum_kernel<double><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const double*>(x), n, w, out);
    else slope_sum_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(x), n, w, out);
    return cuda_status(cudaGetLastError(), "slope_sum launch");
}

extern "C" int64_t mhb_diff_stats_workspace(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = static_cast<int64_t>(mhb::kNumSMs) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return blocks * 5;                 // doubles (one 40-byte record per CTA)
}

extern "C" int32_t mhb_diff_stats_f64(const double* x, int64_t n, double abs_threshold, double* workspace,
                                      int64_t workspace_len, double* out7, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 2, MHB_E_ARG, "diff_stats: at least two samples are needed (n = %lld)", static_cast<long long>(n));
    MHB_REQUIRE(x && workspace && out7, MHB_E_ARG, "diff_stats: null pointer");
    const int64_t need = mhb_diff_stats_workspace(n - 1);
    MHB_REQUIRE(workspace_len >= need, MHB_E_WORKSPACE, "diff_stats: workspace of %lld doubles needed",
                static_cast<long long>(need));
    const int blocks = static_cast<int>(need / 5);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DiffRec* part = reinterpret_cast<DiffRec*>(workspace);
    diff_partial_kernel<<<blocks, 256, 0, s>>>(x, n - 1, abs_threshold, part);
    diff_final_kernel<<<1, 32, 0, s>>>(x, n - 1, part, blocks, out7);
    return cuda_status(cudaGetLastError(), "diff_stats launch");
}

// ---------------------------------------------------------------------------------------------
// Raw sensor counts -> float32 samples on the device.  rolling_apply takes any numeric dtype (util/windows.py:74-91
// casts the window to the reducer's input type); wearable loggers store int16 counts, so a host-side caller ships
// 2 bytes per sample over PCIe and widens here (exact: |count| < 2^15, and `scale` is applied as one rounded
// multiplication, e.g. 1 / 4096 g per count).  HBM-bound: 8 counts per thread through one 128-bit load.
namespace mhb {
namespace {
__global__ void __launch_bounds__(256) widen_i16_kernel(const int16_t* __restrict__ in, int64_t n, float scale,
                                                         float* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t n8 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) ? n / 8 : 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 q = reinterpret_cast<const int4*>(in)[i];
        const int v[4] = {q.x, q.y, q.z, q.w};
        float4 a, b;
        a.x = static_cast<float>(static_cast<short>(v[0] & 0xffff)) * scale;
        a.y = static_cast<float>(v[0] >> 16) * scale;
        a.z = static_cast<float>(static_cast<short>(v[1] & 0xffff)) * scale;
        a.w = static_cast<float>(v[1] >> 16) * scale;
        b.x = static_cast<float>(static_cast<short>(v[2] & 0xffff)) * scale;
        b.y = static_cast<float>(v[2] >> 16) * scale;
        b.z = static_cast<float>(static_cast<short>(v[3] & 0xffff)) * scale;
        b.w = static_cast<float>(v[3] >> 16) * scale;
        reinterpret_cast<float4*>(out)[2 * i] = a;
        reinterpret_cast<float4*>(out)[2 * i + 1] = b;
    }
    for (int64_t i = n8 * 8 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = static_cast<float>(in[i]) * scale;
}
}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_widen_i16_f32(const int16_t* in, int64_t n, float scale, float* out, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "widen_i16: negative size");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(in && out, MHB_E_ARG, "widen_i16: null pointer");
    int64_t blocks = (n / 8 + 255) / 256 + 1;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
    if (blocks > cap) blocks = cap;
    widen_i16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, scale, out);
    return cuda_status(cudaGetLastError(), "widen_i16 launch");
}
