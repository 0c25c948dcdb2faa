// Accelerometer pre-stage: elementwise orientation / magnitude of tri-axial samples (sm_100a).
//
// Replaces the jitted elementwise functions of the reference (src/mhealth/inertial/accelerometer.py):
//   roll(y, z)          :13-25    arctan2(y, z) * 180 / pi
//   pitch(x, y, z)      :44-56    arctan2(-x, sqrt(y*y + z*z)) * 180 / pi
//   magnitude(x, y, z)  :198-225  sqrt(x**2 + y**2 + z**2)
//   magnitude_dot       :236-259  sqrt(x.x + y.y + z.z)  (one scalar)
// the step immediately before windowing for the accelerometer configurations.
// Result types follow numba's: magnitude keeps the input type (float32 arithmetic for float32 input, every
// product and sum rounded separately -- no FMA contraction -- so the result is bit-identical to the reference);
// roll / pitch evaluate arctan2 in the input type and widen to float64 for the scaling by 180 / pi.
// HBM-bound streaming kernels: 128-bit loads where the pointers allow it, grid = a multiple of the SM count.
#include <math_constants.h>

#include "common.cuh"

namespace mhb {

namespace {

enum { kOpMagnitude = 0, kOpRoll = 1, kOpPitch = 2 };

template <typename T>
struct Vec4;
template <>
struct Vec4<float> { using type = float4; };
template <>
struct Vec4<double> { using type = double4; };

__device__ __forceinline__ float mag1(float x, float y, float z) {
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
__device__ __forceinline__ double mag1(double x, double y, double z) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}
__device__ __forceinline__ double roll1(float y, float z) { return static_cast<double>(atan2f(y, z)) * 180.0 / CUDART_PI; }
__device__ __forceinline__ double roll1(double y, double z) { return atan2(y, z) * 180.0 / CUDART_PI; }
__device__ __forceinline__ double pitch1(float x, float y, float z) {
    return static_cast<double>(atan2f(-x, sqrtf(__fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))))) * 180.0 / CUDART_PI;
}
__device__ __forceinline__ double pitch1(double x, double y, double z) {
    return atan2(-x, sqrt(__dadd_rn(__dmul_rn(y, y), __dmul_rn(z, z)))) * 180.0 / CUDART_PI;
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) accel_elementwise_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                                const T* __restrict__ z, int64_t n, void* __restrict__ out) {
    using OutT = typename std::conditional<OP == kOpMagnitude, T, double>::type;
    OutT* o = reinterpret_cast<OutT*>(out);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (OP == kOpMagnitude) o[i] = static_cast<OutT>(mag1(x[i], y[i], z[i]));
        else if (OP == kOpRoll) o[i] = static_cast<OutT>(roll1(y[i], z[i]));
        else o[i] = static_cast<OutT>(pitch1(x[i], y[i], z[i]));
    }
}

// float32 magnitude, 4 samples per thread through 128-bit loads / stores (pointers 16-byte aligned)
__global__ void __launch_bounds__(256) accel_magnitude_v4_kernel(const float4* __restrict__ x, const float4* __restrict__ y,
                                                                 const float4* __restrict__ z, int64_t n4,
                                                                 float4* __restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = x[i], b = y[i], c = z[i];
        float4 r;
        r.x = mag1(a.x, b.x, c.x);
        r.y = mag1(a.y, b.y, c.y);
        r.z = mag1(a.z, b.z, c.z);
        r.w = mag1(a.w, b.w, c.w);
        out[i] = r;
    }
}

// sum of squares of three arrays in float64: per-CTA partials (fixed order), then one CTA folds them
template <typename T>
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                            const T* __restrict__ z, int64_t n, double* __restrict__ part) {
    __shared__ double sh[8];
    double acc = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = static_cast<double>(x[i]), b = static_cast<double>(y[i]), c = static_cast<double>(z[i]);
        acc = fma(a, a, acc);
        acc = fma(b, b, acc);
        acc = fma(c, c, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        part[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(const double* __restrict__ part, int n_part, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_part; i += 32) acc += part[i];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) out[0] = sqrt(acc);
}

int64_t grid_for(int64_t items) {
    int64_t blocks = (items + 255) / 256;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : blocks;
}

template <typename T>
int32_t elementwise_impl(int32_t op, const T* x, const T* y, const T* z, int64_t n, void* out, cudaStream_t stream) {
    const unsigned blocks = static_cast<unsigned>(grid_for(n));
    if (op == kOpMagnitude) {
        if constexpr (sizeof(T) == 4) {
            const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                 reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out);
            if ((al & 15) == 0 && n >= 4) {
                const int64_t n4 = n / 4;
                accel_magnitude_v4_kernel<<<static_cast<unsigned>(grid_for(n4)), 256, 0, stream>>>(
                    reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y),
                    reinterpret_cast<const float4*>(z), n4, reinterpret_cast<float4*>(out));
                const int64_t done = n4 * 4;
                if (done < n)
                    accel_elementwise_kernel<T, kOpMagnitude><<<1, 256, 0, stream>>>(
                        x + done, y + done, z + done, n - done, reinterpret_cast<T*>(out) + done);
                return cuda_status(cudaGetLastError(), "accel magnitude launch");
            }
        }
        accel_elementwise_kernel<T, kOpMagnitude><<<blocks, 256, 0, stream>>>(x, y, z, n, out);
    } else if (op == kOpRoll) {
        accel_elementwise_kernel<T, kOpRoll><<<blocks, 256, 0, stream>>>(x, y, z, n, out);
    } else {
        accel_elementwise_kernel<T, kOpPitch><<<blocks, 256, 0, stream>>>(x, y, z, n, out);
    }
    return cuda_status(cudaGetLastError(), "accel elementwise launch");
}

}  // namespace
}  // namespace mhb

extern "C" int32_t mhb_accel_elementwise(int32_t op, int32_t is_f64, const void* x, const void* y, const void* z,
                                         int64_t n, void* out, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(op >= kOpMagnitude && op <= kOpPitch, MHB_E_FEATURE, "accel_elementwise: unknown op %d", op);
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "accel_elementwise: negative size");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(y && z && out && (x || op == kOpRoll), MHB_E_ARG, "accel_elementwise: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_f64)
        return elementwise_impl<double>(op, static_cast<const double*>(x), static_cast<const double*>(y),
                                        static_cast<const double*>(z), n, out, s);
    return elementwise_impl<float>(op, static_cast<const float*>(x), static_cast<const float*>(y),
                                   static_cast<const float*>(z), n, out, s);
}

extern "C" int64_t mhb_accel_sumsq_workspace(int64_t n) { return mhb::grid_for(n); }

extern "C" int32_t mhb_accel_magnitude_dot(int32_t is_f64, const void* x, const void* y, const void* z, int64_t n,
                                           double* workspace, int64_t workspace_len, double* out, void* stream) {
    using namespace mhb;
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "accel_magnitude_dot: negative size");
    MHB_REQUIRE(out && workspace && (n == 0 || (x && y && z)), MHB_E_ARG, "accel_magnitude_dot: null pointer");
    const int64_t blocks = grid_for(n);
    MHB_REQUIRE(workspace_len >= blocks, MHB_E_WORKSPACE, "accel_magnitude_dot: workspace of %lld doubles needed",
                static_cast<long long>(blocks));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_f64)
        sumsq_partial_kernel<double><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
            static_cast<const double*>(x), static_cast<const double*>(y), static_cast<const double*>(z), n, workspace);
    else
        sumsq_partial_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
            static_cast<const float*>(x), static_cast<const float*>(y), static_cast<const float*>(z), n, workspace);
    sumsq_final_kernel<<<1, 32, 0, s>>>(workspace, static_cast<int>(blocks), out);
    return cuda_status(cudaGetLastError(), "accel_magnitude_dot launch");
}
