// Warp-level bitonic sorting network on register-resident elements (shared by the order-statistics kernels).
#pragma once
#include "common.cuh"

namespace mhb {

// Bitonic sort of 32 * EPL elements held by a warp, EPL consecutive elements per lane (element e = lane * EPL + i):
// compare-exchanges at distance < EPL stay in registers (two FMNMX), larger distances are one shuffle per element.
// Every index is a compile-time constant, so v[] lives in registers.
// One FMNMX / DMNMX per call: a compare-and-select (`b < a ? b : a`) costs FSETP + FSEL and made the ALU pipe the bound
// of the block sort (ncu, round 2: 76 % of the executed instructions were FSETP / FSEL, ALU pipe 86 % busy).  NaN is
// outside the contract of the order statistics (numba's percentile has its own NaN path).
template <typename T>
__device__ __forceinline__ T tmin2(T a, T b);
template <>
__device__ __forceinline__ float tmin2<float>(float a, float b) { return fminf(a, b); }
template <>
__device__ __forceinline__ double tmin2<double>(double a, double b) { return fmin(a, b); }
template <typename T>
__device__ __forceinline__ T tmax2(T a, T b);
template <>
__device__ __forceinline__ float tmax2<float>(float a, float b) { return fmaxf(a, b); }
template <>
__device__ __forceinline__ double tmax2<double>(double a, double b) { return fmax(a, b); }

template <typename T, int EPL>
__device__ __forceinline__ void warp_sort_regs(T* v, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * EPL; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j >= EPL) {
                const int mask = j / EPL;                           // partner lane
                const bool lower = (lane & mask) == 0;
                const bool up = (lane & (k2 / EPL)) == 0;            // k2 >= 2 j >= 2 EPL here
                const bool keep_min = lower == up;
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const T o = __shfl_xor_sync(0xffffffffu, v[i], mask);
                    v[i] = keep_min ? tmin2(v[i], o) : tmax2(v[i], o);
                }
            } else {
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const int l = i ^ j;
                    if (l > i) {
                        // direction of element e = lane * EPL + i: bit k2 of e (a lane bit when k2 >= EPL)
                        const bool up = k2 >= EPL ? (lane & (k2 / EPL)) == 0 : (i & k2) == 0;
                        const T a = v[i], c = v[l];
                        const T mn = tmin2(a, c), mx = tmax2(a, c);
                        v[i] = up ? mn : mx;
                        v[l] = up ? mx : mn;
                    }
                }
            }
        }
    }
}

// sort the p2 = 32 * EPL elements of buf[] (shared memory, already padded) in place
template <typename T, int EPL>
__device__ __forceinline__ void sort_smem_via_regs(T* __restrict__ buf, int lane) {
    T v[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = buf[lane * EPL + i];
    warp_sort_regs<T, EPL>(v, lane);
#pragma unroll
    for (int i = 0; i < EPL; ++i) buf[lane * EPL + i] = v[i];
}

}  // namespace mhb
