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

// ---- float32 specialisation: sign-state network on GROUPS of lanes.
// Measured on the B200 (tools/ubench/sort_pipes.cu, 8 warps per sub-partition): SHFL issues once every 4.0 cycles per
// sub-partition (one warp shuffle per cycle and SM), FMNMX / FMUL once every ~1.0; the generic network above needs 1130
// cycles per 256-element block and sub-partition.  Two changes:
// (1) Sign states.  The generic network spends 2 instructions per cross-lane compare-exchange (FMNMX + predicated FMNMX)
//     and 4 per in-register one whose direction depends on the lane (2 FMNMX + 2 FSEL).  Here every lane holds
//     u = tau * v with a sign state tau = +-1:
//       * in-register stages of a merge level whose direction is `down` for this lane: tau = -1, so the ascending
//         compare-exchange (min to the lower index) of u IS the descending one of v -- no selects;
//       * cross-lane stage: tau = +1 on lanes that keep the minimum and -1 on lanes that keep the maximum; partners
//         then hold opposite states and  u = min(u, -shfl(u))  is  tau * (min or max of the two values): ONE FMNMX per
//         element (the negation is an operand modifier).
//     Changing tau is one FMUL by +-1 per element (exact).  Full warp, 8 elements per lane: 758 cycles per block.
// (2) Fewer, longer lanes.  A block is sorted by a GROUP of GL lanes holding EPL = P2 / GL elements each (a warp sorts
//     32 / GL blocks at once): the lane's own EPL elements are sorted by Batcher's odd-even merge network in registers
//     (19 / 63 / 191 comparators for 8 / 16 / 32 elements), and only log2(GL) (log2(GL) + 1) / 2 merge stages cross lanes --
//     GL = 8, EPL = 32: 48 shuffles per 256-element block instead of 120.
// (-0.0 and +0.0 may come out in either order; they are equal as values.  NaN is outside the contract.)
__device__ __forceinline__ void cex_f32(float& lo, float& hi) {     // ascending compare-exchange
    const float a = lo, c = hi;
    lo = fminf(a, c);
    hi = fmaxf(a, c);
}
// The same with the maximum formed on the FMA pipe: bits(max) = bits(a) + bits(c) - bits(min) (exact: the minimum IS one
// of the two bit patterns), as two IMADs whose multipliers +1 / -1 are RUNTIME values (`one`; with a literal ptxas folds
// them into one IADD3, which is the ALU pipe again).  FMNMX issues once every 2 cycles per sub-partition (ALU pipe, 16
// lanes), IMAD likewise on the heavy FMA pipe: a network that uses this form for one compare-exchange out of three
// measured best (630 -> 595 cycles per 256-element block; two out of three: 614; tools/ubench/sort_pipes.cu).
__device__ __forceinline__ void cex_f32_fma(float& lo, float& hi, int one) {
    const float a = lo, c = hi;
    const float mn = fminf(a, c);
    int t, m;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(__float_as_int(a)), "r"(one), "r"(__float_as_int(c)));
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(m) : "r"(__float_as_int(mn)), "r"(-one), "r"(t));
    lo = mn;
    hi = __int_as_float(m);
}
template <int SEL>
__device__ __forceinline__ void cex_mix(float& lo, float& hi, int one) {
#ifndef MHB_CEX_FMA_OF
#define MHB_CEX_FMA_OF 3          // share of FMA-pipe compare-exchanges: MHB_CEX_FMA_N out of MHB_CEX_FMA_OF
#define MHB_CEX_FMA_N 1
#endif
    if constexpr (SEL % MHB_CEX_FMA_OF >= MHB_CEX_FMA_N) cex_f32(lo, hi);
    else cex_f32_fma(lo, hi, one);
}

// compare-exchanges (i, i + R) for i = I0, I0 + 2 R, ... while i + R < END
template <int I0, int END, int R, bool MIX>
__device__ __forceinline__ void oem_row_f32(float* u, int one) {
    if constexpr (I0 + R < END) {
        if constexpr (MIX) cex_mix<I0 / R>(u[I0], u[I0 + R], one);
        else cex_f32(u[I0], u[I0 + R]);
        oem_row_f32<I0 + 2 * R, END, R, MIX>(u, one);
    }
}

// Batcher's odd-even merge sort of u[LO .. LO + N), ascending, N = 2^m; template recursion, so that every index is a
// compile-time constant and u[] stays in registers.  `one` == 1 at run time (see cex_f32_fma); one == 0 selects the
// plain compare-exchanges everywhere.
template <int LO, int N, int R, bool MIX>
__device__ __forceinline__ void oem_merge_f32(float* u, int one) {
    if constexpr (2 * R < N) {
        oem_merge_f32<LO, N, 2 * R, MIX>(u, one);
        oem_merge_f32<LO + R, N, 2 * R, MIX>(u, one);
        oem_row_f32<LO + R, LO + N, R, MIX>(u, one);
    } else {
        if constexpr (MIX) cex_mix<(LO + R) / 1 + 1>(u[LO], u[LO + R], one);
        else cex_f32(u[LO], u[LO + R]);
    }
}
template <int LO, int N, bool MIX>
__device__ __forceinline__ void oem_sort_f32(float* u, int one) {
    if constexpr (N > 1) {
        oem_sort_f32<LO, N / 2, MIX>(u, one);
        oem_sort_f32<LO + N / 2, N / 2, MIX>(u, one);
        oem_merge_f32<LO, N, 1, MIX>(u, one);
    }
}
template <int N, bool MIX>
__device__ __forceinline__ void lane_sort_f32(float* u, int one) { oem_sort_f32<0, N, MIX>(u, one); }

// Sorts the GL * EPL elements held by each aligned group of GL lanes (element e = l * EPL + i on lane l of the group,
// i = 0 .. EPL - 1) ascending.  l = lane % GL.  The input may be in any order (any element-to-lane mapping).
template <int J, int I, int EPL, bool MIX>
__device__ __forceinline__ void merge_stage_f32(float* u, int one) {      // in-register bitonic merge stage, distance J
    if constexpr (I < EPL) {
        if constexpr ((I ^ J) > I) {
            if constexpr (MIX) cex_mix<I + J>(u[I], u[I ^ J], one);
            else cex_f32(u[I], u[I ^ J]);
        }
        merge_stage_f32<J, I + 1, EPL, MIX>(u, one);
    }
}
template <int J, int EPL, bool MIX>
__device__ __forceinline__ void merge_stages_f32(float* u, int one) {
    if constexpr (J > 0) {
        merge_stage_f32<J, 0, EPL, MIX>(u, one);
        merge_stages_f32<J / 2, EPL, MIX>(u, one);
    }
}

template <int EPL, int GL, bool MIX = false>
__device__ __forceinline__ void group_sort_regs_f32(float* u, int l, int one = 0) {
    bool neg = false;                                   // tau == -1
    auto flip = [&](bool req) {
        const float s = (req != neg) ? -1.0f : 1.0f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) u[i] = __fmul_rn(u[i], s);
        neg = req;
    };
    // merge levels up to EPL: the lane's own elements, descending on odd lanes (as level EPL of the bitonic scheme wants)
    if (GL > 1) flip((l & 1) != 0);
    lane_sort_f32<EPL, MIX>(u, one);
#pragma unroll
    for (int k2 = 2 * EPL; k2 <= GL * EPL; k2 <<= 1) {
        const bool dn = k2 < GL * EPL ? (l & (k2 / EPL)) != 0 : false;     // direction of this merge level for the lane
#pragma unroll
        for (int j = k2 >> 1; j >= EPL; j >>= 1) {
            const int mask = j / EPL;
            const bool upper = (l & mask) != 0;
            flip(dn != upper);                          // keep-min lanes +1, keep-max lanes -1
#pragma unroll
            for (int i = 0; i < EPL; ++i) u[i] = fminf(u[i], -__shfl_xor_sync(0xffffffffu, u[i], mask));
        }
        flip(dn);
        merge_stages_f32<EPL / 2, EPL, MIX>(u, one);
    }
    flip(false);
}

template <int EPL>
__device__ __forceinline__ void warp_sort_regs_f32(float* u, int lane) { group_sort_regs_f32<EPL, 32>(u, lane); }

template <int EPL>
__device__ __forceinline__ void warp_sort_regs_dispatch(float* v, int lane) { warp_sort_regs_f32<EPL>(v, lane); }
template <int EPL>
__device__ __forceinline__ void warp_sort_regs_dispatch(double* v, int lane) { warp_sort_regs<double, EPL>(v, lane); }

// sort the p2 = 32 * EPL elements of buf[] (shared memory, already padded) in place
template <typename T, int EPL>
__device__ __forceinline__ void sort_smem_via_regs(T* __restrict__ buf, int lane) {
    T v[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = buf[lane * EPL + i];
    warp_sort_regs_dispatch<EPL>(v, lane);
#pragma unroll
    for (int i = 0; i < EPL; ++i) buf[lane * EPL + i] = v[i];
}

}  // namespace mhb
