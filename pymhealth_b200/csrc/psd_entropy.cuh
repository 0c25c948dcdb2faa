// Spectral entropy of a window whose power sits in ONE bin other than bin 0 (a noiseless tone).
//
// information.entropy (reference generic/information.py:10-20) is H = -sum p_k ln p_k.  When one p_m is within
// 1e-2 of 1 the straightforward float32 evaluation carries an absolute error of ~1e-7 from the term p_m ln p_m
// (log of a float next to 1), which is a RELATIVE error of 1e-3 on H = 1e-4.  The spectral kernels therefore switch,
// for such windows only (H < kToneEntropy and p_0 < 1/2 -- a dominant bin 0 is already handled through log1p in the
// regular path), to
//     q = sum_{k != m} p_k            (small numbers, each known to float32 relative precision)
//     H = -(1 - q) log1p(-q) - sum_{k != m} p_k ln p_k
// in which every term is small and no cancellation is left.  Never taken on sensor data with any noise floor.
#pragma once
#include "common.cuh"

namespace mhb {

constexpr float kToneEntropy = 0.1f;

// log2 of a NORMAL float32 (>= 2^-126): one MUFU.LG2.  `__log2f` carries the denormal fix-up around it (FSETP, FMUL by
// 2^24, FSEL before, a predicated FADD -24 after): four more instructions per bin, two of them on the half-rate ALU pipe --
// 7 % of the W = 500 kernel's instructions in ncu.  Every per-bin argument of the entropy sums is offset by >= 1e-37 before
// the logarithm, so the flush-to-zero form is exact for them (same MUFU result).
__device__ __forceinline__ float log2_normal(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// A kernel that has no PSD row left when it finds such a window (spectral_w1920.cu) stores this quiet-NaN bit pattern
// in the window's entropy cells instead; window_spectral_kernel, run afterwards in redo mode, re-evaluates the marked
// windows with the formula below.  (An all-zero window's entropy is an ordinary NaN, not this pattern.)
constexpr uint32_t kRedoMarkF32 = 0x7fc01920u;
constexpr unsigned long long kRedoMarkF64 = 0x7ff8000000001920ull;

// Lanes l = 0 .. nl-1 of a group cooperate on one PSD row: psd[k] for k = 1 .. nb-1 (bin 0 is `dc`, float64).
// group_argmax(best, arg) and group_sum(v) reduce over the group and leave the result in every lane.
template <class GroupArgMax, class GroupSum>
__device__ __noinline__ double entropy_dominant_bin(const float* psd, int nb, double dc, double tot, int l, int nl,
                                                    GroupArgMax group_argmax, GroupSum group_sum) {
    float best = -1.f;
    int arg = 0x7fffffff;
    for (int k = 1 + l; k < nb; k += nl) {
        const float v = psd[k];
        if (v > best) {
            best = v;
            arg = k;
        }
    }
    group_argmax(best, arg);
    if (dc >= static_cast<double>(best)) arg = 0;
    const float inv = static_cast<float>(1.0 / tot);
    double r = 0.0, h = 0.0;
    for (int k = 1 + l; k < nb; k += nl) {
        if (k == arg) continue;
        const float y = psd[k];
        const float p = y * inv;
        r += static_cast<double>(y);
        if (p > 0.f) h += static_cast<double>(p * logf(p));
    }
    if (l == 0 && arg != 0) {
        const float p = static_cast<float>(dc) * inv;
        r += dc;
        if (p > 0.f) h += static_cast<double>(p * logf(p));
    }
    r = group_sum(r);
    h = group_sum(h);
    const float q = static_cast<float>(r / tot);
    return static_cast<double>(-(1.0f - q) * log1pf(-q)) - h;
}

}  // namespace mhb
