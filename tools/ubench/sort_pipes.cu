// Micro-benchmark (round 2): what bounds the register-resident block sort of kernel 1b on sm_100a?
// Runs, on register data only (no memory traffic), (0) the generic bitonic network of sort_regs.cuh, (1) the sign-state
// float network, and the instruction classes they are made of in isolation: (2) SHFL.BFLY, (3) FMNMX, (4) FMUL,
// (5) FMUL + SHFL + FMNMX interleaved as a cross-lane stage, (6) in-register compare-exchanges (2 FMNMX).
// Prints cycles per warp-instruction per SM sub-partition (8 warps resident per sub-partition, as in the kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pymhealth_b200/csrc -o sort_pipes tools/ubench/sort_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sort_regs.cuh"

using namespace mhb;

// grouped sort: GL lanes x EPL elements per 256-element block, 32 / GL blocks per warp
template <int EPL, int GL, bool MIX>
__global__ void __launch_bounds__(256, 4) kg(float* out, int n_iter, float s) {
    float v[EPL];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = __sinf(threadIdx.x * 0.37f + i * 1.1f + blockIdx.x);
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
        group_sort_regs_f32<EPL, GL, MIX>(v, lane % GL, static_cast<int>(s));
#pragma unroll
        for (int i = 0; i < EPL; ++i) v[i] = v[i] * s + (float)((lane * 7 + i * 13 + it) & 31);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc += v[i] * (i + 1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int EPL, int GL, bool MIX = false>
void run_g(const char* name, float* out, int n_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = 148 * 4;
    kg<EPL, GL, MIX><<<ctas, 256>>>(out, 16, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kg<EPL, GL, MIX><<<ctas, 256>>>(out, n_iter, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int mhz = 0;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * mhz * 1e3;
    const double blocks_per_iter = 32.0 / GL * (EPL * GL / 256.0);      // in units of 256-element blocks
    printf("%-44s %8.3f ms  %8.1f cycles / 256 elements / sub-partition\n", name, ms, cycles / n_iter / 8.0 / blocks_per_iter);
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float* out, int n_iter, float s) {
    float v[8];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __sinf(threadIdx.x * 0.37f + i * 1.1f + blockIdx.x);
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
        if (MODE == 0) {
            warp_sort_regs<float, 8>(v, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] * s + (float)((lane * 7 + i * 13 + it) & 31);   // unsort (2 ops / element)
        } else if (MODE == 1) {
            warp_sort_regs_f32<8>(v, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] * s + (float)((lane * 7 + i * 13 + it) & 31);
        } else if (MODE == 2) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (u & 15));
        } else if (MODE == 3) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fminf(v[i], v[(i + 1) & 7] + 0.f * s);
        } else if (MODE == 4) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __fmul_rn(v[i], s);
        } else if (MODE == 5) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fminf(__fmul_rn(v[i], s), -__shfl_xor_sync(0xffffffffu, __fmul_rn(v[i], s), 1 + (u & 15)));
        } else if (MODE == 7) {
            // compare-exchange with the maximum formed on the FMA pipe: bits(max) = bits(a) + bits(c) - bits(min), two IMADs
            // whose multipliers (+1, -1) are runtime values (a literal folds into one IADD3 -- ALU pipe again)
            const int one = __float_as_int(s) >> 23 == 127 ? 1 : 2, mone = -one;
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = 1 << (u % 3);
                    if ((i ^ j) > i) {
                        const float a = v[i], c = v[i ^ j];
                        const float mn = fminf(a, c);
                        int t, m;
                        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(__float_as_int(a)), "r"(one), "r"(__float_as_int(c)));
                        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(m) : "r"(__float_as_int(mn)), "r"(mone), "r"(t));
                        v[i] = mn;
                        v[i ^ j] = __int_as_float(m);
                    }
                }
        } else if (MODE == 6) {
            // distances 1, 2, 4 in turn (the same pair twice in a row would be simplified away by the compiler)
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = 1 << (u % 3);
                    if ((i ^ j) > i) {
                        const float a = v[i], c = v[i ^ j];
                        v[i] = fminf(a, c);
                        v[i ^ j] = fmaxf(a, c);
                    }
                }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double instr_per_iter, float* out, int n_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = 148 * 4;
    k<MODE><<<ctas, 256>>>(out, 16, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, n_iter, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int mhz = 0;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * mhz * 1e3;
    const double per_iter_smsp = cycles / n_iter / 8.0;          // 8 warps per sub-partition run the loop concurrently
    printf("%-44s %8.3f ms  %8.1f cycles / iteration / warp-slot", name, ms, per_iter_smsp);
    if (instr_per_iter > 0) printf("  (%.2f cycles per warp-instruction)", per_iter_smsp / instr_per_iter);
    printf("\n");
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    const int n = 4096;
    run<0>("generic bitonic sort (256 elements / warp)", 0, out, n);
    run<1>("sign-state sort (256 elements / warp)", 0, out, n);
    run_g<32, 8>("group sort 8 lanes x 32 (4 blocks / warp)", out, n / 4);
    run_g<32, 8, true>("group sort 8 lanes x 32, 1/3 FMA-pipe maxima", out, n / 4);
    run_g<16, 16>("group sort 16 lanes x 16 (2 blocks / warp)", out, n / 2);
    run_g<8, 32>("group sort 32 lanes x 8", out, n);
    run_g<8, 8>("group sort 8 lanes x 8 (64-element blocks)", out, n);
    run_g<2, 32>("group sort 32 lanes x 2 (64-element blocks)", out, n);
    run<2>("SHFL.BFLY x128", 128, out, n);
    run<3>("FMNMX(+FFMA) x128", 256, out, n);
    run<4>("FMUL x128", 128, out, n);
    run<5>("FMUL+FMUL+SHFL+FMNMX x128", 512, out, n);
    run<6>("in-register compare-exchange (2 FMNMX) x64", 128, out, n);
    run<7>("compare-exchange FMNMX + 2 IMAD x64", 192, out, n);
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
