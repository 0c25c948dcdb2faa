// Micro-benchmark (round 2): scalar vs packed (f32x2) FP32 issue rates on sm_100a, measured with the float2 intrinsics
// of crt/sm_100_rt.h (clean SASS: no register shuffling around the packed instructions), plus the other per-sample
// primitives the window kernels use (3-input min/max, funnel shifts, F2F, DFMA) and two realistic bodies: a radix-5
// butterfly on complex values (scalar) against the same butterfly on two transforms at once (packed).
// Prints warp-instructions per clock per SM sub-partition and G lane-results/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o packed_pipes packed_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
#define NCH 8

struct P2 {
    float2 v;
};

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int n_iter) {
    float x[2 * NCH];
    float2 p[NCH];
    double d[NCH];
    int q[NCH];
#pragma unroll
    for (int i = 0; i < 2 * NCH; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        p[i] = make_float2(x[2 * i], x[2 * i + 1]);
        d[i] = threadIdx.x * 0.001 + i;
        q[i] = threadIdx.x + i;
    }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 2 * NCH; ++i) x[i] = fmaf(x[i], a, b);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) p[i] = __ffma2_rn(p[i], a2, b2);
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 2 * NCH; ++i) x[i] = x[i] + a;
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) p[i] = __fadd2_rn(p[i], a2);
            } else if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < 2 * NCH; ++i) x[i] = x[i] * a;
            } else if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) p[i] = __fmul2_rn(p[i], a2);
            } else if (MODE == 6) {   // half FFMA2, half LOP3 (does the packed op free issue slots for the ALU pipe?)
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    p[i] = __ffma2_rn(p[i], a2, b2);
                    q[i] = (q[i] ^ (it + i)) + 1;
                }
            } else if (MODE == 7) {   // scalar FFMA x2 + the same integer work (same flops as mode 6)
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    x[2 * i] = fmaf(x[2 * i], a, b);
                    x[2 * i + 1] = fmaf(x[2 * i + 1], a, b);
                    q[i] = (q[i] ^ (it + i)) + 1;
                }
            } else if (MODE == 8) {   // 3-input min (FMNMX3)
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    float r;
                    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x[i]), "f"(x[NCH + i]), "f"(a));
                    x[i] = r + b;      // keep a dependency that is not foldable
                }
            } else if (MODE == 9) {   // 2-input min x2 (same work as mode 8)
#pragma unroll
                for (int i = 0; i < NCH; ++i) x[i] = fminf(fminf(x[i], x[NCH + i]), a) + b;
            } else if (MODE == 10) {  // DFMA
#pragma unroll
                for (int i = 0; i < NCH; ++i) d[i] = fma(d[i], (double)a, (double)b);
            } else if (MODE == 11) {  // F2F + DADD
#pragma unroll
                for (int i = 0; i < NCH; ++i) d[i] += (double)x[i];
            } else if (MODE == 12) {  // FADD2 + FFMA2 + FMUL2 mix as in a packed butterfly
#pragma unroll
                for (int i = 0; i < NCH; i += 2) {
                    const float2 s = __fadd2_rn(p[i], p[i + 1]);
                    const float2 t = __fmul2_rn(p[i], a2);
                    p[i] = __ffma2_rn(s, a2, t);
                    p[i + 1] = __fadd2_rn(t, b2);
                }
            } else if (MODE == 13) {  // the same arithmetic, scalar
#pragma unroll
                for (int i = 0; i < 2 * NCH; i += 2) {
                    const float s = x[i] + x[i + 1];
                    const float t = x[i] * a;
                    x[i] = fmaf(s, a, t);
                    x[i + 1] = t + b;
                }
            } else if (MODE == 14) {  // funnel shift mask building + FADD (zero-crossing mask idea)
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const float dd = a - x[i];
                    q[i] = __funnelshift_l(__float_as_int(dd), q[i], 1);
                    x[i] = dd;
                }
            } else if (MODE == 15) {  // MUFU.LG2
#pragma unroll
                for (int i = 0; i < NCH; ++i) x[i] = __log2f(x[i]) + a;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NCH; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += p[i].x + p[i].y + (float)d[i] + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double instr_per_iter, double results_per_iter, int ctas_per_sm) {
    const int nb = 148 * ctas_per_sm;
    float* out;
    cudaMalloc(&out, nb * 256 * 4);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, 64);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, ITERS);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps = (double)nb * 8;
    const double winstr = warps * ITERS * 4 * instr_per_iter;
    const double clk = 1.965e9;
    printf("%-34s ctas/sm=%d  ms=%.3f  warp-instr/clk/SMSP=%.3f  Glane-results/s=%.0f  err=%s\n", name, ctas_per_sm, best,
           winstr / (best * 1e-3) / clk / (148 * 4), warps * 32 * ITERS * 4 * results_per_iter / (best * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int c = 1; c <= 4; c *= 2) {
        run<0>("FFMA x16", 16, 16, c);
        run<1>("FFMA2 x8", 8, 16, c);
        run<2>("FADD x16", 16, 16, c);
        run<3>("FADD2 x8", 8, 16, c);
        run<4>("FMUL x16", 16, 16, c);
        run<5>("FMUL2 x8", 8, 16, c);
        run<6>("FFMA2 x8 + (LOP3,IADD) x8", 24, 16, c);
        run<7>("FFMA x16 + (LOP3,IADD) x8", 32, 16, c);
        run<8>("FMNMX3 + FADD x8", 16, 8, c);
        run<9>("FMNMX x2 + FADD x8", 24, 8, c);
        run<10>("DFMA x8", 8, 8, c);
        run<11>("F2F+DADD x8", 16, 8, c);
        run<12>("packed mix (2 FADD2,FMUL2,FFMA2) x4", 16, 32, c);
        run<13>("scalar mix (same flops) x8", 32, 32, c);
        run<14>("FADD + SHF x8", 16, 8, c);
        run<15>("MUFU.LG2 + FADD x8", 16, 8, c);
    }
    return 0;
}
