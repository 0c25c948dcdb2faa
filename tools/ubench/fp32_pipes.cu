// Micro-benchmark: issue/throughput of scalar vs packed (f32x2) FP32 ops and FP64 ops on sm_100a.
// Prints lane-ops per clock per SM for each variant.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, long long* cyc) {
    float x[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) x[i] = threadIdx.x * 0.001f + i;
    double d[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) d[i] = threadIdx.x * 0.001 + i;
    unsigned long long ab, bb;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(a), "f"(a));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    unsigned long long p[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {          // scalar FFMA, 16 independent
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {   // packed FFMA2, 8 independent (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ab), "l"(bb));
        } else if (MODE == 2) {   // scalar FADD
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) x[i] = x[i] + a;
        } else if (MODE == 3) {   // packed FADD2
#pragma unroll
            for (int i = 0; i < NACC; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ab));
        } else if (MODE == 4) {   // scalar FMUL
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) x[i] = x[i] * a;
        } else if (MODE == 5) {   // packed FMUL2
#pragma unroll
            for (int i = 0; i < NACC; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ab));
        } else if (MODE == 6) {   // DFMA
#pragma unroll
            for (int i = 0; i < NACC; ++i) d[i] = fma(d[i], (double)a, (double)b);
        } else if (MODE == 7) {   // DADD
#pragma unroll
            for (int i = 0; i < NACC; ++i) d[i] = d[i] + (double)a;
        } else if (MODE == 8) {   // mix: 8 FFMA + 8 IADD-ish (LOP3) per iteration
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                x[i] = fmaf(x[i], a, b);
                x[NACC + i] = __int_as_float(__float_as_int(x[NACC + i]) ^ (it + i));
            }
        } else if (MODE == 9) {   // mix: 8 FFMA2 + 8 LOP3 (same flops as 16 FFMA)
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ab), "l"(bb));
                x[i] = __int_as_float(__float_as_int(x[i]) ^ (it + i));
            }
        } else if (MODE == 10) {  // F2F f32->f64 convert + DADD
#pragma unroll
            for (int i = 0; i < NACC; ++i) d[i] += (double)x[i];
        } else if (MODE == 11) {  // MUFU lg2
#pragma unroll
            for (int i = 0; i < NACC; ++i) x[i] = __log2f(x[i]);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += lo + hi + (float)d[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_thread_iter, int ctas_per_sm) {
    int nb = 148 * ctas_per_sm;
    float* out;
    long long* cyc;
    cudaMalloc(&out, nb * 256 * 4);
    cudaMalloc(&cyc, nb * 8);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8];
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nb; ++i) avg += h[i];
    avg /= nb;
    double ops_per_sm = (double)ctas_per_sm * 256 * ITERS * lane_ops_per_thread_iter;
    printf("%-28s ctas/sm=%d  cycles=%.0f  lane-instr/clk/SM=%.1f  ms=%.3f  err=%s\n", name, ctas_per_sm, avg,
           ops_per_sm / avg, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int c = 2; c <= 4; c += 2) {
        run<0>("FFMA x16", 16, c);
        run<1>("FFMA2 x8 (=16 fma)", 8, c);
        run<2>("FADD x16", 16, c);
        run<3>("FADD2 x8", 8, c);
        run<4>("FMUL x16", 16, c);
        run<5>("FMUL2 x8", 8, c);
        run<6>("DFMA x8", 8, c);
        run<7>("DADD x8", 8, c);
        run<8>("FFMA x8 + LOP3 x8", 16, c);
        run<9>("FFMA2 x8 + LOP3 x8", 16, c);
        run<10>("F2F+DADD x8 (2 instr each)", 16, c);
        run<11>("MUFU.LG2 x8", 8, c);
    }
    return 0;
}
