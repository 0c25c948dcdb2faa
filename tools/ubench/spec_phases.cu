// Phase timing of spectral_fast_kernel (development aid): builds the kernel with MHB_PHASE_TIMING and prints the
// cycles each warp spends per phase.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMHB_PHASE_TIMING
#include "../../pymhealth_b200/csrc/abi.cu"
#include "../../pymhealth_b200/csrc/spectral_fast.cu"
#include <vector>
#include <cstdio>
int main() {
    const int64_t n = 30240000 / 4, ns = 24;
    float* x;
    cudaMalloc(&x, sizeof(float) * n * ns);
    std::vector<float> h(n);
    for (int64_t i = 0; i < n; ++i) h[i] = 1.0f + 0.3f * sinf(0.2f * i) + 0.05f * ((i * 2654435761u % 1000) / 1000.0f - 0.5f);
    for (int s = 0; s < ns; ++s) cudaMemcpy(x + s * n, h.data(), sizeof(float) * n, cudaMemcpyHostToDevice);
    mhb_windows g = {ns, n, n, 500, 250};
    const int64_t nw = 1 + (n - 500) / 250;
    float* out;
    cudaMalloc(&out, sizeof(float) * ns * nw * 6);
    int32_t cols[6] = {0, 1, 1, 2, 3, 5}, lo[6] = {0, 5, 30, 5, 3, 0}, hi[6] = {251, 31, 81, 31, 120, 251};
    for (int rep = 0; rep < 2; ++rep) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(mhb::g_phase_cycles, z, sizeof(z));
        cudaMemcpyToSymbol(mhb::g_reduce_cycles, z, sizeof(unsigned long long) * 16);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int st = mhb::spectral_fast_try(x, &g, nw, 0.1, cols, lo, hi, 6, out, 1, nw * 6, 6, 1, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long c[8][8];
        cudaMemcpyFromSymbol(c, mhb::g_phase_cycles, sizeof(c));
        printf("status %d  %.3f ms  %.3f Gwin/s  (%s)\n", st, ms, ns * nw / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        const char* names[5] = {"tile wait", "passA/reduce", "wait B1", "pass B", "wait B2"};
        const double nb = double(ns) * ((nw + 15) / 16);
        for (int ph = 0; ph < 5; ++ph) {
            printf("%-14s", names[ph]);
            for (int wv = 0; wv < 8; ++wv) printf(" %8.0f", c[ph][wv] / nb);
            printf("   cycles/batch per warp\n");
        }
        unsigned long long r[16];
        cudaMemcpyFromSymbol(r, mhb::g_reduce_cycles, sizeof(r));
        printf("reduce sub-phases (cycles/batch per warp): records %.0f |", r[0] / nb / 3);
        for (int c2 = 0; c2 < 6; ++c2) printf(" col%d %.0f", c2, r[1 + c2] / nb);
        printf("\n");
    }
    return 0;
}
