// Kernel 1b (streaming path) -- median / percentiles / IQR of non-overlapping or 50 % overlapping float32 windows with
// NO block-wide barrier: every warp walks its own run of windows.
//
// Replaces rolling_apply(np.median | np.percentile | stats.interquartile_range) (reference
// src/mhealth/generic/stats.py:48-59,158,163 on the driver of util/windows.py:68-91; numba's percentile interpolation,
// numba/np/arraymath.py:1696-1701) for W = g or W = 2 g, S = g (g = gcd(W, S) <= 256): BASELINE configs[1] / [2].
//
// As in window_order_blocks.cu every block of g samples is sorted ONCE and an order statistic of a window is a
// merge-path SELECTION over its one or two sorted blocks.  What changed (round 2, after ncu + tools/ubench/sort_pipes.cu
// showed the batch kernel spending 40 % of its time outside the sorting network: two __syncthreads per batch with the
// selection running on 2 of 8 warps, exposed load latency, and a network bound by the shuffle unit -- one SHFL per 4
// cycles and sub-partition):
//   * a warp owns a CHUNK of consecutive windows of one series and keeps the last five sorted blocks in its own slice of
//     shared memory (slot = block index mod 5): only __syncwarp() separates sort and selection, so all warps of an SM stay
//     busy and CTAs of 8 warps x 5.3 KB fit five to an SM;
//   * a block is sorted by a GROUP of 8 lanes holding 32 elements each (sort_regs.cuh: odd-even merge network in
//     registers, sign-state cross-lane stages) -- a warp sorts 4 consecutive blocks at once with 48 instead of 120
//     shuffles per block;
//   * the selection is cooperative: the 8 lanes of a group run an 8-ary merge-path search (3 rounds of 2 shared loads +
//     ballot instead of 8 dependent binary-search steps by one thread), 4 windows per warp at a time;
//   * ranks and interpolation weights are computed once on the host (same float64 arithmetic as numba's).
// Selected values are samples, so they are exact; the interpolation is the reference's float64 expression.
#include <math_constants.h>

#include "common.cuh"
#include "sort_regs.cuh"

namespace mhb {

namespace {

constexpr int kMaxFeatS = 32;
constexpr int kThreadsOS = 256;
constexpr int kWarpsOS = kThreadsOS / 32;
constexpr int kGL = 8;                   // lanes per sorting group
constexpr int kNG = 32 / kGL;            // blocks sorted at once by a warp
constexpr int kRing = kNG + 1;           // sorted blocks kept per warp

// one selection: elements of rank r and r + 1 of the window, combined as  lower * wl + upper * wu  in float64
// (mode 0), or the window minimum / maximum (modes 1 / 2)
struct Sel {
    int32_t r, mode;
    double wl, wu;
};

struct StreamPlan {
    const float* x;
    int64_t series_stride, nw, chunks_per_series, total_chunks;
    int32_t g, k, n, cw;                 // block length, blocks per window (1 | 2), window length, windows per chunk
    int32_t slot_stride;                 // floats per ring slot
    int32_t one;                         // == 1: opaque multiplier of the FMA-pipe compare-exchange (sort_regs.cuh)
    void* out;
    int64_t o_series, o_window, o_col;
    int32_t n_features;
    int32_t nsel[kMaxFeatS];             // 1, or 2 for IQR (first - second)
    Sel sel[kMaxFeatS][2];
};

__device__ __forceinline__ int pos_of(int e) { return e + (e >> 5); }      // one pad word per 32: conflict-free lane rows

template <typename OutT, int EPL>
__global__ void __launch_bounds__(kThreadsOS, 4) window_order_stream_kernel(const StreamPlan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / kGL, l = lane % kGL;
    float* ring = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * kRing * P.slot_stride;
    const int g = P.g, k = P.k;
    const float inf = CUDART_INF_F;
    const unsigned gshift = static_cast<unsigned>(gi * kGL);

    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kWarpsOS;
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * kWarpsOS + warp; c < P.total_chunks; c += warps_total) {
        const int64_t series = c / P.chunks_per_series;
        const int64_t ci = c - series * P.chunks_per_series;
        const int64_t w0 = ci * P.cw;
        const int64_t left = P.nw - w0;
        const int nwin = left < P.cw ? static_cast<int>(left) : P.cw;
        const int nblk = nwin + k - 1;
        const float* src0 = P.x + series * P.series_stride + w0 * static_cast<int64_t>(g);
        const int64_t obase0 = series * P.o_series + w0 * P.o_window;
        __syncwarp();                                   // the previous chunk's selections are done with the ring

        for (int b0 = 0; b0 < nblk; b0 += kNG) {
            const int blk = b0 + gi;                    // this group's block
            const bool have = blk < nblk;
            // ---- load (any element-to-lane mapping will do: 8 lanes read 32 contiguous bytes) and sort
            float u[EPL];
            const float* src = src0 + static_cast<int64_t>(blk) * g;
            const float* srcl = src + l;
            const int nfull = have ? (g - l + kGL - 1) / kGL : 0;       // elements of this lane: i < nfull
#pragma unroll
            for (int i = 0; i < EPL; ++i) u[i] = i < nfull ? __ldg(srcl + i * kGL) : inf;
            if ((b0 + kNG) * g + lane * 32 < nblk * g && lane * 32 < kNG * g) {      // next iteration's samples -> L2 while this one sorts
                const float* nx = src0 + static_cast<int64_t>(b0 + kNG) * g + lane * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            }
            group_sort_regs_f32<EPL, kGL, true>(u, l, P.one);
            float* slot = ring + (blk % kRing) * P.slot_stride;
            if (have) {
                if constexpr (EPL == 32) {               // pos_of(32 l + i) = 33 l + i: one base, immediate offsets
                    float* row = slot + l * 33;
#pragma unroll
                    for (int i = 0; i < EPL; ++i) row[i] = u[i];
                } else {
#pragma unroll
                    for (int i = 0; i < EPL; ++i) slot[pos_of(l * EPL + i)] = u[i];
                }
            }
            __syncwarp();
            // ---- selection: group gi owns the window that ENDS with its block
            const int wl = blk - (k - 1);
            const bool wact = have && wl >= 0;           // (wl < nwin follows from blk < nblk)
            const float* B = slot;                       // last block of the window
            const float* A = k == 2 ? ring + ((blk + kRing - 1) % kRing) * P.slot_stride : slot;   // first block
            const int gb = k == 2 ? g : 0;               // k == 1: the window is A alone
            for (int j = 0; j < P.n_features; ++j) {
                double acc = 0.0;
                for (int s = 0; s < P.nsel[j]; ++s) {
                    const Sel sl = P.sel[j][s];
                    double pv;
                    if (sl.mode != 0) {
                        float v = 0.f;
                        if (wact) {
                            if (sl.mode == 1) v = (gb && B[0] < A[0]) ? B[0] : A[0];
                            else v = (gb && B[pos_of(g - 1)] > A[pos_of(g - 1)]) ? B[pos_of(g - 1)] : A[pos_of(g - 1)];
                        }
                        pv = static_cast<double>(v);
                    } else {
                        // merge path: i = how many of the r + 1 smallest come from A = number of candidates i in [lo, hi)
                        // with B[r - i] > A[i] (monotone: true ... true, false ... false); 8-ary search, 3 rounds
                        const int r = sl.r;
                        int lo = r + 1 - gb > 0 ? r + 1 - gb : 0;
                        const int hi = r + 1 < g ? r + 1 : g;
                        auto probe = [&](int step) {
                            const int i = lo + (l + 1) * step - 1;
                            bool p = false;
                            if (wact && i < hi) p = B[pos_of(r - i)] > A[pos_of(i)];
                            const unsigned bal = (__ballot_sync(0xffffffffu, p) >> gshift) & 0xffu;
                            lo += __popc(bal) * step;
                        };
                        probe(32);
                        probe(4);
                        probe(1);
                        const int i = lo, jj = r + 1 - lo;
                        float lower = 0.f, upper = 0.f;
                        if (wact) {
                            const float a_last = i > 0 ? A[pos_of(i - 1)] : -inf, b_last = jj > 0 ? B[pos_of(jj - 1)] : -inf;
                            lower = a_last > b_last ? a_last : b_last;
                            const float a_next = i < g ? A[pos_of(i)] : inf, b_next = jj < gb ? B[pos_of(jj)] : inf;
                            upper = a_next < b_next ? a_next : b_next;
                            if (upper == inf) upper = lower;            // r is the top rank
                        }
                        pv = static_cast<double>(lower) * sl.wl + static_cast<double>(upper) * sl.wu;
                    }
                    acc = s == 0 ? pv : acc - pv;
                }
                if (wact && l == 0) store_cell<OutT>(P.out, obase0 + wl * P.o_window + j * P.o_col, acc);
            }
            __syncwarp();                               // ring slots are rewritten by the next iteration
        }
    }
}

}  // namespace

// Returns -100 when this geometry / feature set is not covered here (the caller then uses window_order_blocks.cu).
int32_t window_order_stream_try(const float* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                                const double* h_params, int32_t n_features, const mhb_table* table, void* stream_v) {
    if (n_features <= 0 || n_features > kMaxFeatS) return -100;
    for (int j = 0; j < n_features; ++j)
        if (h_features[j] != MHB_F_MEDIAN && h_features[j] != MHB_F_PERCENTILE && h_features[j] != MHB_F_IQR) return -100;
    const int64_t W = geom->wsize, S = nw == 1 ? geom->wsize : geom->wstep;
    const int64_t g = gcd64(W, S);
    const int64_t k = W / g, hop = S / g;
    if (hop != 1 || k > 2 || g < 16 || g > 256) return -100;
    int64_t p2 = 32;
    while (p2 < g) p2 <<= 1;
    const int epl = static_cast<int>(p2 / kGL);
    StreamPlan P;
    memset(&P, 0, sizeof(P));
    P.x = x;
    P.series_stride = geom->series_stride;
    P.nw = nw;
    P.g = static_cast<int32_t>(g);
    P.k = static_cast<int32_t>(k);
    P.n = static_cast<int32_t>(W);
    // a chunk sorts cw + k - 1 blocks, a multiple of the 4 blocks a warp sorts at once
    P.cw = k == 2 ? 127 : 128;
    P.chunks_per_series = (nw + P.cw - 1) / P.cw;
    P.total_chunks = P.chunks_per_series * geom->n_series;
    int slot = static_cast<int>(p2 + p2 / 32);
    while (slot % 32 != 8) ++slot;                      // the 4 groups' rows start 8 banks apart
    P.slot_stride = slot;
    P.one = 1;
    P.out = table->out;
    P.o_series = table->series_stride;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;
    P.n_features = n_features;
    const int n = static_cast<int>(W);
    auto make_sel = [&](double q) {
        // numba's np.percentile (arraymath.py:1696-1701): rank = 1 + (n - 1) q / 100, lower (1 - m) + upper m
        Sel s;
        memset(&s, 0, sizeof(s));
        if (n == 1 || q == 0.0) {
            s.mode = 1;
        } else if (q == 100.0) {
            s.mode = 2;
        } else {
            const double rank = 1 + (n - 1) * (q / 100.0);
            const double fl = floor(rank);
            const double m = rank - fl;
            s.r = static_cast<int32_t>(fl) - 1;
            s.wl = 1 - m;
            s.wu = m;
        }
        return s;
    };
    for (int j = 0; j < n_features; ++j) {
        P.nsel[j] = 1;
        if (h_features[j] == MHB_F_MEDIAN) {
            Sel s;
            memset(&s, 0, sizeof(s));
            s.r = (n & 1) ? (n >> 1) : (n >> 1) - 1;
            // odd n: the middle element; even n: (a + c) / 2 -- a 0.5 + c 0.5 is the same float64 value (exact halving)
            s.wl = (n & 1) ? 1.0 : 0.5;
            s.wu = (n & 1) ? 0.0 : 0.5;
            P.sel[j][0] = s;
        } else if (h_features[j] == MHB_F_IQR) {
            P.nsel[j] = 2;
            P.sel[j][0] = make_sel(75.0);
            P.sel[j][1] = make_sel(25.0);
        } else {
            P.sel[j][0] = make_sel(h_params ? h_params[j] : 0.0);
        }
    }
    const size_t smem = static_cast<size_t>(kWarpsOS) * kRing * slot * sizeof(float);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    cudaError_t e = cudaSuccess;
#define MHB_GOS_E(OUT, E_)                                                                                       \
    {                                                                                                            \
        auto kern = window_order_stream_kernel<OUT, E_>;                                                         \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));     \
        int per_sm = 0;                                                                                          \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreadsOS, smem); \
        if (e == cudaSuccess) {                                                                                  \
            int64_t ctas = static_cast<int64_t>(kNumSMs) * (per_sm > 0 ? per_sm : 1);                            \
            const int64_t need = (P.total_chunks + kWarpsOS - 1) / kWarpsOS;                                     \
            if (ctas > need) ctas = need;                                                                        \
            kern<<<static_cast<unsigned>(ctas), kThreadsOS, smem, stream>>>(P);                                  \
            e = cudaGetLastError();                                                                              \
        }                                                                                                        \
    }
#define MHB_GOS(OUT)                               \
    switch (epl) {                                 \
        case 4: MHB_GOS_E(OUT, 4) break;           \
        case 8: MHB_GOS_E(OUT, 8) break;           \
        case 16: MHB_GOS_E(OUT, 16) break;         \
        default: MHB_GOS_E(OUT, 32) break;         \
    }
    if (table->out_f32) MHB_GOS(float) else MHB_GOS(double)
#undef MHB_GOS
#undef MHB_GOS_E
    return cuda_status(e, "window_order_stream launch");
}

}  // namespace mhb
