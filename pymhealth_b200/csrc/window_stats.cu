// Kernel 1a -- streaming sliding-window statistics (sm_100a).
//
// Replaces the numba loop of rolling_apply (reference src/mhealth/util/windows.py:68-91) for the
// reducers that are sums / extrema over the window: mean, var, std, min, max, drange, skewness,
// kurtosis(+excess), coeff_var (generic/stats.py:12-163), zero_crossing_count, line_length,
// hjorth_activity (generic/timedom.py:34-95) -- ALL of them in one pass in which every sample
// is read from HBM once, however much the windows overlap.
//
// Decomposition.  g = gcd(W, S) (or a divisor of it): a window is k = W/g consecutive "blocks"
// and successive windows start hop = S/g blocks apart, so per-block partial sums are shared by
// all windows that cover the block.  One CTA walks a chunk of one series stage by stage:
//   TMA   : 1-D bulk copies (cp.async.bulk + mbarrier, NS-deep ring) bring TB blocks (+1 sample)
//           of the series into shared memory; the data never goes through registers twice.
//   phase1: every thread scans one "cell" of m consecutive samples (m | g) sequentially and
//           produces its partial: sum d, d^2, d^3, d^4 (d = x - pivot, float64), min, max,
//           line length and zero crossings inside the cell and across its right edge.
//   phase2: cells -> block partials, written to a ring of recent blocks.
//   phase3: one thread per finished window combines k ring entries (direct sums for small k,
//           differences of a running prefix for large k), turns the shifted power sums into
//           central moments, and stores the requested feature columns.  It is deferred until
//           about half a CTA's worth of windows is pending so it is not a one-warp bubble.
// Parity notes (SURVEY section 8c gotchas): population variance; kurtosis / skewness return 0 for a
// constant window; zero is "not positive" for crossings; no partial tail window.
#include <cstdlib>
#include <math_constants.h>

#include "common.cuh"

namespace mhb {

// window_segments.cu: warp-per-window evaluation of uniform windows (fallback for degenerate block decompositions)
template <typename InT>
int32_t window_stats_direct(const InT* x, const mhb_windows* geom, int64_t nw, const int32_t* h_features,
                            int32_t n_features, double zc_threshold, const mhb_table* table, void* stream_v);

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFeat = 32;
constexpr int kDirectK = 8;      // windows of <= kDirectK blocks are summed directly
constexpr int kNPre = 7;         // S1..S4, LL, ZC get a running prefix when k > kDirectK; the 7th counts non-finite blocks

struct StatsPlan {
    const void* x;
    const void* y;               // magnitude mode (accelerometer.py:198-225 fused into the staging copy): the series is
    const void* z;               // sqrt(x^2 + y^2 + z^2) of three arrays of the same geometry; null otherwise
    int64_t series_len, series_stride, total_elems;
    int64_t nw, win_per_chunk, win_per_segment;
    int32_t chunks_per_series;
    int32_t W, S, g, k, hop, m, cpb, TB, RB, NS;
    int32_t flush;               // finalize once this many windows are pending
    int32_t stage_elems;
    int32_t use_tma;
    int32_t mag_staged;          // magnitude mode through the MAG = true kernels (float32): three arrays per stage slot
    double th;
    double inv_n;
    void* out;
    int64_t o_series, o_window, o_col;
    int32_t n_features;
    int32_t feat[kMaxFeat];
};

// Shared-memory partial record, structure-of-arrays.  The same layout serves the per-cell scratch
// (n = cells per stage) and the block ring (n = RB).
template <typename InT, bool M4, bool TD>
struct Partials {
    double* s1;
    double* s2;
    double* s3;
    double* s4;
    InT* mn;
    InT* mx;
    float* ll;     // line length inside + across the right edge
    float* llb;    // ... line length WITHOUT the right-edge term (the last block of a window contributes this one: a
                   //     subtraction "total - edge" would turn a NaN sample just past the window into NaN - NaN)
    float* zc;     // zero crossings inside + across the right edge (exact small integers)
    float* zcb;

    __device__ unsigned char* carve(unsigned char* p, int n) {
        s1 = reinterpret_cast<double*>(p); p += sizeof(double) * n;
        s2 = reinterpret_cast<double*>(p); p += sizeof(double) * n;
        s3 = s4 = nullptr;
        if (M4) {
            s3 = reinterpret_cast<double*>(p); p += sizeof(double) * n;
            s4 = reinterpret_cast<double*>(p); p += sizeof(double) * n;
        }
        mn = reinterpret_cast<InT*>(p); p += sizeof(InT) * n;
        mx = reinterpret_cast<InT*>(p); p += sizeof(InT) * n;
        ll = llb = zc = zcb = nullptr;
        if (TD) {
            ll = reinterpret_cast<float*>(p); p += sizeof(float) * n;
            llb = reinterpret_cast<float*>(p); p += sizeof(float) * n;
            zc = reinterpret_cast<float*>(p); p += sizeof(float) * n;
            zcb = reinterpret_cast<float*>(p); p += sizeof(float) * n;
        }
        return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    }
};

static size_t partial_bytes(size_t in_size, bool m4, bool td, int n) {
    size_t b = (sizeof(double) * (m4 ? 4 : 2) + in_size * 2 + (td ? sizeof(float) * 4 : 0)) * n;
    return ((b + 15) & ~size_t(15)) + 16;
}

// min / max that PROPAGATE NaN, as np.min / np.max do (fminf / fmaxf would silently drop it): FMNMX.NAN for float32
template <typename T>
__device__ __forceinline__ T tmin(T a, T b);
template <>
__device__ __forceinline__ float tmin<float>(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <>
__device__ __forceinline__ double tmin<double>(double a, double b) { return (a != a || b != b) ? a + b : fmin(a, b); }
template <typename T>
__device__ __forceinline__ T tmax(T a, T b);
template <>
__device__ __forceinline__ float tmax<float>(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <>
__device__ __forceinline__ double tmax<double>(double a, double b) { return (a != a || b != b) ? a + b : fmax(a, b); }

template <typename InT>
struct CellAcc {
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    float ll = 0.f, zc = 0.f;
    InT mn, mx;
    bool nan = false;      // float64 cells: a NaN sample was met (DMNMX cannot propagate it; float32 uses FMNMX.NAN)
};

// per-sample extrema: float32 propagates NaN in the instruction itself; float64 takes the plain DMNMX and one
// predicate-accumulating compare per sample, and the cell's extrema are set to NaN at its end
__device__ __forceinline__ void accum_extrema(CellAcc<float>& a, float v) {
    a.mn = tmin<float>(a.mn, v);
    a.mx = tmax<float>(a.mx, v);
}
__device__ __forceinline__ void accum_extrema(CellAcc<double>& a, double v) {
    a.mn = fmin(a.mn, v);
    a.mx = fmax(a.mx, v);
    a.nan = a.nan || (v != v);
}

template <typename InT, bool M4>
__device__ __forceinline__ void accum_moments(CellAcc<InT>& a, InT v, double c) {
    const double d = static_cast<double>(v) - c;
    a.s1 += d;
    a.s2 = fma(d, d, a.s2);
    if (M4) {
        const double d2 = d * d;
        a.s3 = fma(d2, d, a.s3);
        a.s4 = fma(d2, d2, a.s4);
    }
    accum_extrema(a, v);
}

// time-domain pair terms between neighbours (prev, v); pos flags are 1.0f / 0.0f
template <typename InT>
__device__ __forceinline__ void accum_pair(CellAcc<InT>& a, InT v, InT prev, float pos, float pos_prev) {
    a.ll += fabsf(static_cast<float>(v - prev));
    const float dp = pos - pos_prev;
    a.zc = fmaf(dp, dp, a.zc);
}

template <typename InT>
__device__ __forceinline__ InT round_down_threshold(double th);
template <>
__device__ __forceinline__ float round_down_threshold<float>(double th) {
    // x > th (x float, th double)  <=>  x > largest float <= th
    return __double2float_rd(th);
}
template <>
__device__ __forceinline__ double round_down_threshold<double>(double th) {
    return th;
}

// sqrt(x**2 + y**2 + z**2) in the input type with every operation rounded, exactly as accel.cu's standalone kernel
template <typename InT>
__device__ __forceinline__ InT magnitude3(InT x, InT y, InT z);
template <>
__device__ __forceinline__ float magnitude3<float>(float x, float y, float z) {
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
template <>
__device__ __forceinline__ double magnitude3<double>(double x, double y, double z) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// Chunk / stage geometry.  Everything a stage needs is 32-bit arithmetic relative to the chunk start (a chunk is
// at most 64 stages of <= 24 KB); the 64-bit quantities are folded once per CTA.
template <typename InT>
struct ChunkDesc {
    static constexpr int A = 16 / sizeof(InT);
    int64_t goff0;      // global element offset of the chunk's first sample
    int32_t lead0;      // goff0 mod A
    int32_t avail;      // samples of the series from the chunk start on (clamped)
    int32_t room;       // elements of the whole buffer from the aligned chunk start on (clamped)
    __device__ __forceinline__ void set(const StatsPlan& P, int64_t series_base, int64_t chunk_s0) {
        goff0 = series_base + chunk_s0;
        lead0 = static_cast<int32_t>(goff0 & (A - 1));
        const int64_t av = P.series_len - chunk_s0;
        avail = av > 0x7fffffff ? 0x7fffffff : static_cast<int32_t>(av);
        const int64_t rm = P.total_elems - (goff0 - lead0);
        room = rm > 0x7fffffff ? 0x7fffffff : static_cast<int32_t>(rm);
    }
};

template <typename InT>
struct StageDesc {
    static constexpr int A = 16 / sizeof(InT);
    int32_t rel;        // first sample of the stage, relative to the chunk start
    int32_t lead;       // (goff0 + rel) mod A: offset of the stage's first sample inside the staged copy
    int32_t nblk;       // blocks in this stage
    int32_t cnt;        // samples belonging to blocks
    int32_t has_next;   // the sample after the stage exists in the series
    int32_t n_load;     // elements copied by TMA (multiple of 16 bytes)
    int32_t tma;        // stage is loaded by TMA (else cooperative guarded copy)

    __device__ __forceinline__ void set(const StatsPlan& P, const ChunkDesc<InT>& ck, int32_t st, int32_t stage_samples,
                                        int32_t blocks_left, int32_t TB) {
        nblk = blocks_left < TB ? blocks_left : TB;
        cnt = nblk * P.g;
        rel = st * stage_samples;
        has_next = (rel + cnt < ck.avail) ? 1 : 0;
        lead = (ck.lead0 + rel) & (A - 1);
        n_load = (lead + cnt + has_next + A - 1) & ~(A - 1);
        tma = (P.use_tma && (ck.lead0 + rel - lead) + n_load <= ck.room) ? 1 : 0;
    }
};

__device__ __forceinline__ int wrap(int i, int n) { return i >= n ? i - n : i; }

// ---------------------------------------------------------------------------------------------
// GEO = 1: the hot geometry (cpb = 10, k = 2, hop = 1, TB = 25 -- W = 500 / S = 250 with 25-sample cells) with every
// loop bound known at compile time; GEO = 2: the same for W = 1920 / S = 64 (BASELINE configs[3]: cpb = 4 cells of 16, k = 30,
// hop = 1, TB = 64); GEO = 0: bounds from the plan.
// MAG = true: magnitude mode with the three axes staged by TMA side by side (x | y | z, stage_elems apart) and combined
// as phase 1 reads them; MAG = false with P.y set: the guarded-copy path combines them while copying.
template <typename InT, typename OutT, bool M4, bool TD, int MCELL /*0 = runtime m; -8 = 2 x float4*/, int GEO, bool MAG = false>
__global__ void __launch_bounds__(kThreads, 3) window_stats_kernel(const StatsPlan P) {
    static_assert(!MAG || MCELL >= 0, "magnitude mode uses the scalar cells");
    constexpr int SB = MAG ? 3 : 1;          // arrays per stage slot
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int m = MCELL > 0 ? MCELL : (MCELL < 0 ? -MCELL : P.m);
    const int g_cpb = GEO == 1 ? 10 : GEO == 2 ? 4 : P.cpb, g_k = GEO == 1 ? 2 : GEO == 2 ? 30 : P.k, g_hop = GEO ? 1 : P.hop,
              g_TB = GEO == 1 ? 25 : GEO == 2 ? 64 : P.TB;
    using Part = Partials<InT, M4, TD>;

    // ---- carve shared memory
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // NS barriers (<= 8)
    unsigned char* ptr = smem_raw + 128;
    InT* stage_buf = reinterpret_cast<InT*>(ptr);
    ptr += static_cast<size_t>(P.NS) * SB * P.stage_elems * sizeof(InT);
    const int ncell_max = g_TB * g_cpb;
    Part cell, ring;
    if (g_cpb > 1) ptr = cell.carve(ptr, ncell_max);
    ptr = ring.carve(ptr, P.RB);
    double* pre = reinterpret_cast<double*>(ptr);                           // [kNPre][RB + 1] (k > kDirectK)
    const int R1 = P.RB + 1;
    __shared__ double carry[kNPre];

    // ---- which chunk
    const int64_t series = blockIdx.x / P.chunks_per_series;
    const int32_t chunk = blockIdx.x - static_cast<int32_t>(series) * P.chunks_per_series;
    const int64_t w0 = static_cast<int64_t>(chunk) * P.win_per_chunk;
    const int64_t w1 = min(P.nw, w0 + P.win_per_chunk);
    const int32_t nwin = static_cast<int32_t>(w1 - w0);
    const int32_t n_blocks = (nwin - 1) * g_hop + g_k;
    const int32_t n_stages = (n_blocks + g_TB - 1) / g_TB;
    const bool use_prefix = g_k > kDirectK;
    const InT* xg = reinterpret_cast<const InT*>(P.x);
    const int64_t series_base = series * P.series_stride;
    const int64_t chunk_s0 = w0 * P.S;                    // first sample of the chunk (series-relative)
    const int32_t stage_samples = g_TB * P.g;
    ChunkDesc<InT> ck;
    ck.set(P, series_base, chunk_s0);
    const InT* xck = xg + ck.goff0;                       // chunk start
    const InT* yck = MAG ? reinterpret_cast<const InT*>(P.y) + ck.goff0 : nullptr;
    const InT* zck = MAG ? reinterpret_cast<const InT*>(P.z) + ck.goff0 : nullptr;
    // thread 0: one expect_tx for the whole slot, one bulk copy per array
    auto issue_stage = [&](int sl, const StageDesc<InT>& sd) {
        InT* dst = stage_buf + static_cast<size_t>(sl) * SB * P.stage_elems;
        const uint32_t bytes = sd.n_load * sizeof(InT);
        mbar_arrive_expect_tx(&full[sl], SB * bytes);
        bulk_g2s(dst, xck + (sd.rel - sd.lead), bytes, &full[sl]);
        if (MAG) {
            bulk_g2s(dst + P.stage_elems, yck + (sd.rel - sd.lead), bytes, &full[sl]);
            bulk_g2s(dst + 2 * P.stage_elems, zck + (sd.rel - sd.lead), bytes, &full[sl]);
        }
    };
    // sample i of a staged array: the magnitude of the three staged axes in magnitude mode
    auto ld = [&](const InT* q, int i) -> InT {
        if (MAG) return magnitude3<InT>(q[i], q[i + P.stage_elems], q[i + 2 * P.stage_elems]);
        return q[i];
    };

    if (tid == 0) {
        for (int i = 0; i < P.NS; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    if (tid < kNPre) carry[tid] = 0.0;
    if (use_prefix && tid < kNPre) pre[tid * R1] = 0.0;                    // prefix before block 0
    __syncthreads();

    if (tid == 0) {
        for (int st = 0; st < P.NS && st < n_stages; ++st) {
            StageDesc<InT> d;
            d.set(P, ck, st, stage_samples, n_blocks - st * g_TB, g_TB);
            if (d.tma) issue_stage(st, d);
        }
    }

    const InT th = round_down_threshold<InT>(P.th > 0.0 ? P.th : 0.0);
    double c = 0.0;             // pivot of the shifted power sums: first sample of the chunk
    int32_t blocks_done = 0;    // blocks whose partials are in the ring
    int32_t ring_head = 0;      // ring slot of block `blocks_done`
    int32_t emitted = 0;        // windows already written
    int32_t ring_emit = 0;      // ring slot of the first block of window `emitted`
    int32_t pre_emit = 0;       // prefix slot of that block
    int slot = 0;
    uint32_t parity = 0;

    for (int st = 0; st < n_stages; ++st) {
        StageDesc<InT> d;
        d.set(P, ck, st, stage_samples, n_blocks - st * g_TB, g_TB);
        InT* buf = stage_buf + static_cast<size_t>(slot) * SB * P.stage_elems;
        if (d.tma) {
            mbar_wait(&full[slot], parity);
        } else {
            // guarded cooperative copy: unaligned base pointer or the last few samples of the buffer
            const int n = d.cnt + d.has_next;
            if (MAG) {
                const InT* ys = yck + d.rel;
                const InT* zs = zck + d.rel;
                for (int i = tid; i < n; i += kThreads) {
                    buf[d.lead + i] = xck[d.rel + i];
                    buf[d.lead + i + P.stage_elems] = ys[i];
                    buf[d.lead + i + 2 * P.stage_elems] = zs[i];
                }
            } else if (P.y) {
                // magnitude mode: the three axes are read straight from global memory (coalesced, four independent
                // loads per axis in flight per thread) and only their magnitude is staged
                const InT* yg = reinterpret_cast<const InT*>(P.y) + ck.goff0 + d.rel;
                const InT* zg = reinterpret_cast<const InT*>(P.z) + ck.goff0 + d.rel;
                const InT* xs = xck + d.rel;
                InT* dstm = buf + d.lead;
                int i = tid;
                for (; i + 3 * kThreads < n; i += 4 * kThreads) {
                    InT vx[4], vy[4], vz[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        vx[u] = __ldg(xs + i + u * kThreads);
                        vy[u] = __ldg(yg + i + u * kThreads);
                        vz[u] = __ldg(zg + i + u * kThreads);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) dstm[i + u * kThreads] = magnitude3<InT>(vx[u], vy[u], vz[u]);
                }
                for (; i < n; i += kThreads) dstm[i] = magnitude3<InT>(xs[i], yg[i], zg[i]);
            } else {
                for (int i = tid; i < n; i += kThreads) buf[d.lead + i] = xck[d.rel + i];
            }
            __syncthreads();
        }
        const InT* s = buf + d.lead;
        if (st == 0) {
            // pivot of the shifted power sums: the first sample of the chunk's SEGMENT (a fixed number of windows counted
            // from the start of the series; a chunk is a power-of-two fraction of one), so that the sums -- and with
            // them every bit of the table -- do not depend on how a call is cut into chunks, i.e. on how many series
            // share the launch.  Any finite pivot is exact; a NaN / inf sample there must not poison the chunk.
            const int64_t pidx = series_base + (w0 / P.win_per_segment) * P.win_per_segment * P.S;
            if (P.y) c = static_cast<double>(magnitude3<InT>(xg[pidx], reinterpret_cast<const InT*>(P.y)[pidx],
                                                             reinterpret_cast<const InT*>(P.z)[pidx]));
            else c = static_cast<double>(xg[pidx]);
            if (!isfinite(c)) c = 0.0;
        }

        // ---------------- phase 1: one cell per thread
        const int ncell = d.nblk * g_cpb;
        for (int ce = tid; ce < ncell; ce += kThreads) {
            const InT* p = s + ce * m;
            CellAcc<InT> a;
            InT prev = ld(p, 0);
            a.mn = prev;
            a.mx = prev;
            float pos_prev = (TD && prev > th) ? 1.f : 0.f;
            accum_moments<InT, M4>(a, prev, c);
            if (MCELL < 0) {
                // power-of-two blocks: the cell is -MCELL/4 aligned 128-bit shared loads
                InT vv[MCELL < 0 ? -MCELL : 1];
#pragma unroll
                for (int i = 0; i < (MCELL < 0 ? -MCELL : 0) / 4; ++i) {
                    const float4 q = reinterpret_cast<const float4*>(p)[i];
                    vv[4 * i + 0] = static_cast<InT>(q.x);
                    vv[4 * i + 1] = static_cast<InT>(q.y);
                    vv[4 * i + 2] = static_cast<InT>(q.z);
                    vv[4 * i + 3] = static_cast<InT>(q.w);
                }
#pragma unroll
                for (int i = 1; i < (MCELL < 0 ? -MCELL : 1); ++i) {
                    const InT v = vv[i];
                    accum_moments<InT, M4>(a, v, c);
                    if (TD) {
                        const float pos = v > th ? 1.f : 0.f;
                        accum_pair<InT>(a, v, prev, pos, pos_prev);
                        pos_prev = pos;
                    }
                    prev = v;
                }
            } else if (MCELL > 0) {
#pragma unroll
                for (int i = 1; i < (MCELL > 0 ? MCELL : 1); ++i) {
                    const InT v = ld(p, i);
                    accum_moments<InT, M4>(a, v, c);
                    if (TD) {
                        const float pos = v > th ? 1.f : 0.f;
                        accum_pair<InT>(a, v, prev, pos, pos_prev);
                        pos_prev = pos;
                    }
                    prev = v;
                }
            } else {
#pragma unroll 4
                for (int i = 1; i < m; ++i) {
                    const InT v = ld(p, i);
                    accum_moments<InT, M4>(a, v, c);
                    if (TD) {
                        const float pos = v > th ? 1.f : 0.f;
                        accum_pair<InT>(a, v, prev, pos, pos_prev);
                        pos_prev = pos;
                    }
                    prev = v;
                }
            }
            float llb = 0.f, zcb = 0.f;
            if (TD && (ce + 1 < ncell || d.has_next)) {
                const InT nx = ld(p, m);
                llb = fabsf(static_cast<float>(nx - prev));
                zcb = ((nx > th) != (prev > th)) ? 1.f : 0.f;
            }
            const Part& dst = (g_cpb > 1) ? cell : ring;
            const int idx = (g_cpb > 1) ? ce : wrap(ring_head + ce, P.RB);
            dst.s1[idx] = a.s1;
            dst.s2[idx] = a.s2;
            if (M4) {
                dst.s3[idx] = a.s3;
                dst.s4[idx] = a.s4;
            }
            if (sizeof(InT) == 8 && a.nan) a.mn = a.mx = static_cast<InT>(CUDART_NAN);
            dst.mn[idx] = a.mn;
            dst.mx[idx] = a.mx;
            if (TD) {
                dst.ll[idx] = a.ll + llb;
                dst.llb[idx] = a.ll;
                dst.zc[idx] = a.zc + zcb;
                dst.zcb[idx] = a.zc;
            }
        }
        __syncthreads();      // stage buffer fully consumed; cell partials visible

        // ---------------- refill this slot with stage st + NS
        if (tid == 0 && st + P.NS < n_stages) {
            StageDesc<InT> nd;
            nd.set(P, ck, st + P.NS, stage_samples, n_blocks - (st + P.NS) * g_TB, g_TB);
            if (nd.tma) issue_stage(slot, nd);
        }
        if (++slot == P.NS) {
            slot = 0;
            parity ^= 1;
        }

        // ---------------- phase 2: cells -> blocks; one (block, quantity pair) per thread
        if (g_cpb > 1) {
            constexpr int NG = 2 + (M4 ? 1 : 0) + (TD ? 2 : 0);     // {s1,s2} {mn,mx} [{s3,s4}] [{ll,llb} {zc,zcb}]
            const int total = d.nblk * NG;
            for (int idx = tid; idx < total; idx += kThreads) {
                int grp = 0, b = idx;
                while (b >= d.nblk) {
                    b -= d.nblk;
                    ++grp;
                }
                const int c0 = b * g_cpb;
                const int r = wrap(ring_head + b, P.RB);
                if (grp == 0) {
                    double u = 0, v = 0;
                    for (int i = 0; i < g_cpb; ++i) {
                        u += cell.s1[c0 + i];
                        v += cell.s2[c0 + i];
                    }
                    ring.s1[r] = u;
                    ring.s2[r] = v;
                } else if (grp == 1) {
                    InT u = cell.mn[c0], v = cell.mx[c0];
                    for (int i = 1; i < g_cpb; ++i) {
                        u = tmin<InT>(u, cell.mn[c0 + i]);
                        v = tmax<InT>(v, cell.mx[c0 + i]);
                    }
                    ring.mn[r] = u;
                    ring.mx[r] = v;
                } else if (M4 && grp == 2) {
                    double u = 0, v = 0;
                    for (int i = 0; i < g_cpb; ++i) {
                        u += cell.s3[c0 + i];
                        v += cell.s4[c0 + i];
                    }
                    ring.s3[r] = u;
                    ring.s4[r] = v;
                } else if (TD && grp == (M4 ? 3 : 2)) {
                    double u = 0;                    // float cell terms, float64 across cells
                    for (int i = 0; i < g_cpb - 1; ++i) u += static_cast<double>(cell.ll[c0 + i]);
                    ring.ll[r] = static_cast<float>(u + static_cast<double>(cell.ll[c0 + g_cpb - 1]));
                    ring.llb[r] = static_cast<float>(u + static_cast<double>(cell.llb[c0 + g_cpb - 1]));
                } else if (TD) {
                    float u = 0;
                    for (int i = 0; i < g_cpb - 1; ++i) u += cell.zc[c0 + i];
                    ring.zc[r] = u + cell.zc[c0 + g_cpb - 1];
                    ring.zcb[r] = u + cell.zcb[c0 + g_cpb - 1];
                }
            }
            __syncthreads();
        }

        // ---------------- running prefix of the additive quantities (large k only)
        if (use_prefix) {
            const int warp = tid >> 5, lane = tid & 31;
            constexpr int nq = 2 + (M4 ? 2 : 0) + (TD ? 2 : 0);
            if (warp <= nq) {                       // one warp per additive quantity + one for the non-finite count
                // quantity order: s1, s2, [s3, s4], [ll, zc]; q == nq: number of blocks whose partials are not finite.
                // Such a block enters the prefixes as ZERO (a NaN there would poison every later window of the chunk
                // through pre[hi] - pre[lo]); windows that cover one are summed directly in phase 3 instead.
                const int q = warp;
                const bool is_cnt = q == nq;
                const bool is_f = TD && q >= nq - 2 && !is_cnt;
                const double* srcd = q == 0 ? ring.s1 : q == 1 ? ring.s2 : (M4 && q == 2) ? ring.s3 : ring.s4;
                const float* srcf = (q == nq - 2) ? ring.ll : ring.zc;
                const int qs = is_cnt ? kNPre - 1 : q;                            // prefix / carry slot
                double run = carry[qs];
                int pslot = pre_emit + (blocks_done - emitted * g_hop);          // prefix slot of block blocks_done
                while (pslot >= R1) pslot -= R1;                                  // (no integer division in the stage loop)
                while (pslot < 0) pslot += R1;
                for (int base = 0; base < d.nblk; base += 32) {
                    const int b = base + lane;
                    double v = 0.0;
                    if (b < d.nblk) {
                        const int r = wrap(ring_head + b, P.RB);
                        bool bad = !isfinite(ring.s1[r]) || !isfinite(ring.s2[r]);
                        if (M4) bad = bad || !isfinite(ring.s3[r]) || !isfinite(ring.s4[r]);
                        if (TD) bad = bad || !isfinite(ring.ll[r]);
                        if (is_cnt) v = bad ? 1.0 : 0.0;
                        else if (!bad) v = is_f ? static_cast<double>(srcf[r]) : srcd[r];
                    }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double u = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += u;
                    }
                    if (b < d.nblk) pre[qs * R1 + wrap(pslot + b + 1, R1)] = run + v;
                    run += __shfl_sync(0xffffffffu, v, 31);
                }
                if (lane == 0) carry[qs] = run;
            }
            __syncthreads();
        }
        blocks_done += d.nblk;
        ring_head = wrap(ring_head + d.nblk, P.RB);

        // ---------------- phase 3: finished windows (deferred until `flush` are pending)
        int32_t ready = 0;
        if (blocks_done >= g_k) {
            ready = (blocks_done - g_k) / g_hop + 1;
            ready = ready < nwin ? ready : nwin;
        }
        const bool do_flush = (ready - emitted >= P.flush) || (st == n_stages - 1);
        if (do_flush) {
            constexpr int nq = 2 + (M4 ? 2 : 0) + (TD ? 2 : 0);
            for (int wl = emitted + tid; wl < ready; wl += kThreads) {
                const int rel = (wl - emitted) * g_hop;          // < RB by construction
                const int r0 = wrap(ring_emit + rel, P.RB);
                double S1, S2, S3 = 0, S4 = 0, LL = 0, ZC = 0;
                InT mn = ring.mn[r0], mx = ring.mx[r0];
                int r = r0;
                bool direct = !use_prefix;
                int lo = 0, hi = 0;
                if (use_prefix) {
                    lo = wrap(pre_emit + rel, R1);
                    hi = lo + g_k;
                    while (hi >= R1) hi -= R1;
                    direct = pre[(kNPre - 1) * R1 + hi] != pre[(kNPre - 1) * R1 + lo];      // a non-finite block inside
                }
                // LL / ZC: blocks 0 .. k-2 with their right edges, the last block without (that edge belongs to the next
                // window only)
                if (direct) {
                    S1 = ring.s1[r0];
                    S2 = ring.s2[r0];
                    if (M4) {
                        S3 = ring.s3[r0];
                        S4 = ring.s4[r0];
                    }
                    for (int j = 1; j < g_k; ++j) {
                        if (TD) {
                            LL += ring.ll[r];
                            ZC += ring.zc[r];
                        }
                        r = wrap(r + 1, P.RB);
                        S1 += ring.s1[r];
                        S2 += ring.s2[r];
                        if (M4) {
                            S3 += ring.s3[r];
                            S4 += ring.s4[r];
                        }
                        mn = tmin<InT>(mn, ring.mn[r]);
                        mx = tmax<InT>(mx, ring.mx[r]);
                    }
                } else {
                    S1 = pre[0 * R1 + hi] - pre[0 * R1 + lo];
                    S2 = pre[1 * R1 + hi] - pre[1 * R1 + lo];
                    if (M4) {
                        S3 = pre[2 * R1 + hi] - pre[2 * R1 + lo];
                        S4 = pre[3 * R1 + hi] - pre[3 * R1 + lo];
                    }
                    if (TD) {
                        const int hi1 = hi > 0 ? hi - 1 : R1 - 1;
                        LL = pre[(nq - 2) * R1 + hi1] - pre[(nq - 2) * R1 + lo];
                        ZC = pre[(nq - 1) * R1 + hi1] - pre[(nq - 1) * R1 + lo];
                    }
                    for (int j = 1; j < g_k; ++j) {
                        r = wrap(r + 1, P.RB);
                        mn = tmin<InT>(mn, ring.mn[r]);
                        mx = tmax<InT>(mx, ring.mx[r]);
                    }
                }
                if (TD) {
                    LL += ring.llb[r];
                    ZC += ring.zcb[r];
                }

                // shifted power sums -> central moments
                const double n = static_cast<double>(P.W);
                const double dl = S1 * P.inv_n;
                const double mean = c + dl;
                double M2 = S2 - S1 * dl;
                if (M2 < 0.0 || mn == mx) M2 = 0.0;     // constant window: exactly zero, like the two-pass form
                const double var = M2 * P.inv_n;
                const double sd = sqrt(var);
                double skew = 0.0, kurt = 0.0;
                if (M4 && var > 0.0) {
                    const double dl2 = dl * dl;
                    const double M3 = S3 - 3.0 * dl * S2 + 2.0 * n * dl2 * dl;
                    const double M4v = S4 - 4.0 * dl * S3 + 6.0 * dl2 * S2 - 3.0 * n * dl2 * dl2;
                    const double inv_var = 1.0 / var;
                    skew = (M3 * P.inv_n) * inv_var / sd;
                    kurt = (M4v * P.inv_n) * inv_var * inv_var;
                }
                const int64_t obase = series * P.o_series + (w0 + wl) * P.o_window;
                for (int j = 0; j < P.n_features; ++j) {
                    double v;
                    switch (P.feat[j]) {
                        case MHB_F_MEAN: v = mean; break;
                        case MHB_F_VAR:
                        case MHB_F_HJORTH_ACTIVITY: v = var; break;
                        case MHB_F_STD: v = sd; break;
                        case MHB_F_MIN: v = static_cast<double>(mn); break;
                        case MHB_F_MAX: v = static_cast<double>(mx); break;
                        case MHB_F_DRANGE: v = static_cast<double>(mx) - static_cast<double>(mn); break;
                        case MHB_F_SKEWNESS: v = skew; break;
                        case MHB_F_KURTOSIS: v = kurt; break;
                        case MHB_F_KURTOSIS_EXCESS: v = kurt - 3.0; break;
                        case MHB_F_COEFF_VAR: v = sd / mean; break;
                        case MHB_F_ZERO_CROSSINGS: v = ZC; break;
                        case MHB_F_LINE_LENGTH: v = LL; break;
                        case MHB_F_SUM: v = mean * n; break;
                        default: v = 0.0; break;
                    }
                    store_cell<OutT>(P.out, obase + j * P.o_col, v);
                }
            }
            const int adv = (ready - emitted) * g_hop;       // blocks released (may exceed RB when hop > k)
            ring_emit += adv;
            while (ring_emit >= P.RB) ring_emit -= P.RB;
            pre_emit += adv;
            while (pre_emit >= R1) pre_emit -= R1;
            emitted = ready;
        }
        // The next iteration writes ring / prefix slots of NEW blocks only after a barrier that every
        // thread reaches after finishing this phase (cpb > 1: the barrier after phase 1); when phase 1
        // itself writes the ring, close the iteration with one.
        if (g_cpb == 1) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
struct CellChoice {
    int m;
    int mcell_template;   // 0 = runtime, >0 unrolled scalar cell, <0 float4 cell
};

CellChoice choose_cell(int64_t g, bool allow_vec4) {
    CellChoice c;
    if (allow_vec4 && g % 16 == 0 && getenv("MHB_STATS_CELL8") == nullptr) {
        // power-of-two-ish blocks have no odd divisor: 4 x 128-bit loads per thread.  Lane stride 64 B gives a 4-way
        // conflict on LDS.128 (16 wavefronts per warp load, 512 samples per 64 cycles: far above the 4 B/sample HBM
        // can feed), and a 16-sample cell halves the per-stage overhead per sample of the 8-sample one.
        c.m = 16;
        c.mcell_template = -16;
        return c;
    }
    if (allow_vec4 && g % 8 == 0) {
        // 2 x 128-bit loads per thread (lane stride 32 B: 2-way conflict on LDS.128)
        c.m = 8;
        c.mcell_template = -8;
        return c;
    }
    // scalar shared-memory reads at a lane stride of m words: conflict degree gcd(m, 32)
    int best = 1, best_score = -1;
    for (int m = 1; m <= 32; ++m) {
        if (g % m) continue;
        int conflict = static_cast<int>(gcd64(m, 32));
        int score = (conflict == 1 ? 3000 : conflict == 2 ? 2000 : conflict == 4 ? 1000 : 0) + m;
        if (m == 1 && g > 1) score = 1;   // a cell of one sample wastes the thread
        if (score > best_score) {
            best_score = score;
            best = m;
        }
    }
    c.m = best;
    c.mcell_template = (best == 25) ? best : 0;
    return c;
}

template <typename InT, typename OutT, bool M4, bool TD>
cudaError_t launch_with_cell(const StatsPlan& P, int mt, dim3 grid, size_t smem, cudaStream_t stream) {
#define MHB_LAUNCH(MC, GEO)                                                                             \
    {                                                                                                   \
        auto kern = window_stats_kernel<InT, OutT, M4, TD, MC, GEO>;                                    \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                 \
        kern<<<grid, kThreads, smem, stream>>>(P);                                                      \
        return cudaGetLastError();                                                                      \
    }
    if (P.y && P.mag_staged) {
        if constexpr (sizeof(InT) == 4) {
#define MHB_LAUNCH_MAG(MC, GEO)                                                                         \
    {                                                                                                   \
        auto kern = window_stats_kernel<InT, OutT, M4, TD, MC, GEO, true>;                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                 \
        kern<<<grid, kThreads, smem, stream>>>(P);                                                      \
        return cudaGetLastError();                                                                      \
    }
            if (mt == 25) {
                if (P.cpb == 10 && P.k == 2 && P.hop == 1 && P.TB == 25) MHB_LAUNCH_MAG(25, 1)
                MHB_LAUNCH_MAG(25, 0)
            }
            MHB_LAUNCH_MAG(0, 0)
#undef MHB_LAUNCH_MAG
        }
    }
    if (mt == -8) {
        if constexpr (sizeof(InT) == 4) MHB_LAUNCH(-8, 0)
    }
    if (mt == -16) {
        if constexpr (sizeof(InT) == 4) {
            if (P.cpb == 4 && P.k == 30 && P.hop == 1 && P.TB == 64 && getenv("MHB_STATS_NOGEO2") == nullptr) MHB_LAUNCH(-16, 2)
            MHB_LAUNCH(-16, 0)
        }
    }
    if (mt == 25) {
        if (P.cpb == 10 && P.k == 2 && P.hop == 1 && P.TB == 25) MHB_LAUNCH(25, 1)
        MHB_LAUNCH(25, 0)
    }
    MHB_LAUNCH(0, 0)
#undef MHB_LAUNCH
}

template <typename InT>
int32_t window_stats_impl(const InT* x, const mhb_windows* geom, const int32_t* h_features, int32_t n_features,
                          double zc_threshold, const mhb_table* table, void* stream_v, const InT* y = nullptr,
                          const InT* z = nullptr) {
    MHB_REQUIRE(geom && table, MHB_E_ARG, "window_stats: null geometry/table");
    MHB_REQUIRE(geom->wsize >= 1 && geom->wstep >= 1, MHB_E_ARG, "window_stats: wsize and wstep must be >= 1");
    MHB_REQUIRE(geom->n_series >= 0 && geom->series_len >= 0 && geom->series_stride >= geom->series_len,
                MHB_E_ARG, "window_stats: bad series geometry");
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxFeat, MHB_E_ARG, "window_stats: 0..%d features per call", kMaxFeat);
    MHB_REQUIRE(n_features == 0 || h_features, MHB_E_ARG, "window_stats: null feature list");
    StatsPlan P;
    memset(&P, 0, sizeof(P));
    bool m4 = false, td = false;
    for (int j = 0; j < n_features; ++j) {
        const int f = h_features[j];
        MHB_REQUIRE((f >= MHB_F_MEAN && f <= MHB_F_SUM), MHB_E_FEATURE,
                    "window_stats: feature id %d is not in the streaming family", f);
        if (f == MHB_F_SKEWNESS || f == MHB_F_KURTOSIS || f == MHB_F_KURTOSIS_EXCESS) m4 = true;
        if (f == MHB_F_ZERO_CROSSINGS || f == MHB_F_LINE_LENGTH) td = true;
        P.feat[j] = f;
    }
    const int64_t nw = n_windows_host(geom->series_len, geom->wsize, geom->wstep);
    if (nw == 0 || geom->n_series == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(x && table->out, MHB_E_ARG, "window_stats: null data/output pointer");
    MHB_REQUIRE((y == nullptr) == (z == nullptr), MHB_E_ARG, "window_stats: magnitude mode takes both y and z");
    P.n_features = n_features;
    P.x = x;
    P.y = y;
    P.z = z;
    P.series_len = geom->series_len;
    P.series_stride = geom->series_stride;
    P.total_elems = (geom->n_series - 1) * geom->series_stride + geom->series_len;
    P.nw = nw;
    P.W = geom->wsize;
    P.S = nw == 1 ? geom->wsize : geom->wstep;      // a single window: the hop is irrelevant, keep g = W
    P.th = zc_threshold;
    P.inv_n = 1.0 / static_cast<double>(geom->wsize);
    P.out = table->out;
    P.o_series = table->series_stride;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;

    // block size: a divisor of gcd(W, S) small enough for one stage, with k + hop bounded
    int64_t g = gcd64(P.W, P.S);
    constexpr int64_t kMaxStageBytes = 25 * 1024;      // 25 blocks of 250 float32 samples fit one stage
    // magnitude mode, float32: the three axes are staged side by side (a third of the stage budget each, scalar cells)
    const bool mag_staged = y != nullptr && sizeof(InT) == 4 && getenv("MHB_MAG_COPY") == nullptr;
    P.mag_staged = mag_staged ? 1 : 0;
    const int64_t max_block = kMaxStageBytes / static_cast<int64_t>(sizeof(InT));
    if (g > max_block) {
        int64_t best = 1;
        for (int64_t dv = 1; dv * dv <= g; ++dv)
            if (g % dv == 0) {
                if (dv <= max_block && dv > best) best = dv;
                if (g / dv <= max_block && g / dv > best) best = g / dv;
            }
        g = best;
    }
    P.g = static_cast<int32_t>(g);
    const int64_t k64 = P.W / g, hop64 = P.S / g;
    MHB_REQUIRE(!(y && k64 + hop64 > 1024), MHB_E_UNSUPPORTED,
                "window_stats: magnitude mode needs gcd(wsize, wstep) >= (wsize + wstep) / 1024");
    if (k64 + hop64 > 1024)        // e.g. co-prime W and S: no block sharing to exploit, evaluate every window directly
        return window_stats_direct<InT>(x, geom, nw, h_features, n_features, zc_threshold, table, stream_v);
    P.k = static_cast<int32_t>(k64);
    P.hop = static_cast<int32_t>(hop64);
    const CellChoice cc = choose_cell(g, sizeof(InT) == 4 && geom->series_stride % 4 == 0 && !mag_staged);
    P.m = cc.m;
    P.cpb = static_cast<int32_t>(g / cc.m);
    int64_t tb = P.cpb <= kThreads ? kThreads / P.cpb : 1;
    const int64_t tb_cap = kMaxStageBytes / static_cast<int64_t>(g * sizeof(InT));
    if (tb > tb_cap) tb = tb_cap;
    if (tb < 1) tb = 1;
    P.TB = static_cast<int32_t>(tb);
    // deferred finalisation: let about half a CTA of windows pile up before phase 3 runs
    {
        const int64_t per_stage = P.TB / P.hop + 1;
        // heavily overlapping windows (k > kDirectK: one finished window per block) finalize in bigger groups
        int64_t flush = (P.k > kDirectK ? (3 * kThreads) / 4 : kThreads / 2) - per_stage;
        const int64_t cap = (1536 - P.k - P.TB) / P.hop - per_stage - 1;   // keep the ring <= 1536 blocks
        if (flush > cap) flush = cap;
        if (flush < 1) flush = 1;
        P.flush = static_cast<int32_t>(flush);
        P.RB = static_cast<int32_t>((flush + per_stage + 1) * P.hop + P.k + P.TB + 1);
    }
    // ... and trade the second stage buffer for the longer ring (the other CTAs of the SM cover the copy latency)
    P.NS = P.k > kDirectK ? 1 : 2;
    constexpr int A = 16 / sizeof(InT);
    P.stage_elems = ((P.TB * P.g + 1 + (A - 1) + A - 1) / A) * A + A;
    P.use_tma = (reinterpret_cast<uintptr_t>(x) % 16 == 0 && y == nullptr) ? 1 : 0;
    if (mag_staged) {
        P.use_tma = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) % 16 == 0) ? 1 : 0;
        // the same stage geometry as the single-array kernel (so the chunking, and with it every rounding, is the same),
        // three arrays per slot, ONE slot: the two CTAs of an SM cover each other's copy latency (measured: 2.2 ms
        // against 3.9 ms with two slots and one CTA per SM, tools/perf_magnitude.py)
        P.NS = 1;
    }

    size_t smem = 128 + static_cast<size_t>(P.NS) * (mag_staged ? 3 : 1) * P.stage_elems * sizeof(InT);
    if (P.cpb > 1) smem += partial_bytes(sizeof(InT), m4, td, P.TB * P.cpb);
    smem += partial_bytes(sizeof(InT), m4, td, P.RB);
    if (P.k > kDirectK) smem += sizeof(double) * kNPre * (P.RB + 1);
    smem += 64;
    MHB_REQUIRE(!(y && smem > 220 * 1024), MHB_E_UNSUPPORTED, "window_stats: magnitude mode: window geometry too awkward");
    if (smem > 220 * 1024)           // ring + prefix arrays of an awkward (k, hop) do not fit: evaluate every window directly
        return window_stats_direct<InT>(x, geom, nw, h_features, n_features, zc_threshold, table, stream_v);

    // chunking: enough CTAs to fill the machine several times, each long enough to amortise the
    // pipeline fill and the (k - hop) halo blocks, short enough to keep the pivot local
    const int64_t win_per_stage = P.TB / P.hop > 0 ? P.TB / P.hop : 1;
    const int64_t total_windows = nw * geom->n_series;
    const int64_t target_ctas = static_cast<int64_t>(kNumSMs) * 12;
    // a chunk is segment / 2^j windows (segment = 64 stages): halve while the launch has too few CTAs
    const int64_t seg = 64 * win_per_stage;
    int64_t wpc = seg;
    while (wpc % 2 == 0 && wpc / 2 >= 8 * win_per_stage && (wpc / 2) % win_per_stage == 0 &&
           (total_windows + wpc - 1) / wpc < target_ctas)
        wpc /= 2;
    P.win_per_segment = seg;
    P.win_per_chunk = wpc;
    const int64_t cps = (nw + wpc - 1) / wpc;
    MHB_REQUIRE(cps * geom->n_series < (1LL << 31), MHB_E_UNSUPPORTED, "window_stats: too many chunks");
    P.chunks_per_series = static_cast<int32_t>(cps);
    const dim3 grid(static_cast<unsigned>(cps * geom->n_series));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

    cudaError_t e;
    const bool f32 = table->out_f32 != 0;
#define MHB_DISPATCH(M4v, TDv)                                                                        \
    e = f32 ? launch_with_cell<InT, float, M4v, TDv>(P, cc.mcell_template, grid, smem, stream)        \
            : launch_with_cell<InT, double, M4v, TDv>(P, cc.mcell_template, grid, smem, stream)
    if (m4 && td) {
        MHB_DISPATCH(true, true);
    } else if (m4) {
        MHB_DISPATCH(true, false);
    } else if (td) {
        MHB_DISPATCH(false, true);
    } else {
        MHB_DISPATCH(false, false);
    }
#undef MHB_DISPATCH
    return cuda_status(e, "window_stats launch");
}

}  // namespace

}  // namespace mhb

extern "C" int32_t mhb_window_stats_f32(const float* x, const mhb_windows* geom, const int32_t* h_features,
                                        int32_t n_features, double zc_threshold, const mhb_table* table,
                                        void* stream) {
    return mhb::window_stats_impl<float>(x, geom, h_features, n_features, zc_threshold, table, stream);
}

extern "C" int32_t mhb_window_stats_f64(const double* x, const mhb_windows* geom, const int32_t* h_features,
                                        int32_t n_features, double zc_threshold, const mhb_table* table,
                                        void* stream) {
    return mhb::window_stats_impl<double>(x, geom, h_features, n_features, zc_threshold, table, stream);
}

// Fused pre-stage of SURVEY 8f-1: the statistics of magnitude(x, y, z) (inertial/accelerometer.py:198-225) without
// materialising the magnitude series.  x, y, z share the geometry (same length, stride, series count).
extern "C" int32_t mhb_window_stats_magnitude_f32(const float* x, const float* y, const float* z, const mhb_windows* geom,
                                                  const int32_t* h_features, int32_t n_features, double zc_threshold,
                                                  const mhb_table* table, void* stream) {
    MHB_REQUIRE(y && z, MHB_E_ARG, "window_stats_magnitude: null axis pointer");
    return mhb::window_stats_impl<float>(x, geom, h_features, n_features, zc_threshold, table, stream, y, z);
}

extern "C" int32_t mhb_window_stats_magnitude_f64(const double* x, const double* y, const double* z, const mhb_windows* geom,
                                                  const int32_t* h_features, int32_t n_features, double zc_threshold,
                                                  const mhb_table* table, void* stream) {
    MHB_REQUIRE(y && z, MHB_E_ARG, "window_stats_magnitude: null axis pointer");
    return mhb::window_stats_impl<double>(x, geom, h_features, n_features, zc_threshold, table, stream, y, z);
}
