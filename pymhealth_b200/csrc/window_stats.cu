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
//   phase2: cells -> block partials, written to a ring of the last TB + k + hop blocks.
//   phase3: one thread per finished window combines k ring entries (direct sums for small k,
//           differences of a running prefix for large k), turns the shifted power sums into
//           central moments, and stores the requested feature columns.
// Parity notes (SURVEY section 8c gotchas): population variance; kurtosis / skewness return 0 for a
// constant window; zero is "not positive" for crossings; no partial tail window.
#include "common.cuh"

namespace mhb {

namespace {

enum : int { Q_S1 = 0, Q_S2, Q_S3, Q_S4, Q_LL, Q_ZC, Q_LLB, Q_ZCB, Q_MN, Q_MX, NQ };
constexpr int kThreads = 256;
constexpr int kMaxFeat = 32;
constexpr int kDirectK = 8;      // windows of <= kDirectK blocks are summed directly
constexpr int kNAdd = 6;         // Q_S1..Q_ZC are additive and get a running prefix

struct StatsPlan {
    const void* x;
    int64_t series_len, series_stride, total_elems;
    int64_t nw, win_per_chunk;
    int32_t chunks_per_series;
    int32_t W, S, g, k, hop, m, cpb, TB, RB, NS;
    int32_t stage_elems;
    int32_t use_tma;
    double th;
    void* out;
    int64_t o_series, o_window, o_col;
    int32_t n_features;
    int32_t feat[kMaxFeat];
};

template <typename InT>
struct CellAcc {
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    float ll = 0.f;
    int zc = 0;
    InT mn, mx;
};

template <typename InT, bool M4, bool TD>
__device__ __forceinline__ void accum(CellAcc<InT>& a, InT v, InT prev, bool has_prev, double c, InT th) {
    const double d = static_cast<double>(v) - c;
    const double d2 = d * d;
    a.s1 += d;
    a.s2 += d2;
    if (M4) {
        a.s3 = fma(d2, d, a.s3);
        a.s4 = fma(d2, d2, a.s4);
    }
    a.mn = v < a.mn ? v : a.mn;
    a.mx = v > a.mx ? v : a.mx;
    if (TD && has_prev) {
        a.ll += fabsf(static_cast<float>(v - prev));
        a.zc += ((v > th) != (prev > th)) ? 1 : 0;
    }
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

struct StageDesc {
    int64_t goff;       // global element offset of the stage's first sample
    int64_t a0;         // aligned-down element offset the copy starts at
    int32_t lead;       // goff - a0
    int32_t nblk;       // blocks in this stage
    int32_t cnt;        // samples belonging to blocks
    int32_t has_next;   // the sample after the stage exists in the series
    int32_t n_load;     // elements copied by TMA (multiple of 16 bytes)
    int32_t tma;        // stage is loaded by TMA (else cooperative guarded copy)
};

template <typename InT>
__device__ __forceinline__ StageDesc describe_stage(const StatsPlan& P, int64_t series, int64_t blk_begin,
                                                    int32_t n_blocks, int32_t st) {
    constexpr int A = 16 / sizeof(InT);
    StageDesc d;
    const int32_t b0 = st * P.TB;
    d.nblk = min(P.TB, n_blocks - b0);
    d.cnt = d.nblk * P.g;
    const int64_t s0 = (blk_begin + b0) * static_cast<int64_t>(P.g);
    d.has_next = (s0 + d.cnt < P.series_len) ? 1 : 0;
    d.goff = series * P.series_stride + s0;
    d.a0 = d.goff & ~static_cast<int64_t>(A - 1);
    d.lead = static_cast<int32_t>(d.goff - d.a0);
    d.n_load = (d.lead + d.cnt + d.has_next + A - 1) & ~(A - 1);
    d.tma = (P.use_tma && d.a0 + d.n_load <= P.total_elems) ? 1 : 0;
    return d;
}

// ---------------------------------------------------------------------------------------------
template <typename InT, typename OutT, bool M4, bool TD, int MCELL /*0 = runtime m*/>
__global__ void __launch_bounds__(kThreads, 2) window_stats_kernel(const StatsPlan P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int m = MCELL ? MCELL : P.m;

    // ---- carve shared memory
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // NS barriers (<= 8)
    unsigned char* ptr = smem_raw + 128;
    InT* stage_buf = reinterpret_cast<InT*>(ptr);
    ptr += static_cast<size_t>(P.NS) * P.stage_elems * sizeof(InT);
    ptr = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ptr) + 15) & ~uintptr_t(15));
    const int ncell_max = P.TB * P.cpb;
    double* cell = reinterpret_cast<double*>(ptr);                          // [NQ][ncell_max] (cpb > 1)
    if (P.cpb > 1) ptr += sizeof(double) * NQ * ncell_max;
    double* ring = reinterpret_cast<double*>(ptr);                          // [NQ][RB]
    ptr += sizeof(double) * NQ * P.RB;
    double* pre = reinterpret_cast<double*>(ptr);                           // [kNAdd][RB + 1] (k > kDirectK)
    __shared__ double carry[kNAdd];

    // ---- which chunk
    const int64_t series = blockIdx.x / P.chunks_per_series;
    const int32_t chunk = blockIdx.x % P.chunks_per_series;
    const int64_t w0 = static_cast<int64_t>(chunk) * P.win_per_chunk;
    const int64_t w1 = min(P.nw, w0 + P.win_per_chunk);
    const int32_t nwin = static_cast<int32_t>(w1 - w0);
    const int64_t blk_begin = w0 * P.hop;
    const int32_t n_blocks = (nwin - 1) * P.hop + P.k;
    const int32_t n_stages = (n_blocks + P.TB - 1) / P.TB;
    const bool use_prefix = P.k > kDirectK;
    const InT* xg = reinterpret_cast<const InT*>(P.x);

    if (tid == 0) {
        for (int i = 0; i < P.NS; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    if (tid < kNAdd) carry[tid] = 0.0;
    if (use_prefix && tid < kNAdd) pre[tid * (P.RB + 1)] = 0.0;           // prefix before block 0
    __syncthreads();

    if (tid == 0) {
        for (int st = 0; st < P.NS && st < n_stages; ++st) {
            const StageDesc d = describe_stage<InT>(P, series, blk_begin, n_blocks, st);
            if (d.tma) {
                mbar_arrive_expect_tx(&full[st], d.n_load * sizeof(InT));
                bulk_g2s(stage_buf + static_cast<size_t>(st) * P.stage_elems, xg + d.a0, d.n_load * sizeof(InT),
                         &full[st]);
            }
        }
    }

    const InT th = round_down_threshold<InT>(P.th > 0.0 ? P.th : 0.0);
    double c = 0.0;            // pivot of the shifted power sums: first sample of the chunk
    int32_t blocks_done = 0;
    int32_t emitted = 0;

    for (int st = 0; st < n_stages; ++st) {
        const int slot = st % P.NS;
        const StageDesc d = describe_stage<InT>(P, series, blk_begin, n_blocks, st);
        InT* buf = stage_buf + static_cast<size_t>(slot) * P.stage_elems;
        if (d.tma) {
            mbar_wait(&full[slot], (st / P.NS) & 1);
        } else {
            // guarded cooperative copy: unaligned base pointer or the last few samples of the buffer
            const int n = d.cnt + d.has_next;
            for (int i = tid; i < n; i += kThreads) buf[d.lead + i] = xg[d.goff + i];
            __syncthreads();
        }
        const InT* s = buf + d.lead;
        if (st == 0) c = static_cast<double>(s[0]);

        // ---------------- phase 1: one cell per thread
        const int ncell = d.nblk * P.cpb;
        for (int ce = tid; ce < ncell; ce += kThreads) {
            const InT* p = s + ce * m;
            CellAcc<InT> a;
            a.mn = p[0];
            a.mx = p[0];
            InT prev = p[0];
            accum<InT, M4, TD>(a, prev, prev, false, c, th);
            if (MCELL) {
#pragma unroll
                for (int i = 1; i < (MCELL ? MCELL : 1); ++i) {
                    const InT v = p[i];
                    accum<InT, M4, TD>(a, v, prev, true, c, th);
                    prev = v;
                }
            } else {
#pragma unroll 4
                for (int i = 1; i < m; ++i) {
                    const InT v = p[i];
                    accum<InT, M4, TD>(a, v, prev, true, c, th);
                    prev = v;
                }
            }
            double llb = 0.0, zcb = 0.0;
            if (TD && (ce + 1 < ncell || d.has_next)) {
                const InT nx = p[m];
                llb = fabsf(static_cast<float>(nx - prev));
                zcb = ((nx > th) != (prev > th)) ? 1.0 : 0.0;
            }
            double* dst;
            int idx, stride;
            if (P.cpb > 1) {
                dst = cell; idx = ce; stride = ncell_max;
            } else {
                dst = ring; idx = (blocks_done + ce) % P.RB; stride = P.RB;
            }
            dst[Q_S1 * stride + idx] = a.s1;
            dst[Q_S2 * stride + idx] = a.s2;
            dst[Q_S3 * stride + idx] = a.s3;
            dst[Q_S4 * stride + idx] = a.s4;
            dst[Q_LL * stride + idx] = static_cast<double>(a.ll) + llb;      // "full": inside + right edge
            dst[Q_ZC * stride + idx] = static_cast<double>(a.zc) + zcb;
            dst[Q_LLB * stride + idx] = llb;
            dst[Q_ZCB * stride + idx] = zcb;
            dst[Q_MN * stride + idx] = static_cast<double>(a.mn);
            dst[Q_MX * stride + idx] = static_cast<double>(a.mx);
        }
        __syncthreads();      // stage buffer fully consumed; cell partials visible

        // ---------------- refill this slot with stage st + NS
        if (tid == 0 && st + P.NS < n_stages) {
            const StageDesc nd = describe_stage<InT>(P, series, blk_begin, n_blocks, st + P.NS);
            if (nd.tma) {
                mbar_arrive_expect_tx(&full[slot], nd.n_load * sizeof(InT));
                bulk_g2s(buf, xg + nd.a0, nd.n_load * sizeof(InT), &full[slot]);
            }
        }

        // ---------------- phase 2: cells -> blocks (quantity-major so warps stay uniform)
        if (P.cpb > 1) {
            const int total = d.nblk * NQ;
            for (int idx = tid; idx < total; idx += kThreads) {
                const int q = idx / d.nblk;
                const int b = idx - q * d.nblk;
                const double* src = cell + q * ncell_max + b * P.cpb;
                double r;
                if (q <= Q_ZC) {
                    r = 0.0;
                    for (int i = 0; i < P.cpb; ++i) r += src[i];
                } else if (q <= Q_ZCB) {
                    r = src[P.cpb - 1];
                } else if (q == Q_MN) {
                    r = src[0];
                    for (int i = 1; i < P.cpb; ++i) r = fmin(r, src[i]);
                } else {
                    r = src[0];
                    for (int i = 1; i < P.cpb; ++i) r = fmax(r, src[i]);
                }
                ring[q * P.RB + (blocks_done + b) % P.RB] = r;
            }
            __syncthreads();
        }

        // ---------------- running prefix of the additive quantities (large k only)
        if (use_prefix) {
            const int warp = tid >> 5, lane = tid & 31;
            if (warp == (st & 7)) {                 // rotate the serial work over the SM sub-partitions
                for (int q = 0; q < kNAdd; ++q) {
                    double run = carry[q];
                    for (int base = 0; base < d.nblk; base += 32) {
                        const int b = base + lane;
                        double v = (b < d.nblk) ? ring[q * P.RB + (blocks_done + b) % P.RB] : 0.0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const double u = __shfl_up_sync(0xffffffffu, v, o);
                            if (lane >= o) v += u;
                        }
                        if (b < d.nblk) pre[q * (P.RB + 1) + (blocks_done + b + 1) % (P.RB + 1)] = run + v;
                        run += __shfl_sync(0xffffffffu, v, 31);
                    }
                    __syncwarp();
                    if (lane == 0) carry[q] = run;
                }
            }
            __syncthreads();
        }
        blocks_done += d.nblk;

        // ---------------- phase 3: finished windows
        int32_t ready = 0;
        if (blocks_done >= P.k) ready = min(nwin, (blocks_done - P.k) / P.hop + 1);
        for (int wl = emitted + tid; wl < ready; wl += kThreads) {
            const int b0 = wl * P.hop;
            double S1, S2, S3 = 0, S4 = 0, LL = 0, ZC = 0;
            const int last = (b0 + P.k - 1) % P.RB;
            if (!use_prefix) {
                S1 = S2 = 0.0;
                for (int j = 0; j < P.k; ++j) {
                    const int r = (b0 + j) % P.RB;
                    S1 += ring[Q_S1 * P.RB + r];
                    S2 += ring[Q_S2 * P.RB + r];
                    if (M4) {
                        S3 += ring[Q_S3 * P.RB + r];
                        S4 += ring[Q_S4 * P.RB + r];
                    }
                    if (TD) {
                        LL += ring[Q_LL * P.RB + r];
                        ZC += ring[Q_ZC * P.RB + r];
                    }
                }
            } else {
                const int R1 = P.RB + 1;
                const int hi = (b0 + P.k) % R1, lo = b0 % R1;
                S1 = pre[Q_S1 * R1 + hi] - pre[Q_S1 * R1 + lo];
                S2 = pre[Q_S2 * R1 + hi] - pre[Q_S2 * R1 + lo];
                if (M4) {
                    S3 = pre[Q_S3 * R1 + hi] - pre[Q_S3 * R1 + lo];
                    S4 = pre[Q_S4 * R1 + hi] - pre[Q_S4 * R1 + lo];
                }
                if (TD) {
                    LL = pre[Q_LL * R1 + hi] - pre[Q_LL * R1 + lo];
                    ZC = pre[Q_ZC * R1 + hi] - pre[Q_ZC * R1 + lo];
                }
            }
            if (TD) {
                LL -= ring[Q_LLB * P.RB + last];
                ZC -= ring[Q_ZCB * P.RB + last];
            }
            double mn = ring[Q_MN * P.RB + b0 % P.RB], mx = ring[Q_MX * P.RB + b0 % P.RB];
            for (int j = 1; j < P.k; ++j) {
                const int r = (b0 + j) % P.RB;
                mn = fmin(mn, ring[Q_MN * P.RB + r]);
                mx = fmax(mx, ring[Q_MX * P.RB + r]);
            }

            // shifted power sums -> central moments
            const double n = static_cast<double>(P.W);
            const double dl = S1 / n;
            const double mean = c + dl;
            double M2 = S2 - S1 * dl;
            if (M2 < 0.0 || mn == mx) M2 = 0.0;        // constant window: exactly zero, like the two-pass form
            const double var = M2 / n;
            const double sd = sqrt(var);
            double skew = 0.0, kurt = 0.0;
            if (M4 && var > 0.0) {
                const double M3 = S3 - 3.0 * dl * S2 + 2.0 * n * dl * dl * dl;
                const double M4v = S4 - 4.0 * dl * S3 + 6.0 * dl * dl * S2 - 3.0 * n * dl * dl * dl * dl;
                skew = (M3 / n) / (sd * sd * sd);
                kurt = (M4v / n) / (var * var);
            }
            const int64_t obase = series * P.o_series + (w0 + wl) * P.o_window;
            for (int j = 0; j < P.n_features; ++j) {
                double v;
                switch (P.feat[j]) {
                    case MHB_F_MEAN: v = mean; break;
                    case MHB_F_VAR:
                    case MHB_F_HJORTH_ACTIVITY: v = var; break;
                    case MHB_F_STD: v = sd; break;
                    case MHB_F_MIN: v = mn; break;
                    case MHB_F_MAX: v = mx; break;
                    case MHB_F_DRANGE: v = mx - mn; break;
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
        emitted = ready > emitted ? ready : emitted;
        // no barrier needed here: the next iteration only writes ring / prefix entries of NEW blocks,
        // which never alias entries a still-pending window of this iteration reads (RB >= TB+k+hop+1),
        // and its first block-level write happens after the barrier that follows phase 1.
        if (P.cpb == 1) __syncthreads();   // ...except when phase 1 itself writes the ring
    }
}

// ---------------------------------------------------------------------------------------------
struct CellChoice {
    int m;
    int mcell_template;   // 0 = runtime
};

CellChoice choose_cell(int64_t g) {
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
    CellChoice c;
    c.m = best;
    c.mcell_template = (best == 25 || best == 8 || best == 16 || best == 32) ? best : 0;
    return c;
}

template <typename InT, typename OutT, bool M4, bool TD>
cudaError_t launch_with_cell(const StatsPlan& P, int mt, dim3 grid, size_t smem, cudaStream_t stream) {
#define MHB_LAUNCH(MC)                                                                                  \
    {                                                                                                   \
        auto kern = window_stats_kernel<InT, OutT, M4, TD, MC>;                                         \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                 \
        kern<<<grid, kThreads, smem, stream>>>(P);                                                      \
        return cudaGetLastError();                                                                      \
    }
    switch (mt) {
        case 25: MHB_LAUNCH(25)
        case 8: MHB_LAUNCH(8)
        case 16: MHB_LAUNCH(16)
        case 32: MHB_LAUNCH(32)
        default: MHB_LAUNCH(0)
    }
#undef MHB_LAUNCH
}

template <typename InT>
int32_t window_stats_impl(const InT* x, const mhb_windows* geom, const int32_t* h_features, int32_t n_features,
                          double zc_threshold, const mhb_table* table, void* stream_v) {
    MHB_REQUIRE(geom && table, MHB_E_ARG, "window_stats: null geometry/table");
    MHB_REQUIRE(geom->wsize >= 1 && geom->wstep >= 1, MHB_E_ARG, "window_stats: wsize and wstep must be >= 1");
    MHB_REQUIRE(geom->n_series >= 0 && geom->series_len >= 0 && geom->series_stride >= geom->series_len,
                MHB_E_ARG, "window_stats: bad series geometry");
    MHB_REQUIRE(n_features >= 0 && n_features <= kMaxFeat, MHB_E_ARG, "window_stats: 0..%d features per call", kMaxFeat);
    MHB_REQUIRE(n_features == 0 || h_features, MHB_E_ARG, "window_stats: null feature list");
    const int64_t nw = n_windows_host(geom->series_len, geom->wsize, geom->wstep);
    if (nw == 0 || geom->n_series == 0 || n_features == 0) return MHB_OK;
    MHB_REQUIRE(x && table->out, MHB_E_ARG, "window_stats: null data/output pointer");

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
    P.n_features = n_features;
    P.x = x;
    P.series_len = geom->series_len;
    P.series_stride = geom->series_stride;
    P.total_elems = (geom->n_series - 1) * geom->series_stride + geom->series_len;
    P.nw = nw;
    P.W = geom->wsize;
    P.S = geom->wstep;
    P.th = zc_threshold;
    P.out = table->out;
    P.o_series = table->series_stride;
    P.o_window = table->window_stride;
    P.o_col = table->column_stride;

    // block size: a divisor of gcd(W, S) small enough for one stage, with k + hop bounded
    int64_t g = gcd64(P.W, P.S);
    constexpr int64_t kMaxStageBytes = 32 * 1024;
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
    MHB_REQUIRE(k64 + hop64 <= 4096, MHB_E_UNSUPPORTED,
                "window_stats: wsize=%d wstep=%d needs %lld + %lld blocks per window/hop (> 4096); "
                "use the large-window path",
                P.W, P.S, (long long)k64, (long long)hop64);
    P.k = static_cast<int32_t>(k64);
    P.hop = static_cast<int32_t>(hop64);
    const CellChoice cc = choose_cell(g);
    P.m = cc.m;
    P.cpb = static_cast<int32_t>(g / cc.m);
    int64_t tb = P.cpb <= kThreads ? kThreads / P.cpb : 1;
    tb = tb < kMaxStageBytes / static_cast<int64_t>(g * sizeof(InT)) ? tb : kMaxStageBytes / (g * sizeof(InT));
    if (tb < 1) tb = 1;
    if (tb > 256) tb = 256;
    P.TB = static_cast<int32_t>(tb);
    P.RB = P.TB + P.k + P.hop + 1;
    P.NS = 3;
    constexpr int A = 16 / sizeof(InT);
    P.stage_elems = ((P.TB * P.g + 1 + (A - 1) + A - 1) / A) * A + A;
    P.use_tma = (reinterpret_cast<uintptr_t>(x) % 16 == 0) ? 1 : 0;

    size_t smem = 128 + static_cast<size_t>(P.NS) * P.stage_elems * sizeof(InT) + 16;
    if (P.cpb > 1) smem += sizeof(double) * NQ * P.TB * P.cpb;
    smem += sizeof(double) * NQ * P.RB;
    if (P.k > kDirectK) smem += sizeof(double) * kNAdd * (P.RB + 1);
    MHB_REQUIRE(smem <= 220 * 1024, MHB_E_UNSUPPORTED, "window_stats: geometry needs %zu bytes of shared memory", smem);

    // chunking: enough CTAs to fill the machine several times, each long enough to amortise the
    // pipeline fill and the (k - hop) halo blocks, short enough to keep the pivot local
    const int64_t win_per_stage = P.TB / P.hop > 0 ? P.TB / P.hop : 1;
    const int64_t total_windows = nw * geom->n_series;
    const int64_t target_ctas = static_cast<int64_t>(kNumSMs) * 8;
    int64_t wpc = (total_windows + target_ctas - 1) / target_ctas;
    if (wpc < 8 * win_per_stage) wpc = 8 * win_per_stage;
    if (wpc > 64 * win_per_stage) wpc = 64 * win_per_stage;
    if (wpc > nw) wpc = nw;
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
