// Kernel 3 -- location traces (sm_100a): haversine family, successive / home distances, per-segment
// (subject-day) feature rows with stay-point assignment and label entropy, label statistics.
//
// Reference: src/mhealth/location/distance.py:4-59 (haversine + three gufuncs, float64 only),
// location/features.py:43-113 (distance from home, proportion home stay, successive distance),
// location/distribution.py:28-102 (location variance, label counting, label entropy).
// Radius of gyration and stay-point assignment are EXTENSIONS (no reference implementation; the
// definition is oracle/location_ext.py).
//
// All arithmetic is float64, like the reference's gufunc signatures: latitude / longitude differences
// of ~1e-7 rad at 1 Hz GPS do not survive float32, and bit-exact counts (points within `limit` of
// home, stay-point labels) need the same comparison the float64 reference makes.
#include <cmath>
#include <math_constants.h>

#include "common.cuh"

namespace mhb {

namespace {

constexpr double kEarthDiameterKm = 12742.018;            // distance.py:8,18  (2 * 6371.009)
constexpr double kDeg = 0.017453292519943295;             // pi / 180, np.radians

// sin / asin of the SMALL arguments this path lives on (half-differences of neighbouring fixes are ~1e-7 rad, a trace
// spans ~1e-3 rad): below the cut the odd Taylor polynomial is exact to < 1e-19 relative (last dropped term
// x^16 / 17! and ~0.022 y^10), so it is as accurate as the library routine (<= 1 ulp) without its range reduction
// (53 -> ~12 instructions per call); larger arguments take the library path.
__device__ __forceinline__ double sin_fast(double x) {
    if (fabs(x) > 0.5) return sin(x);
    const double s = x * x;
    double p = 1.0 / 1307674368000.0;               // 1 / 15!
    p = fma(p, s, -1.0 / 6227020800.0);             // 1 / 13!
    p = fma(p, s, 1.0 / 39916800.0);                // 1 / 11!
    p = fma(p, s, -1.0 / 362880.0);                 // 1 / 9!
    p = fma(p, s, 1.0 / 5040.0);                    // 1 / 7!
    p = fma(p, s, -1.0 / 120.0);                    // 1 / 5!
    p = fma(p, s, 1.0 / 6.0);                       // 1 / 3!
    return fma(-x * s, p, x);
}
// cos of a latitude in radians: |x| <= pi/2 always holds for valid data, so two Taylor kernels replace the library's
// general argument reduction: cos x on |x| <= pi/4 (last dropped term x^18 / 18! < 3e-18), and sin(pi/2 - |x|) with
// pi/2 split in two doubles above that.  Anything outside falls back to cos().
__device__ __forceinline__ double cos_lat(double x) {
    const double ax = fabs(x);
    if (!(ax <= 1.5707963267948966)) return cos(x);
    if (ax <= 0.7853981633974483) {
        const double s = x * x;
        double p = 1.0 / 20922789888000.0;           // 1 / 16!
        p = fma(p, s, -1.0 / 87178291200.0);         // 1 / 14!
        p = fma(p, s, 1.0 / 479001600.0);            // 1 / 12!
        p = fma(p, s, -1.0 / 3628800.0);             // 1 / 10!
        p = fma(p, s, 1.0 / 40320.0);                // 1 / 8!
        p = fma(p, s, -1.0 / 720.0);                 // 1 / 6!
        p = fma(p, s, 1.0 / 24.0);                   // 1 / 4!
        p = fma(p, s, -0.5);
        return fma(p, s, 1.0);
    }
    const double y = (1.5707963267948966 - ax) + 6.123233995736766e-17;      // pi/2 = hi + lo
    const double s = y * y;
    double p = 1.0 / 355687428096000.0;              // 1 / 17!
    p = fma(p, s, -1.0 / 1307674368000.0);           // 1 / 15!
    p = fma(p, s, 1.0 / 6227020800.0);               // 1 / 13!
    p = fma(p, s, -1.0 / 39916800.0);                // 1 / 11!
    p = fma(p, s, 1.0 / 362880.0);                   // 1 / 9!
    p = fma(p, s, -1.0 / 5040.0);                    // 1 / 7!
    p = fma(p, s, 1.0 / 120.0);                      // 1 / 5!
    p = fma(p, s, -1.0 / 6.0);                       // 1 / 3!
    return fma(y * s, p, y);
}
__device__ __forceinline__ double asin_fast(double y) {
    if (!(fabs(y) < 0.02)) return asin(y);
    const double s = y * y;
    double p = 105.0 / 3456.0;
    p = fma(p, s, 15.0 / 336.0);
    p = fma(p, s, 3.0 / 40.0);
    p = fma(p, s, 1.0 / 6.0);
    return fma(y * s, p, y);
}

// distance.py:4-19 -- every input converted to radians BEFORE the subtraction
__device__ __forceinline__ double haversine(double lat1, double lon1, double lat2, double lon2) {
    // products rounded BEFORE the subtraction, as numpy does (no fused multiply-subtract): the step
    // between two 1 Hz fixes is ~1e-7 rad, and a contracted fma(lat2, k, -a1) changes it by ~1e-10 relative
    const double a1 = __dmul_rn(lat1, kDeg), a2 = __dmul_rn(lat2, kDeg);
    const double o1 = __dmul_rn(lon1, kDeg), o2 = __dmul_rn(lon2, kDeg);
    const double sa = sin_fast(__dsub_rn(a2, a1) / 2.0);
    const double so = sin_fast(__dsub_rn(o2, o1) / 2.0);
    const double h = sa * sa + (cos_lat(a1) * cos_lat(a2) * (so * so));
    return kEarthDiameterKm * asin_fast(sqrt(h));
}

__global__ void haversine_elementwise_kernel(const double* __restrict__ lat1, const double* __restrict__ lon1,
                                             const double* __restrict__ lat2, const double* __restrict__ lon2,
                                             int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = haversine(lat1[i], lon1[i], lat2[i], lon2[i]);
}

__global__ void haversine_vector_kernel(double lat, double lon, const double* __restrict__ latcol,
                                        const double* __restrict__ loncol, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = haversine(lat, lon, latcol[i], loncol[i]);
}

__global__ void haversine_outer_kernel(const double* __restrict__ lat1, const double* __restrict__ lon1, int64_t n,
                                       const double* __restrict__ lat2, const double* __restrict__ lon2, int64_t m,
                                       double* __restrict__ out) {
    const int64_t total = n * m;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = idx / m, j = idx - i * m;
        out[idx] = haversine(lat1[i], lon1[i], lat2[j], lon2[j]);
    }
}

// features.py:98-113: dist[0] = 0, dist[i] = haversine(p[i-1], p[i])
__global__ void successive_distance_kernel(const double* __restrict__ lat, const double* __restrict__ lon, int64_t n,
                                           double* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = i == 0 ? 0.0 : haversine(lat[i - 1], lon[i - 1], lat[i], lon[i]);
}

__global__ void zero_segment_starts_kernel(const int64_t* __restrict__ offs, int64_t n_segments, int64_t n,
                                           double* __restrict__ out) {
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n_segments;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t a = offs[k];
        if (a < n && a < offs[k + 1]) out[a] = 0.0;
    }
}

// information.py:10-20 term for one count out of n: p = c/n + 1e-30; p ln p
__device__ __forceinline__ double plogp(double cnt, double n) {
    const double p = cnt / n + 1e-30;
    return p * log(p);
}

// haversine split in two so that shared sub-expressions are evaluated once per point:
//   hav_h   = sin^2(dlat / 2) + cos(lat1) cos(lat2) sin^2(dlon / 2)   (inputs already in radians, cosines supplied)
//   hav_km  = 2 r asin(sqrt(h))
// Same operations in the same order as haversine() above.
__device__ __forceinline__ double hav_h(double a1, double o1, double c1, double a2, double o2, double c2) {
    const double sa = sin_fast(__dsub_rn(a2, a1) / 2.0);
    const double so = sin_fast(__dsub_rn(o2, o1) / 2.0);
    return sa * sa + (c1 * c2 * (so * so));
}
__device__ __forceinline__ double hav_km(double h) { return kEarthDiameterKm * asin_fast(sqrt(h)); }

// Comparison of a distance with a threshold through h: d(h) = 2 r asin(sqrt(h)) is monotone, so away from the
// threshold the comparison of h with sin^2(D / 2r) decides; inside a 1e-9 relative guard band the distance itself is
// evaluated and compared exactly as the reference would (bit-exact counts and labels).
struct HavThreshold {
    double d, h, lo, hi;
    __device__ __forceinline__ void set(double dist) {
        d = dist;
        const double sn = sin(dist / kEarthDiameterKm);
        h = sn * sn;
        lo = h * (1.0 - 1e-9);
        hi = h * (1.0 + 1e-9);
        if (!(dist >= 0.0)) lo = hi = h = -1.0;              // negative / NaN thresholds: always evaluate exactly
        if (dist >= kEarthDiameterKm * 1.5) lo = hi = h = -1.0;    // beyond asin's range: evaluate exactly
    }
    __device__ __forceinline__ bool greater(double hv) const {     // hav_km(hv) > d
        if (h < 0.0 || (hv >= lo && hv <= hi)) return hav_km(hv) > d;
        return hv > h;
    }
    __device__ __forceinline__ bool less(double hv) const {        // hav_km(hv) < d
        if (h < 0.0 || (hv >= lo && hv <= hi)) return hav_km(hv) < d;
        return hv < h;
    }
};

// One warp per segment.
__global__ void __launch_bounds__(256, 4) location_segments_kernel(
    const double* __restrict__ lat, const double* __restrict__ lon, const int64_t* __restrict__ t,
    const int64_t* __restrict__ offs, int64_t n_segments, const double* __restrict__ home, double limit,
    double stay_dist, int64_t stay_min, double* __restrict__ rows, int64_t* __restrict__ labels) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    HavThreshold th_home, th_stay;
    th_home.set(limit);
    th_stay.set(stay_dist);
    for (int64_t seg = warp0; seg < n_segments; seg += nwarps) {
        const int64_t a = offs[seg], b = offs[seg + 1];
        const int64_t n = b - a;
        double* row = rows + seg * MHB_SEG_COLUMNS;
        if (n <= 0) {
            if (lane == 0) {
                row[0] = 0.0;
                for (int c = 1; c < MHB_SEG_COLUMNS; ++c) row[c] = CUDART_NAN;
                row[5] = row[7] = row[8] = 0.0;
            }
            continue;
        }
        const double* la = lat + a;
        const double* lo = lon + a;
        const double nd = static_cast<double>(n);
        // ---- pass A: centroid (np.mean: plain sums)
        double s1 = 0, s2 = 0;
        for (int64_t i = lane; i < n; i += 32) {
            s1 += la[i];
            s2 += lo[i];
        }
        const double mlat = warp_sum(s1) / nd, mlon = warp_sum(s2) / nd;
        // ---- pass B: variances, gyration, path length, home statistics.  Per point: ONE cosine (its own latitude,
        // shared by the three haversines it takes part in; the previous point's comes from the neighbouring lane),
        // six sines, and asin / sqrt only where the distance itself is needed (gyration, path length).
        const double hlat = home[2 * seg], hlon = home[2 * seg + 1];
        const double am = __dmul_rn(mlat, kDeg), om = __dmul_rn(mlon, kDeg), cm = cos_lat(am);
        const double ah = __dmul_rn(hlat, kDeg), oh = __dmul_rn(hlon, kDeg), ch = cos_lat(ah);
        double vlat = 0, vlon = 0, gyr = 0, path = 0, hmax = 0;
        unsigned long long near = 0;
        double carry_c = 0.0;                       // cos(latitude) of the last point of the previous 32-point step
        for (int64_t base = 0; base < n; base += 32) {
            const int64_t i = base + lane;
            const bool in = i < n;
            const double x = in ? la[i] : 0.0, y = in ? lo[i] : 0.0;
            const double ax = __dmul_rn(x, kDeg), oy = __dmul_rn(y, kDeg);
            const double cx = cos_lat(ax);
            double cprev = __shfl_up_sync(0xffffffffu, cx, 1);
            if (lane == 0) cprev = carry_c;
            carry_c = __shfl_sync(0xffffffffu, cx, 31);
            if (in) {
                const double dx = x - mlat, dy = y - mlon;
                vlat += dx * dx;
                vlon += dy * dy;
                const double dc = hav_km(hav_h(ax, oy, cx, am, om, cm));           // haversine(x, y, mlat, mlon)
                gyr += dc * dc;
                if (i > 0) {
                    const double ap = __dmul_rn(la[i - 1], kDeg), op = __dmul_rn(lo[i - 1], kDeg);
                    path += hav_km(hav_h(ap, op, cprev, ax, oy, cx));              // haversine(p[i-1], p[i])
                }
                const double hh = hav_h(ah, oh, ch, ax, oy, cx);                   // haversine_vector(home, points)
                hmax = fmax(hmax, hh);
                near += th_home.less(hh) ? 1ull : 0ull;                            // strict <, features.py:83-84
            }
        }
        vlat = warp_sum(vlat);
        vlon = warp_sum(vlon);
        gyr = warp_sum(gyr);
        path = warp_sum(path);
        near = warp_sum(near);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
        const double dmax = hav_km(hmax);           // the distance is monotone in h: max d = d(max h)

        // ---- pass C: stay points (anchor scan, 32 candidates per step) + label entropy on the fly
        const int64_t* tt = t + a;
        int64_t i = 0, n_stay = 0, n_noise = 0;
        double hsum = 0.0;                          // sum over stay labels of p ln p
        while (i < n) {
            const double aa = __dmul_rn(la[i], kDeg), ao = __dmul_rn(lo[i], kDeg), ac = cos_lat(aa);
            int64_t j = i + 1;
            while (j < n) {
                const int64_t q = j + lane;
                bool stop = q >= n;
                if (!stop) {
                    const double aq = __dmul_rn(la[q], kDeg), oq = __dmul_rn(lo[q], kDeg);
                    stop = th_stay.greater(hav_h(aa, ao, ac, aq, oq, cos_lat(aq)));    // haversine(anchor, p[q]) > stay_dist
                }
                const unsigned ball = __ballot_sync(0xffffffffu, stop);
                if (ball) {
                    j += __ffs(ball) - 1;
                    break;
                }
                j += 32;
            }
            if (j > n) j = n;
            const bool is_stay = (tt[j - 1] - tt[i]) >= stay_min;
            if (labels) {
                const int64_t lab = is_stay ? n_stay : -1;
                for (int64_t q = i + lane; q < j; q += 32) labels[a + q] = lab;
            }
            if (is_stay) {
                hsum += plogp(static_cast<double>(j - i), nd);
                ++n_stay;
            } else {
                n_noise += j - i;
            }
            i = j;
        }
        if (lane == 0) {
            const int64_t n_labels = n_stay + (n_noise > 0 ? 1 : 0);
            double h = hsum;
            if (n_noise > 0) h += plogp(static_cast<double>(n_noise), nd);
            h = -h;
            row[0] = nd;
            row[1] = path;
            row[2] = vlat / nd + vlon / nd;                         // distribution.py:39
            row[3] = sqrt(gyr / nd);
            row[4] = dmax;
            row[5] = static_cast<double>(near);
            row[6] = static_cast<double>(near) / nd;
            row[7] = static_cast<double>(n_stay);
            row[8] = static_cast<double>(n_labels);
            row[9] = h;
            row[10] = n_labels > 1 ? h / log(static_cast<double>(n_labels)) : CUDART_NAN;
        }
    }
}

// ---- label statistics (distribution.py:58-102): dense histogram over [lmin, lmax]
__global__ void zero_i64_kernel(int64_t* p, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        p[i] = 0;
}

__global__ void label_hist_kernel(const int64_t* __restrict__ labels, int64_t n, int64_t lmin, int64_t lmax,
                                  int64_t* __restrict__ counts) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t l = labels[i];
        if (l >= lmin && l <= lmax)
            atomicAdd(reinterpret_cast<unsigned long long*>(counts + (l - lmin)), 1ull);
    }
}

__global__ void __launch_bounds__(256) label_reduce_kernel(const int64_t* __restrict__ counts, int64_t range, int64_t n,
                                                           int64_t n_override, double* __restrict__ out3) {
    __shared__ double sh[8];
    __shared__ unsigned long long su[8];
    double h = 0.0;
    unsigned long long distinct = 0;
    const double nd = static_cast<double>(n);
    for (int64_t i = threadIdx.x; i < range; i += blockDim.x) {
        const int64_t c = counts[i];
        if (c > 0) {
            h += plogp(static_cast<double>(c), nd);
            ++distinct;
        }
    }
    h = warp_sum(h);
    distinct = warp_sum(distinct);
    if ((threadIdx.x & 31) == 0) {
        sh[threadIdx.x >> 5] = h;
        su[threadIdx.x >> 5] = distinct;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double H = 0.0;
        unsigned long long D = 0;
        for (int i = 0; i < 8; ++i) {
            H += sh[i];
            D += su[i];
        }
        H = -H;
        const double nc = n_override > 0 ? static_cast<double>(n_override) : static_cast<double>(D);
        out3[0] = static_cast<double>(D);
        out3[1] = H;
        out3[2] = H / log(nc);              // distribution.py:102 (0/0 -> nan for a single label, as in the reference)
    }
}

__global__ void minmax_init_kernel(const int64_t* v, int64_t* out2) {
    out2[0] = v[0];
    out2[1] = v[0];
}

__global__ void minmax_i64_kernel(const int64_t* __restrict__ v, int64_t n, int64_t* __restrict__ out2) {
    long long lo = v[0], hi = v[0];
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const long long x = v[i];
        lo = x < lo ? x : lo;
        hi = x > hi ? x : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<long long*>(out2), lo);
        atomicMax(reinterpret_cast<long long*>(out2 + 1), hi);
    }
}

inline unsigned grid_for(int64_t n, int threads = 256) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace mhb

using namespace mhb;

extern "C" int32_t mhb_haversine_elementwise(const double* lat1, const double* lon1, const double* lat2,
                                             const double* lon2, int64_t n, double* out, void* stream) {
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "haversine_elementwise: negative length");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(lat1 && lon1 && lat2 && lon2 && out, MHB_E_ARG, "haversine_elementwise: null pointer");
    haversine_elementwise_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(lat1, lon1, lat2, lon2, n, out);
    return cuda_status(cudaGetLastError(), "haversine_elementwise");
}

extern "C" int32_t mhb_haversine_vector(double lat, double lon, const double* latcol, const double* loncol, int64_t n,
                                        double* out, void* stream) {
    MHB_REQUIRE(n >= 0, MHB_E_ARG, "haversine_vector: negative length");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(latcol && loncol && out, MHB_E_ARG, "haversine_vector: null pointer");
    haversine_vector_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(lat, lon, latcol, loncol, n, out);
    return cuda_status(cudaGetLastError(), "haversine_vector");
}

extern "C" int32_t mhb_haversine_outer(const double* lat1, const double* lon1, int64_t n, const double* lat2,
                                       const double* lon2, int64_t m, double* out, void* stream) {
    MHB_REQUIRE(n >= 0 && m >= 0, MHB_E_ARG, "haversine_outer: negative length");
    if (n == 0 || m == 0) return MHB_OK;
    MHB_REQUIRE(lat1 && lon1 && lat2 && lon2 && out, MHB_E_ARG, "haversine_outer: null pointer");
    haversine_outer_kernel<<<grid_for(n * m), 256, 0, static_cast<cudaStream_t>(stream)>>>(lat1, lon1, n, lat2, lon2, m, out);
    return cuda_status(cudaGetLastError(), "haversine_outer");
}

extern "C" int32_t mhb_successive_distance(const double* lat, const double* lon, const int64_t* seg_offsets,
                                           int64_t n_segments, int64_t n, double* out, void* stream) {
    MHB_REQUIRE(n >= 0 && n_segments >= 0, MHB_E_ARG, "successive_distance: negative size");
    if (n == 0) return MHB_OK;
    MHB_REQUIRE(lat && lon && out, MHB_E_ARG, "successive_distance: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    successive_distance_kernel<<<grid_for(n), 256, 0, s>>>(lat, lon, n, out);
    if (seg_offsets && n_segments > 0) zero_segment_starts_kernel<<<grid_for(n_segments), 256, 0, s>>>(seg_offsets, n_segments, n, out);
    return cuda_status(cudaGetLastError(), "successive_distance");
}

extern "C" int32_t mhb_location_segments(const double* lat, const double* lon, const int64_t* t,
                                         const int64_t* seg_offsets, int64_t n_segments, const double* home,
                                         double home_limit_km, double stay_dist_km, int64_t stay_min_seconds,
                                         double* out_rows, int64_t* labels_out, void* stream) {
    MHB_REQUIRE(n_segments >= 0, MHB_E_ARG, "location_segments: negative segment count");
    if (n_segments == 0) return MHB_OK;
    MHB_REQUIRE(lat && lon && t && seg_offsets && home && out_rows, MHB_E_ARG, "location_segments: null pointer");
    int64_t ctas = (n_segments + 7) / 8;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    if (ctas > cap) ctas = cap;
    location_segments_kernel<<<static_cast<unsigned>(ctas), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        lat, lon, t, seg_offsets, n_segments, home, home_limit_km, stay_dist_km, stay_min_seconds, out_rows, labels_out);
    return cuda_status(cudaGetLastError(), "location_segments");
}

extern "C" int32_t mhb_label_stats(const int64_t* labels, int64_t n, int64_t label_min, int64_t label_max,
                                   int64_t n_clusters_override, int64_t* workspace_counts, double* out3, void* stream) {
    MHB_REQUIRE(n > 0 && label_max >= label_min, MHB_E_ARG, "label_stats: empty input or bad label range");
    MHB_REQUIRE(labels && workspace_counts && out3, MHB_E_ARG, "label_stats: null pointer");
    const int64_t range = label_max - label_min + 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    zero_i64_kernel<<<grid_for(range), 256, 0, s>>>(workspace_counts, range);
    label_hist_kernel<<<grid_for(n), 256, 0, s>>>(labels, n, label_min, label_max, workspace_counts);
    label_reduce_kernel<<<1, 256, 0, s>>>(workspace_counts, range, n, n_clusters_override, out3);
    return cuda_status(cudaGetLastError(), "label_stats");
}

extern "C" int32_t mhb_minmax_i64(const int64_t* v, int64_t n, int64_t* out2, void* stream) {
    MHB_REQUIRE(n > 0, MHB_E_ARG, "minmax_i64: empty input");
    MHB_REQUIRE(v && out2, MHB_E_ARG, "minmax_i64: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    minmax_init_kernel<<<1, 1, 0, s>>>(v, out2);
    minmax_i64_kernel<<<grid_for(n), 256, 0, s>>>(v, n, out2);
    return cuda_status(cudaGetLastError(), "minmax_i64");
}
