// K10: automatic peak selection for a batch of spectra - the reference's AutoPeakSelector (nmrfit/utils.py:670-783,
// reached through Data.select_peaks('auto'), containers.py:159-161), the step that turns a phased spectrum into the
// Peak records (loc, width, height, area, bounds) from which the fit's bounds and weights are built.
//
// Per spectrum the reference (a) upsamples u 100x by linear interpolation onto np.linspace(w.min(), w.max(), 100 N)
// (utils.py:711-714), (b) smooths with a Savitzky-Golay filter (11 points, degree 4; :716), (c) takes a constant
// baseline by peakutils.baseline(.., 0) (:718) - clip to the running mean until it moves by < 0.1 %, at most 100 times -
// (d) keeps as maxima the samples strictly above every neighbour within `window` ppm (scipy.signal.argrelmax with an
// order of tens of thousands of samples, O(M * order) in scipy; :728-738), and per maximum (e) finds the nearest
// falling and rising half-height crossings (:748-753), (f) a local baseline over loc +- 2 widths (:760-767) and
// (g) the Simpson area over that window (:770).  All of it is data-parallel over the upsampled samples:
//
//   peaks_upsample_kernel   (a) one thread per upsampled sample: searchsorted + the interp1d expression
//   peaks_smooth_kernel     (b) interior by the symmetric correlation ndimage evaluates, 5 + 5 edge samples by the
//                               degree-4 least-squares fit of the first / last 11 samples
//   peaks_baseline_kernel   (c) one CTA per spectrum; an iteration is one reduction of min(y, running minimum of means)
//   peaks_blockmax_kernel + peaks_maxima_kernel   (d) candidates (above both neighbours) check their window through
//                               per-1,024-sample maxima: a few hundred loads each instead of 2 * order
//   peaks_cross_kernel      (e) one CTA per (peak, spectrum): closest crossing of either kind, lowest index on ties
//   peaks_measure_kernel    (f, g) one CTA per (peak, spectrum): local baseline iterations, then composite Simpson on
//                               the (irregular to rounding) abscissae as scipy.integrate.simpson evaluates it
//
// HBM-bound integer / compare work with a few flops per sample; the upsampled signal (800 B per input point) stays
// in L2 for one spectrum and streams from HBM for a batch.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "nmrfit_internal.h"

namespace nmrfit {

namespace {

constexpr int kPkThreads = 256;
constexpr int kPkBlock = 1024;                             // samples per block maximum

// np.linspace(start, stop, M)[j]: arange(M) * step + start with the last sample set to stop
__device__ __forceinline__ double wu_at(long long j, long long M, double start, double stop, double step) {
    return j == M - 1 ? stop : __dadd_rn(__dmul_rn((double)j, step), start);
}

__device__ __forceinline__ double block_sum(double x, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nw; ++k) t += sm[k];
    return t;
}

// (a) uu[b][i] = interp1d(w, u)(wu_i): scipy.interpolate.interp1d linear evaluation -
//     hi = clip(searchsorted(x, x_new, 'left'), 1, N-1); lo = hi - 1; slope = (y_hi - y_lo) / (x_hi - x_lo);
//     y_new = slope * (x_new - x_lo) + y_lo      (every operation rounded on its own)
__global__ void __launch_bounds__(kPkThreads)
peaks_upsample_kernel(const double* __restrict__ w, const double* __restrict__ u, int N, long long M, double* __restrict__ uu) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPkThreads + threadIdx.x;
    if (i >= M) return;
    const double* wb = w + (size_t)b * N;
    const double* ub = u + (size_t)b * N;
    const double start = wb[0], stop = wb[N - 1];
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(M - 1));
    const double x = wu_at(i, M, start, stop, step);
    int lo = 0, hi = N;                                    // first index with w[idx] >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (wb[mid] < x) lo = mid + 1; else hi = mid;
    }
    int ih = lo < 1 ? 1 : (lo > N - 1 ? N - 1 : lo);
    const int il = ih - 1;
    const double slope = __ddiv_rn(__dsub_rn(ub[ih], ub[il]), __dsub_rn(wb[ih], wb[il]));
    uu[(size_t)b * M + i] = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, wb[il])), ub[il]);
}

// (b) Savitzky-Golay(11, 4), mode='interp'.  c[0..10]: scipy.signal.savgol_coeffs(11, 4); interior samples as
// ndimage.correlate1d evaluates a symmetric kernel: x_i c_5 + sum_{m=5..1} c_{5+m} (x_{i-m} + x_{i+m});
// edge[0][r][k], edge[1][r][k]: the linear maps of the degree-4 polynomial fit of the first / last 11 samples
// evaluated at positions 0..4 / 6..10 (scipy's _fit_edges_polyfit).
struct SgCoef { double c[11]; double edge[2][5][11]; };
__global__ void __launch_bounds__(kPkThreads)
peaks_smooth_kernel(const double* __restrict__ uu, long long M, SgCoef k, double* __restrict__ us) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPkThreads + threadIdx.x;
    if (i >= M) return;
    const double* x = uu + (size_t)b * M;
    double out;
    if (M < 11) {
        out = x[i];                                        // (the reference would raise; callers never get here)
    } else if (i < 5 || i >= M - 5) {
        const int side = i < 5 ? 0 : 1;
        const int r = side == 0 ? (int)i : (int)(i - (M - 5));
        const double* e = k.edge[side][r];
        const double* xe = side == 0 ? x : x + (M - 11);
        double t = 0.0;
        for (int q = 0; q < 11; ++q) t = fma(e[q], xe[q], t);
        out = t;
    } else {
        double t = __dmul_rn(x[i], k.c[5]);
#pragma unroll
        for (int m = 5; m >= 1; --m) t = __dadd_rn(t, __dmul_rn(k.c[5 + m], __dadd_rn(x[i - m], x[i + m])));
        out = t;
    }
    us[(size_t)b * M + i] = out;
}

// (c) peakutils.baseline(y, 0)[0]: coefficient c = 1; repeat <= max_it times: c' = mean(y); stop if |c' - c| / |c| < tol;
// c = c'; y = min(y, c).  y after k updates is min(y_0, smallest accepted mean), so nothing is written back.
// On the very first test succeeding peakutils returns the signal itself: [0] is then y_0[0].
__device__ __forceinline__ double baseline0(const double* __restrict__ y, long long n, int max_it, double tol, double* sm) {
    double c = 1.0, cmin = CUDART_INF;
    bool first = true;
    for (int it = 0; it < max_it; ++it) {
        double s = 0.0;
        for (long long j = threadIdx.x; j < n; j += blockDim.x) s += fmin(y[j], cmin);
        const double c_new = block_sum(s, sm) / (double)n;
        if (fabs(c_new - c) / fabs(c) < tol) return first ? y[0] : c;
        c = c_new;
        cmin = fmin(cmin, c);
        first = false;
    }
    return c;
}

__global__ void __launch_bounds__(1024)
peaks_baseline_kernel(const double* __restrict__ us, long long M, int max_it, double tol, double* __restrict__ base) {
    __shared__ double sm[32];
    const int b = blockIdx.x;
    const double v = baseline0(us + (size_t)b * M, M, max_it, tol, sm);
    if (threadIdx.x == 0) base[b] = v;
}

// (d) maxima of every block of kPkBlock samples
__global__ void __launch_bounds__(kPkThreads)
peaks_blockmax_kernel(const double* __restrict__ us, long long M, int nblk, double* __restrict__ bmax) {
    __shared__ double sm[kPkThreads / 32];
    const int b = blockIdx.y, blk = blockIdx.x;
    const double* x = us + (size_t)b * M;
    double m = -CUDART_INF;
    for (int e = threadIdx.x; e < kPkBlock; e += kPkThreads) {
        const long long j = (long long)blk * kPkBlock + e;
        if (j < M) m = fmax(m, x[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kPkThreads / 32; ++k) m = fmax(m, sm[k]);
        bmax[(size_t)b * nblk + blk] = m;
    }
}

// scipy.signal.argrelmax(x, order)[0], mode 'clip': x_i > x_j for every j within `order` of i (clipped to the ends;
// an end sample is compared with itself and never qualifies).  One thread per sample; the few that beat both
// neighbours check their window: whole blocks through bmax, the two partial blocks sample by sample.
__global__ void __launch_bounds__(kPkThreads)
peaks_maxima_kernel(const double* __restrict__ us, const double* __restrict__ bmax, long long M, int nblk, long long order,
                    int max_out, long long* __restrict__ out, int* __restrict__ count) {
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * kPkThreads + threadIdx.x;
    if (i < 1 || i >= M - 1) return;
    const double* x = us + (size_t)b * M;
    const double xi = x[i];
    if (!(xi > x[i - 1] && xi > x[i + 1])) return;
    const long long lo = max(0LL, i - order), hi = min(M - 1, i + order);
    const double* bm = bmax + (size_t)b * nblk;
    // blocks entirely inside [lo, hi] that do not hold i
    const long long b_lo = (lo + kPkBlock - 1) / kPkBlock, b_hi = (hi + 1) / kPkBlock;     // [b_lo, b_hi)
    const long long bi = i / kPkBlock;
    for (long long q = b_lo; q < b_hi; ++q)
        if (q != bi && !(bm[q] < xi)) return;
    auto scan = [&](long long s, long long e) {            // samples [s, e] except i
        for (long long j = s; j <= e; ++j)
            if (j != i && !(x[j] < xi)) return false;
        return true;
    };
    if (b_lo >= b_hi) {                                    // no whole block inside: the window itself
        if (!scan(lo, hi)) return;
    } else {
        if (!scan(lo, b_lo * kPkBlock - 1)) return;
        if (!scan(b_hi * kPkBlock, hi)) return;
        if (bi >= b_lo && bi < b_hi && !scan(bi * kPkBlock, min(M - 1, (bi + 1) * kPkBlock - 1))) return;
    }
    const int slot = atomicAdd(count + b, 1);
    if (slot < max_out) out[(size_t)b * max_out + slot] = i;
}

// gather: v[b][k] = uu[b][idx[b][k]]
__global__ void peaks_gather_kernel(const double* __restrict__ uu, long long M, const long long* __restrict__ idx, int max_out,
                                    const int* __restrict__ count, double* __restrict__ v) {
    const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= max_out || k >= count[b]) return;
    v[(size_t)b * max_out + k] = uu[(size_t)b * M + idx[(size_t)b * max_out + k]];
}

// (e) nearest falling (d < 0) and rising (d > 0) half-height crossing of peak k of spectrum b:
//     d_j = sign(h/2 - (uu_j - base)) - sign(h/2 - (uu_{j+1} - base)), j = 0..M-2;
//     x_right = wu[argmin over {d < 0} of |wu_j - loc|], x_left likewise over {d > 0} - the first index on ties.
struct PeakIn { long long i; double height; };
__device__ __forceinline__ int sgn(double a) { return (a > 0.0) - (a < 0.0); }
__global__ void __launch_bounds__(kPkThreads)
peaks_cross_kernel(const double* __restrict__ uu, const double* __restrict__ w, int N, long long M,
                   const double* __restrict__ base, const long long* __restrict__ pk_i, const double* __restrict__ pk_h,
                   const int* __restrict__ n_pk, int max_peaks, long long* __restrict__ cross /*[B][max][2]: right, left*/) {
    __shared__ double sd[2][kPkThreads / 32];
    __shared__ long long sj[2][kPkThreads / 32];
    const int b = blockIdx.y, k = blockIdx.x;
    if (k >= n_pk[b]) return;
    const double* y = uu + (size_t)b * M;
    const double start = w[(size_t)b * N], stop = w[(size_t)b * N + N - 1];
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(M - 1));
    const long long pi = pk_i[(size_t)b * max_peaks + k];
    const double loc = wu_at(pi, M, start, stop, step);
    const double half = pk_h[(size_t)b * max_peaks + k] / 2.0;
    const double bs = base[b];
    double bd[2] = {CUDART_INF, CUDART_INF};
    long long bj[2] = {-1, -1};
    for (long long j = threadIdx.x; j < M - 1; j += kPkThreads) {
        const int d = sgn(half - (y[j] - bs)) - sgn(half - (y[j + 1] - bs));
        if (d == 0) continue;
        const int which = d < 0 ? 0 : 1;
        const double dist = fabs(wu_at(j, M, start, stop, step) - loc);
        if (dist < bd[which]) { bd[which] = dist; bj[which] = j; }     // ascending j per thread: strict keeps the first
    }
    for (int which = 0; which < 2; ++which) {
        double d0 = bd[which];
        long long j0 = bj[which];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d0, o);
            const long long oj = __shfl_xor_sync(0xffffffffu, j0, o);
            if (oj >= 0 && (j0 < 0 || od < d0 || (od == d0 && oj < j0))) { d0 = od; j0 = oj; }
        }
        if ((threadIdx.x & 31) == 0) { sd[which][threadIdx.x >> 5] = d0; sj[which][threadIdx.x >> 5] = j0; }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int which = threadIdx.x;
        double d0 = sd[which][0];
        long long j0 = sj[which][0];
        for (int q = 1; q < kPkThreads / 32; ++q) {
            const double od = sd[which][q];
            const long long oj = sj[which][q];
            if (oj >= 0 && (j0 < 0 || od < d0 || (od == d0 && oj < j0))) { d0 = od; j0 = oj; }
        }
        cross[((size_t)b * max_peaks + k) * 2 + which] = j0;
    }
}

// (f, g) for an accepted peak: bounds = loc -+ 2 width; the samples with bounds[0] <= wu <= bounds[1]; local baseline;
// height = uu[i] - baseline; area = scipy.integrate.simpson(uu[idx] - baseline, x = wu[idx]).
struct PeakOut { double loc, width, b0, b1, baseline, height, area; long long lo, hi; int ok; int pad; };
__global__ void __launch_bounds__(kPkThreads)
peaks_measure_kernel(const double* __restrict__ uu, const double* __restrict__ w, int N, long long M,
                     const long long* __restrict__ pk_i, const int* __restrict__ n_pk, int max_peaks,
                     const long long* __restrict__ cross, int max_it, double tol, PeakOut* __restrict__ out) {
    __shared__ double sm[kPkThreads / 32];
    const int b = blockIdx.y, k = blockIdx.x;
    if (k >= n_pk[b]) return;
    const double* y = uu + (size_t)b * M;
    const double start = w[(size_t)b * N], stop = w[(size_t)b * N + N - 1];
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(M - 1));
    auto X = [&](long long j) { return wu_at(j, M, start, stop, step); };
    PeakOut& o = out[(size_t)b * max_peaks + k];
    const long long pi = pk_i[(size_t)b * max_peaks + k];
    const long long jr = cross[((size_t)b * max_peaks + k) * 2], jl = cross[((size_t)b * max_peaks + k) * 2 + 1];
    const double loc = X(pi);
    if (jr < 0 || jl < 0) {                                // no crossing of one kind: the reference raises here
        if (threadIdx.x == 0) { o.loc = loc; o.ok = -1; }
        return;
    }
    const double xr = X(jr), xl = X(jl);
    if (!(xl < xr)) {                                      // utils.py:755: the peak is dropped
        if (threadIdx.x == 0) { o.loc = loc; o.ok = 0; }
        return;
    }
    const double width = __dsub_rn(xr, xl);
    const double b0 = __dsub_rn(loc, __dmul_rn(2.0, width)), b1 = __dadd_rn(loc, __dmul_rn(2.0, width));
    // first sample >= b0 and last sample <= b1 (wu is non-decreasing)
    long long lo = 0, hi = M;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (X(mid) < b0) lo = mid + 1; else hi = mid; }
    const long long first = lo;
    lo = 0; hi = M;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (X(mid) <= b1) lo = mid + 1; else hi = mid; }
    const long long last = lo - 1;
    const long long n = last - first + 1;
    const double bl = baseline0(y + first, n, max_it, tol, sm);
    // composite Simpson on irregular abscissae (scipy.integrate._quadrature._basic_simpson / simpson, scipy >= 1.11):
    // pairs of intervals (2q, 2q+1, 2q+2), q = 0..(n_simpson - 3)/2, over the first n (odd) or n - 1 (even) samples
    const long long ns = (n & 1) ? n : n - 1;
    double acc = 0.0;
    for (long long q = threadIdx.x; 2 * q + 2 < ns; q += kPkThreads) {
        const long long j = first + 2 * q;
        const double x0 = X(j), x1 = X(j + 1), x2 = X(j + 2);
        const double h0 = x1 - x0, h1 = x2 - x1;
        const double hsum = h0 + h1, hprod = h0 * h1;
        const double h0divh1 = h1 != 0.0 ? h0 / h1 : 0.0;
        const double t0 = 2.0 - (h0divh1 != 0.0 ? 1.0 / h0divh1 : 0.0);
        const double t1 = hsum * (hprod != 0.0 ? hsum / hprod : 0.0);
        const double t2 = 2.0 - h0divh1;
        acc += hsum / 6.0 * ((y[j] - bl) * t0 + (y[j + 1] - bl) * t1 + (y[j + 2] - bl) * t2);
    }
    double area = block_sum(acc, sm);
    if (threadIdx.x == 0) {
        if (n == 2) {
            area = 0.5 * (X(last) - X(last - 1)) * ((y[last] - bl) + (y[last - 1] - bl));
        } else if (n >= 4 && (n & 1) == 0) {               // even count: Cartwright's correction for the last interval
            const double h0 = X(last - 1) - X(last - 2), h1 = X(last) - X(last - 1);
            double num = 2.0 * h1 * h1 + 3.0 * h0 * h1, den = 6.0 * (h1 + h0);
            const double alpha = den != 0.0 ? num / den : 0.0;
            num = h1 * h1 + 3.0 * h0 * h1; den = 6.0 * h0;
            const double beta = den != 0.0 ? num / den : 0.0;
            num = h1 * h1 * h1; den = 6.0 * h0 * (h0 + h1);
            const double eta = den != 0.0 ? num / den : 0.0;
            area += alpha * (y[last] - bl) + beta * (y[last - 1] - bl) - eta * (y[last - 2] - bl);
        }
        o.loc = loc; o.width = width; o.b0 = b0; o.b1 = b1; o.baseline = bl; o.height = y[pi] - bl; o.area = area;
        o.lo = first; o.hi = last; o.ok = 1;
    }
}

__global__ void peaks_probe_kernel(const double* __restrict__ w, int N, long long M, const double* __restrict__ uu,
                                   const double* __restrict__ us, int b, const long long* __restrict__ idx, int n,
                                   double* __restrict__ out /*[3][n]*/) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double start = w[(size_t)b * N], stop = w[(size_t)b * N + N - 1];
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(M - 1));
    const long long j = idx[k];
    out[k] = wu_at(j, M, start, stop, step);
    out[n + k] = uu[(size_t)b * M + j];
    out[2 * n + k] = us[(size_t)b * M + j];
}

}  // namespace

// ---- launchers (called from cabi.cu) --------------------------------------------------------------------------
cudaError_t launch_peaks_front(const double* w, const double* u, int B, int N, long long M, const double* sg /*[11+110]*/,
                               double* uu, double* us, double* bmax, int nblk, double* base, int max_it, double tol,
                               long long order, int max_out, long long* maxima, int* n_maxima, double* uu_at_maxima,
                               cudaStream_t st) {
    SgCoef k;
    for (int i = 0; i < 11; ++i) k.c[i] = sg[i];
    for (int s = 0; s < 2; ++s)
        for (int r = 0; r < 5; ++r)
            for (int q = 0; q < 11; ++q) k.edge[s][r][q] = sg[11 + (s * 5 + r) * 11 + q];
    const dim3 grid((unsigned)((M + kPkThreads - 1) / kPkThreads), B);
    peaks_upsample_kernel<<<grid, kPkThreads, 0, st>>>(w, u, N, M, uu);
    peaks_smooth_kernel<<<grid, kPkThreads, 0, st>>>(uu, M, k, us);
    peaks_baseline_kernel<<<B, 1024, 0, st>>>(us, M, max_it, tol, base);
    peaks_blockmax_kernel<<<dim3(nblk, B), kPkThreads, 0, st>>>(us, M, nblk, bmax);
    cudaError_t e = cudaMemsetAsync(n_maxima, 0, sizeof(int) * B, st);
    if (e != cudaSuccess) return e;
    peaks_maxima_kernel<<<grid, kPkThreads, 0, st>>>(us, bmax, M, nblk, order, max_out, maxima, n_maxima);
    peaks_gather_kernel<<<dim3((max_out + 127) / 128, B), 128, 0, st>>>(uu, M, maxima, max_out, n_maxima, uu_at_maxima);
    count_launches(6);
    return cudaGetLastError();
}

cudaError_t launch_peaks_back(const double* w, const double* uu, int B, int N, long long M, const double* base,
                              const long long* pk_i, const double* pk_h, const int* n_pk, int max_peaks, long long* cross,
                              int max_it, double tol, void* out, cudaStream_t st) {
    const dim3 grid(max_peaks, B);
    peaks_cross_kernel<<<grid, kPkThreads, 0, st>>>(uu, w, N, M, base, pk_i, pk_h, n_pk, max_peaks, cross);
    peaks_measure_kernel<<<grid, kPkThreads, 0, st>>>(uu, w, N, M, pk_i, n_pk, max_peaks, cross, max_it, tol,
                                                     reinterpret_cast<PeakOut*>(out));
    count_launches(2);
    return cudaGetLastError();
}

cudaError_t launch_peaks_probe(const double* w, int N, long long M, const double* uu, const double* us, int b,
                               const long long* idx, int n, double* out, cudaStream_t st) {
    peaks_probe_kernel<<<(n + 127) / 128, 128, 0, st>>>(w, N, M, uu, us, b, idx, n, out);
    count_launches(1);
    return cudaGetLastError();
}

size_t peaks_out_bytes() { return sizeof(PeakOut); }

}  // namespace nmrfit
