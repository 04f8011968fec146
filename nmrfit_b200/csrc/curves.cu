// K4/K5: curve evaluation on a grid - the pieces FitUtility.generate_result strings
// together on the host in the reference (utils.py:226-295), plus the standalone
// ps2 / voigt / Kramers-Kronig entry points that stay importable from Python.
//
// The Kramers-Kronig counterpart (reference equations.py:9-80: adaptive quadrature
// of [V(w-x) - V(w+x)]/x over [0, inf), ~6.6 ms per grid point on a CPU core) is
// evaluated in closed form: the Hilbert transform of the Lorentzian is the
// dispersion Lorentzian t/(1+t^2), that of the Gaussian is (2/sqrt(pi)) Dawson(s);
// yoff cancels in the integrand.
//
// generate_result is bound by HBM writes: (2P + 4) doubles out and 1 double in per
// grid point ((2*24+5)*8 B * 262,144 = 111 MB at BASELINE config 5).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include "nmrfit_internal.h"
#include "nmrfit_math.cuh"

namespace nmrfit {

// phi_i = p0 + (p1*i)/n  (proc_autophase.py:30-31); forward: (u + iv) e^{i phi}; inverse: (u + iv) e^{-i phi}
__global__ void ps2_kernel(const double* __restrict__ u, const double* __restrict__ v, int n, double p0, double p1,
                           int inv, double* __restrict__ re, double* __restrict__ im) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sn, cs;
    sincos(p0 + (p1 * (double)i) / (double)n, &sn, &cs);
    if (inv) sn = -sn;
    double a = u[i], b = v[i];
    re[i] = a * cs - b * sn;
    im[i] = a * sn + b * cs;
}

__device__ __forceinline__ double body_real(const PeakCoef& c, double w) {
    double d = w - c.loc, d2 = d * d;
    double rq = rcp_pos(fma(d2, c.kL2, 1.0));
    return fma(c.aG, exp_neg<0>(d2 * c.nkG2, nullptr), c.aL * rq);
}

__device__ __forceinline__ double body_imag(const PeakCoef& c, double kL, double kG, double w) {
    double d = w - c.loc;
    double rq = rcp_pos(fma(d * d, c.kL2, 1.0));
    double daw = dawson(d * kG, NMRFIT_DAW_TAB, NMRFIT_DAW_TAIL);
    return fma(c.aL * (d * kL), rq, (c.aG * kTwoOverSqrtPi) * daw);
}

__global__ void voigt_kernel(const double* __restrict__ w, int n, double r, double yoff, double width, double loc,
                             double a, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PeakCoef c = make_coef(r, width, loc, a);
    out[i] = yoff + body_real(c, w[i]);
}

__global__ void kk_kernel(const double* __restrict__ w, int n, double r, double width, double loc, double a,
                          double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PeakCoef c = make_coef(r, width, loc, a);
    double iw = 2.0 / width;
    out[i] = body_imag(c, iw, iw * kSqrtLn2, w[i]);
}

// A CTA of 256 threads covers kGenPoints consecutive grid points x kGenGroups groups of peaks: thread (x, y) evaluates
// the peaks k = y, y + kGenGroups, ... at point x.  Splitting the peaks over threads quarters every thread's serial
// chain and quadruples the warps in flight (262,144 points x 24 peaks is only one wave of one-thread-per-point CTAs:
// the first version spent its 34 us on latency, not on its 111 MB of stores).  The groups' partial sums of V and I
// meet in shared memory and are added in group order.  Every store is coalesced along the grid.  The D parameters
// travel as a kernel argument when they fit (up to kGenInlinePeaks peaks: no allocation, no copy, nothing to wait
// for on the host); larger fits read them from `params_dev`.
constexpr int kGenPoints = 64;
constexpr int kGenGroups = 4;
constexpr int kGenThreads = kGenPoints * kGenGroups;
constexpr int kGenMaxPeaks = 256;
constexpr int kGenInlinePeaks = 64;
// (the per-peak coefficients {loc, aL, aG, aG 2/sqrt(pi), kL, kG} are computed on the host and travel in the argument:
// every warp works on ONE peak at a time, so its lanes read them from the constant bank as broadcasts - no prologue in
// which 24 threads of each of 4,096 CTAs divide and everyone else waits at a barrier)
struct GenParams { double p0, p1, yoff, pad; double c[kGenInlinePeaks][6]; };

template <bool INLINE>
__global__ void __launch_bounds__(kGenThreads)
generate_result_kernel(const GenParams gp, const double* __restrict__ params_dev, int P, const double* __restrict__ w,
                       int n, double* __restrict__ real, double* __restrict__ imag, double* __restrict__ V,
                       double* __restrict__ I, double* __restrict__ u, double* __restrict__ v) {
    extern __shared__ __align__(16) double sm[];   // (!INLINE: [P][8], then) the groups' partial sums [2][kGenGroups][kGenPoints]
    const int tid = threadIdx.x, x = tid % kGenPoints, y = tid / kGenPoints;
    double p0, p1, yoff;
    if (INLINE) {
        p0 = gp.p0; p1 = gp.p1; yoff = gp.yoff;
    } else {
        p0 = params_dev[0]; p1 = params_dev[1]; yoff = params_dev[3];
        const double r = params_dev[2];
        for (int k = tid; k < P; k += kGenThreads) {
            double width = params_dev[4 + 3 * k], loc = params_dev[5 + 3 * k], a = params_dev[6 + 3 * k];
            PeakCoef c = make_coef(r, width, loc, a);
            double iw = 2.0 / width;
            double* o = sm + k * 8;
            o[0] = c.loc; o[1] = c.aL; o[2] = c.aG; o[3] = c.aG * kTwoOverSqrtPi; o[4] = iw; o[5] = iw * kSqrtLn2; o[6] = 0; o[7] = 0;
        }
        __syncthreads();
    }
    double* part = sm + (INLINE ? 0 : (size_t)P * 8);      // [2 buffers][2 sums][kGenGroups][kGenPoints]
    // (written as a grid-stride loop over chunks of kGenPoints points, one barrier per chunk with double-buffered partial
    // sums; launched with one chunk per CTA - see launch_generate_result)
    const int n_chunks = (n + kGenPoints - 1) / kGenPoints;
    int buf = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, buf ^= 1) {
        const int i = chunk * kGenPoints + x;
        const bool ok = i < n;
        const double wi = ok ? w[i] : 0.0;
        double vs = 0.0, is = 0.0;
        if (ok) {
            for (int k = y; k < P; k += kGenGroups) {
                // one reciprocal serves the Lorentzian and its dispersion counterpart; the Gaussian is evaluated only
                // within 6.5 units of s of the centre (beyond: < 4.5e-19 of its height, below half an ulp of anything it
                // is added to) - a warp's neighbouring points are on the same side of that cut almost everywhere, so the
                // branch is uniform; Dawson's integral keeps its full range (its tail decays only as 1/s)
                double2 c01, c23, c45;                     // (loc, aL), (aG, aG * 2/sqrt(pi)), (kL, kG)
                if (INLINE) {
                    c01 = make_double2(gp.c[k][0], gp.c[k][1]);
                    c23 = make_double2(gp.c[k][2], gp.c[k][3]);
                    c45 = make_double2(gp.c[k][4], gp.c[k][5]);
                } else {
                    c01 = *reinterpret_cast<const double2*>(sm + k * 8);
                    c23 = *reinterpret_cast<const double2*>(sm + k * 8 + 2);
                    c45 = *reinterpret_cast<const double2*>(sm + k * 8 + 4);
                }
                const double d = wi - c01.x;
                const double t = d * c45.x, s = d * c45.y;
                const double rq = rcp_pos(fma(t, t, 1.0));
                const double lor = c01.y * rq;
                double body = lor;
                if (fabs(s) <= kGaussCut) body = fma(c23.x, exp_neg<0>(-(s * s), nullptr), lor);
                const double re = yoff + body;       // utils.py:267: every contribution carries yoff
                const double im = fma(lor, t, c23.y * dawson(s, NMRFIT_DAW_TAB, NMRFIT_DAW_TAIL));
                real[(size_t)k * n + i] = re;
                imag[(size_t)k * n + i] = im;
                vs += re;                            // utils.py:276-277: both sums accumulate
                is += im;
            }
        }
        double* pb = part + buf * (2 * kGenGroups * kGenPoints);
        pb[y * kGenPoints + x] = vs;
        pb[(kGenGroups + y) * kGenPoints + x] = is;
        __syncthreads();
        if (y != 0 || !ok) continue;
        vs = 0.0;
        is = 0.0;
#pragma unroll
        for (int g = 0; g < kGenGroups; ++g) {
            vs += pb[g * kGenPoints + x];
            is += pb[(kGenGroups + g) * kGenPoints + x];
        }
        V[i] = vs;
        I[i] = is;
        double sn, cs;
        sincos(p0 + (p1 * (double)i) / (double)n, &sn, &cs);   // utils.py:284: ramp over the UPSAMPLED length
        u[i] = vs * cs + is * sn;
        v[i] = is * cs - vs * sn;
    }
}

cudaError_t launch_ps2(const double* u, const double* v, int n, double p0, double p1, int inv, double* re, double* im,
                       cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_launches(1);
    ps2_kernel<<<(n + 255) / 256, 256, 0, st>>>(u, v, n, p0, p1, inv, re, im);
    return cudaGetLastError();
}

cudaError_t launch_voigt(const double* w, int n, double r, double yoff, double width, double loc, double a, double* out,
                         cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_launches(1);
    voigt_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, n, r, yoff, width, loc, a, out);
    return cudaGetLastError();
}

cudaError_t launch_kk(const double* w, int n, double r, double width, double loc, double a, double* out,
                      cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_launches(1);
    kk_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, n, r, width, loc, a, out);
    return cudaGetLastError();
}

// params_host: the D = 4 + 3P fitted parameters in host memory
cudaError_t launch_generate_result(const double* params_host, int P, const double* w, int n, double* real,
                                   double* imag, double* V, double* I, double* u, double* v, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (P > kGenMaxPeaks) return cudaErrorInvalidValue;
    count_launches(1);
    const int D = 4 + 3 * P;
    // one chunk per CTA (the loop in the kernel takes any grid).  Measured, 24 peaks x 262,144 points: 28.7 us; a resident
    // set of 148 x 6 CTAs striding over the chunks 31.8 us, four chunks per CTA 33.8 us - the hardware's CTA scheduler and
    // neighbouring CTAs writing neighbouring memory do better than either
    const unsigned n_chunks = (n + kGenPoints - 1) / kGenPoints;
    const unsigned grid = n_chunks;
    if (P <= kGenInlinePeaks) {
        GenParams gp{};
        gp.p0 = params_host[0]; gp.p1 = params_host[1]; gp.yoff = params_host[3];
        for (int k = 0; k < P; ++k) {
            // make_coef's expressions, on the host (IEEE divisions and products: the same doubles)
            const double r = params_host[2], width = params_host[4 + 3 * k], a = params_host[6 + 3 * k];
            const double iw = 2.0 / width;
            const double aG = a * (1.0 - r) * (iw * kSqrtLn2OverPi);
            gp.c[k][0] = params_host[5 + 3 * k];
            gp.c[k][1] = a * r * (iw / kPi);
            gp.c[k][2] = aG;
            gp.c[k][3] = aG * kTwoOverSqrtPi;
            gp.c[k][4] = iw;
            gp.c[k][5] = iw * kSqrtLn2;
        }
        const size_t smem_inline = (size_t)2 * 2 * kGenGroups * kGenPoints * sizeof(double);
        generate_result_kernel<true><<<grid, kGenThreads, smem_inline, st>>>(gp, nullptr, P, w, n, real, imag, V, I, u, v);
        return cudaGetLastError();
    }
    const size_t smem = ((size_t)P * 8 + 2 * 2 * kGenGroups * kGenPoints) * sizeof(double);
    double* pd = nullptr;
    cudaError_t e = cudaMallocAsync(&pd, sizeof(double) * D, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(pd, params_host, sizeof(double) * D, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        GenParams gp{};
        generate_result_kernel<false><<<grid, kGenThreads, smem, st>>>(gp, pd, P, w, n, real, imag, V, I, u, v);
        e = cudaGetLastError();
    }
    cudaFreeAsync(pd, st);
    return e;
}

}  // namespace nmrfit
