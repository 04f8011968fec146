// K9: phase estimation before the fit, for a batch of spectra and a list of phase candidates.
//
// Reference: Data.shift_phase(method='brute'|'auto') (containers.py:51-78).
//   brute  Data._brute_phase (containers.py:98-110): for each p0 in arange(-pi, pi, step) phase the spectrum
//          (ps2 with p1 = 0), error = sqrt((mean(V[:n]) - mean(V[-n:]))^2) with n = max(1, N/5000); keep the first
//          smallest error among candidates whose real part points up (max V > |min V|).
//   auto   proc_autophase.approximate_phase (proc_autophase.py:107-139): scipy Nelder-Mead on the ACME score
//          (_ps_acme_score, proc_autophase.py:142-187) - entropy of the normalised |dV| + 1000 x squared negative
//          excursions.  The simplex stays on the host (scipy, as in the reference); the score is evaluated here.
//
// One CTA per (candidate, spectrum): threads stride over the N points (coalesced u, v reads: 16 B per point per
// candidate, served by L2 after the first candidate), V_i = u_i cos(phi_i) - v_i sin(phi_i) with
// phi_i = p0 + (p1*i)/N (proc_autophase.py:30-31), fixed-shape reductions (no atomics: reproducible).
// The reference computes e^{i phi} with the host libm; CUDA's sincos differs from it by <= 1 ulp, so scores and
// errors agree to ~1e-15 relative rather than bit for bit - the tests state the tolerance.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "nmrfit_internal.h"

namespace nmrfit {

constexpr int kPhThreads = 256;

__device__ __forceinline__ double block_sum(double x, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < kPhThreads / 32; ++k) t += sm[k];
    return t;
}
__device__ __forceinline__ double block_max(double x, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    __syncthreads();
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    double t = sm[0];
    for (int k = 1; k < kPhThreads / 32; ++k) t = fmax(t, sm[k]);
    return t;
}

// np.sum of n <= 128 doubles at(off), at(off + 1), ... as numpy computes it: pairwise summation's base case - eight
// running sums over blocks of eight, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail.
template <typename F>
__device__ __forceinline__ double numpy_sum_block(F at, int off, int n) {
    double s;
    if (n < 8) {
        s = 0.0;
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, at(off + i));
    } else {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = at(off + k);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], at(off + i + k));
        s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                      __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) s = __dadd_rn(s, at(off + i));
    }
    return s;
}

// np.mean of ANY number of doubles as numpy computes it (V[:n].mean() of containers.py:103-104 with n = N/5000, i.e.
// n > 128 from 645,000 points on): above 128 elements pairwise summation splits at n/2 rounded down to a multiple of
// eight and adds the halves' sums; the recursion is unrolled on a small explicit stack (depth <= log2(n/128) + 1).
template <typename F>
__device__ __forceinline__ double numpy_mean(F at, int n) {
    if (n <= 128) return numpy_sum_block(at, 0, n) / (double)n;
    int off[32], len[32], state[32];
    double left[32];
    int sp = 0;
    off[0] = 0; len[0] = n; state[0] = 0; left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        if (len[sp] <= 128) {
            ret = numpy_sum_block(at, off[sp], len[sp]);
            --sp;
            continue;
        }
        int n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (state[sp] == 0) {                              // descend into the left half
            state[sp] = 1;
            off[sp + 1] = off[sp]; len[sp + 1] = n2; state[sp + 1] = 0;
            ++sp;
        } else if (state[sp] == 1) {                       // left half done: keep it, descend into the right half
            left[sp] = ret;
            state[sp] = 2;
            off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; state[sp + 1] = 0;
            ++sp;
        } else {                                           // both done
            ret = __dadd_rn(left[sp], ret);
            --sp;
        }
    }
    return ret / (double)n;
}

// brute scan: err[b][k], ok[b][k] for candidate p0_k (p1 = 0)
__global__ void __launch_bounds__(kPhThreads)
phase_brute_kernel(const double* __restrict__ u, const double* __restrict__ v, int N, const double* __restrict__ cands,
                   int K, double* __restrict__ err, int* __restrict__ ok) {
    __shared__ double sm[kPhThreads / 32];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double* ub = u + (size_t)b * N;
    const double* vb = v + (size_t)b * N;
    double sn, cs;
    sincos(cands[k], &sn, &cs);                            // p1 = 0: one rotation for the whole spectrum
    double hi = -CUDART_INF, lo = CUDART_INF;
    for (int i = tid; i < N; i += kPhThreads) {
        const double V = __dsub_rn(__dmul_rn(ub[i], cs), __dmul_rn(vb[i], sn));   // Re[(u + iv)(c + is)], numpy's order
        hi = fmax(hi, V);
        lo = fmin(lo, V);
    }
    hi = block_max(hi, sm);
    lo = -block_max(-lo, sm);
    if (tid == 0) {
        const int n = max(1, N / 5000);
        auto head = [&](int i) { return __dsub_rn(__dmul_rn(ub[i], cs), __dmul_rn(vb[i], sn)); };
        auto tail = [&](int i) { const int j = N - n + i; return __dsub_rn(__dmul_rn(ub[j], cs), __dmul_rn(vb[j], sn)); };
        const double d = __dsub_rn(numpy_mean(head, n), numpy_mean(tail, n));
        err[(size_t)b * K + k] = sqrt(__dmul_rn(d, d));
        ok[(size_t)b * K + k] = hi > fabs(lo);
    }
}

// first smallest error among the upward candidates (containers.py:105-108); none -> p0 = 0, error = inf
__global__ void phase_brute_select_kernel(const double* __restrict__ cands, int K, const double* __restrict__ err,
                                          const int* __restrict__ ok, double* __restrict__ best_p0,
                                          double* __restrict__ best_err) {
    const int b = blockIdx.x;
    double be = CUDART_INF, bp = 0.0;
    for (int k = 0; k < K; ++k) {
        const double e = err[(size_t)b * K + k];
        if (e < be && ok[(size_t)b * K + k]) { be = e; bp = cands[k]; }
    }
    best_p0[b] = bp;
    if (best_err) best_err[b] = be;
}

// ACME score[b][k] for candidate (p0_k, p1_k) in RADIANS
__global__ void __launch_bounds__(kPhThreads)
phase_acme_kernel(const double* __restrict__ u, const double* __restrict__ v, int N, const double* __restrict__ ph, int K,
                  double* __restrict__ score) {
    __shared__ double sm[kPhThreads / 32];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double* ub = u + (size_t)b * N;
    const double* vb = v + (size_t)b * N;
    const double p0 = ph[2 * k], p1 = ph[2 * k + 1];
    auto real_at = [&](int i) {
        double sn, cs;
        sincos(p0 + (p1 * (double)i) / (double)N, &sn, &cs);
        return __dsub_rn(__dmul_rn(ub[i], cs), __dmul_rn(vb[i], sn));
    };
    // pass 1: T = sum |dV|/2, negative excursions
    double T = 0.0, sumas = 0.0, pen = 0.0;
    for (int i = tid; i < N; i += kPhThreads) {
        const double V = real_at(i);
        if (i + 1 < N) T += fabs((real_at(i + 1) - V) / 2.0);
        const double as_ = V - fabs(V);
        sumas += as_;
        pen += (as_ / 2.0) * (as_ / 2.0);
    }
    T = block_sum(T, sm);
    sumas = block_sum(sumas, sm);
    pen = block_sum(pen, sm);
    // pass 2: entropy -sum p log p with p = ds/T (zeros contribute nothing: proc_autophase.py:171)
    double h = 0.0;
    for (int i = tid; i + 1 < N; i += kPhThreads) {
        const double p = fabs((real_at(i + 1) - real_at(i)) / 2.0) / T;
        if (p != 0.0) h += -p * log(p);
    }
    h = block_sum(h, sm);
    if (tid == 0) score[(size_t)b * K + k] = h + 1000.0 * (sumas < 0.0 ? pen : 0.0);
}

cudaError_t launch_phase_brute(const double* u, const double* v, int B, int N, const double* cands_dev, int K,
                               double* err, int* ok, double* best_p0, double* best_err, cudaStream_t st) {
    phase_brute_kernel<<<dim3(K, B), kPhThreads, 0, st>>>(u, v, N, cands_dev, K, err, ok);
    phase_brute_select_kernel<<<B, 1, 0, st>>>(cands_dev, K, err, ok, best_p0, best_err);
    count_launches(2);
    return cudaGetLastError();
}

cudaError_t launch_phase_acme(const double* u, const double* v, int B, int N, const double* ph_dev, int K, double* score,
                              cudaStream_t st) {
    phase_acme_kernel<<<dim3(K, B), kPhThreads, 0, st>>>(u, v, N, ph_dev, K, score);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace nmrfit
