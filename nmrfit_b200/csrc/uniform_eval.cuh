// The two device routines of the uniform-axis objective, shared by the two-pass kernels
// (objective_uniform.cu) and the fused swarm kernel (swarm_fused.cu) so that both evaluate a particle
// with the SAME arithmetic in the SAME order - their objective values are bit-identical.
//
//   prepare_particle  per-particle constants: span coefficients of its peaks, phase tables and, per
//                     region of 32*R points, the near-peak mask, the far-field polynomial and the
//                     phase anchor.  The destinations are plain pointers: global memory in the
//                     two-pass path, shared memory in the fused kernel.
//   eval_region       one warp's squared weighted residual over its region, from those constants
//                     and the staged (u, v, weights) of the tile.
#pragma once
#include <cuda_runtime.h>
#include "nmrfit_math.cuh"
#include "uniform_common.cuh"

namespace nmrfit {

// xs: the particle's D parameters; sw: its spectrum's w plane (N points); cs: shared [P][8] scratch that ends up
// holding the span coefficients; coef_out: optional second copy of them; part [kPartDoubles]; far [NRP][kFarTerms];
// anchor [NRP][2]; mask [NRP][MW+1].  NR regions cover the axis, NRP >= NR slots are filled (the rest neutral).
// Called by all `nthreads` (>= 128) threads of a CTA; contains __syncthreads().
// `pairs` (optional shared scratch of NRP*P*kPairDoubles doubles, only when NRP*P + NRP <= nthreads): the
// (region, peak) series are computed one pair per thread instead of one region per thread - same values, same order
// of accumulation, but P times shorter on the critical path (the fused swarm kernel has one CTA per particle).
constexpr int kPairDoubles = kFarTerms + 1;
template <int R>
__device__ __forceinline__ void prepare_particle(const double* __restrict__ xs, const double* __restrict__ sw, double h,
                                                 double w_ulp, int N, int P, int NR, int NRP, int tid, int nthreads,
                                                 double* __restrict__ cs, double* __restrict__ coef_out,
                                                 double* __restrict__ part, double* __restrict__ far,
                                                 double* __restrict__ anchor, unsigned* __restrict__ mask,
                                                 double* __restrict__ pairs = nullptr, int r_lo = 0, int r_hi = -1,
                                                 int slot_nw = 1 << 30, size_t slot_stride = 0) {
    // regions [r_lo, r_hi) are filled, at index r - r_lo of far / anchor / mask (default: all NRP slots).
    // Tile-major destinations (the streamed evaluation kernel reads one tile's regions of a whole particle group as
    // one block): local region rl lands at slot (rl / slot_nw) * slot_stride + rl % slot_nw instead of rl.
    if (r_hi < 0) r_hi = NRP;
    const int nr = r_hi - r_lo;
    const int MW = (P + 31) / 32;
    const double p0 = xs[0], p1 = xs[1];
    constexpr double H = 16.0 * R;

    for (int k = tid; k < P; k += nthreads) {
        SpanCoef c = make_span_coef(xs[2], xs[4 + 3 * k], xs[5 + 3 * k], xs[6 + 3 * k], h, w_ulp, R);
        if (c.exact) c = null_span_coef();                 // the span loop adds zero; the peak is handled after it
        double* o = cs + k * 8;
        o[0] = c.loc; o[1] = c.kL; o[2] = c.kG; o[3] = c.aL; o[4] = c.aG; o[5] = c.dT; o[6] = c.thr; o[7] = c.c2;
        if (coef_out) {
            double* g = coef_out + k * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = o[i];
        }
    }
    // phi_i = p0 + (p1*i)/N with i = i_r + lane*R + j:  anchor(i_r) * e^{i p1 (lane R)/N} * (e^{i p1/N})^j
    for (int e = nthreads - 1 - tid; e < 33; e += nthreads) {   // the last warps do these while the first does the peaks
        double sn, cn;
        sincos(e < 32 ? (p1 * (double)(e * R)) / (double)N : p1 / (double)N, &sn, &cn);
        part[2 * e] = cn;
        part[2 * e + 1] = sn;
    }
    if (tid == 64) part[66] = (double)P * xs[3];           // yoff is added once per peak (equations.py:147,195)
    __syncthreads();
    if (tid == 64) {
        int n = 0;
        for (int k = 0; k < P; ++k) n += cs[k * 8 + 6] < 0.0;      // thr < 0 marks a nulled (exact-path) peak
        part[67] = (double)n;
    }
    if (pairs) {
        // one (region, peak) pair per thread; the region anchors by the threads at the other end
        if (tid < nr * P) {
            const int rl = tid / P, k = tid - rl * P, r = r_lo + rl;
            double* pr = pairs + (size_t)tid * kPairDoubles;
            double kind = -1.0;                            // -1: padding region or exact-path peak (neither near nor far)
            if (r < NR) {
                const double* o = cs + k * 8;
                SpanCoef c;
                c.loc = o[0]; c.kL = o[1]; c.kG = o[2]; c.aL = o[3]; c.aG = o[4]; c.dT = o[5]; c.thr = o[6]; c.c2 = o[7];
                if (!(c.thr < 0.0)) {
                    const double w_c = fma(0.5 * (32 * R - 1), h, sw[r * 32 * R]);
                    double v[kFarTerms];
                    kind = (double)far_terms(w_c - c.loc, c, H, v);
#pragma unroll
                    for (int n = 0; n < kFarTerms; ++n) pr[1 + n] = v[n];
                }
            }
            pr[0] = kind;
        }
        const int ral = nthreads - 1 - tid;
        if (ral < nr) {
            const int ra = r_lo + ral;
            double sn = 0.0, cn = 1.0;
            if (ra < NR) sincos(p0 + (p1 * (double)(ra * 32 * R)) / (double)N, &sn, &cn);
            anchor[ral * 2] = cn;
            anchor[ral * 2 + 1] = sn;
        }
        __syncthreads();
        if (tid < nr) {
            const int r = tid;                             // local index
            double C[kFarTerms];
#pragma unroll
            for (int n = 0; n < kFarTerms; ++n) C[n] = 0.0;
            unsigned any_far = 0;
            unsigned* mk = mask + (size_t)r * (MW + 1);
            for (int wd = 0; wd < MW; ++wd) {
                unsigned m = 0;
                const int kend = min(P, wd * 32 + 32);
                for (int k = wd * 32; k < kend; ++k) {
                    const double* pr = pairs + (size_t)(r * P + k) * kPairDoubles;
                    const int kind = (int)pr[0];
                    if (kind < 0) continue;
                    if (kind == kFarNear) { m |= 1u << (k & 31); continue; }
                    any_far = 1u;
                    if (kind == kFarSeries) {
                        double v[kFarTerms];
#pragma unroll
                        for (int n = 0; n < kFarTerms; ++n) v[n] = pr[1 + n];
                        far_add(cs[k * 8 + 3], v, C);
                    }
                }
                mk[wd] = m;
            }
            mk[MW] = any_far;
            double* fc = far + (size_t)r * kFarTerms;
#pragma unroll
            for (int n = 0; n < kFarTerms; ++n) fc[n] = C[n];
        }
        return;
    }
    for (int rl = tid; rl < nr; rl += nthreads) {
        const int r = r_lo + rl;
        const size_t slot = (size_t)(rl / slot_nw) * slot_stride + (size_t)(rl % slot_nw);
        double C[kFarTerms];
#pragma unroll
        for (int n = 0; n < kFarTerms; ++n) C[n] = 0.0;
        unsigned any_far = 0;
        unsigned* mk = mask + slot * (MW + 1);
        double sn = 0.0, cn = 1.0;
        if (r < NR) {
            const int ir = r * 32 * R;
            const double w_c = fma(0.5 * (32 * R - 1), h, sw[ir]);
            for (int wd = 0; wd < MW; ++wd) {
                unsigned m = 0;
                const int kend = min(P, wd * 32 + 32);
                for (int k = wd * 32; k < kend; ++k) {
                    const double* o = cs + k * 8;
                    SpanCoef c;
                    c.loc = o[0]; c.kL = o[1]; c.kG = o[2]; c.aL = o[3]; c.aG = o[4]; c.dT = o[5]; c.thr = o[6]; c.c2 = o[7];
                    if (c.thr < 0.0) continue;             // exact-path peak: neither near nor far
                    if (far_accumulate(w_c - c.loc, c, H, C)) any_far = 1u;
                    else m |= 1u << (k & 31);
                }
                mk[wd] = m;
            }
            sincos(p0 + (p1 * (double)ir) / (double)N, &sn, &cn);
        } else {
            for (int wd = 0; wd < MW; ++wd) mk[wd] = 0u;
        }
        mk[MW] = any_far;
        double* fc = far + slot * kFarTerms;
#pragma unroll
        for (int n = 0; n < kFarTerms; ++n) fc[n] = C[n];
        anchor[slot * 2] = cn;
        anchor[slot * 2 + 1] = sn;
    }
}

// Shared-memory placement of a tile's staged points.  Point e of the tile (thread t = e / R, j = e % R) lives in
// row j at column t ^ j (double2 (u, v) array) resp. t ^ 2j (weights array): the evaluation reads row j with
// consecutive t (a permutation inside each aligned group of 16 columns: conflict-free), and the staging loop,
// whose consecutive lanes hold consecutive e - eight different rows of the same column - spreads over eight
// different banks instead of colliding 8-way (ncu before: 18.3 M of 66.6 M shared wavefronts were such conflicts).
__device__ __forceinline__ int stage_slot_uv(int t, int j, int stride) { return j * stride + (t ^ j); }
__device__ __forceinline__ int stage_slot_wt(int t, int j, int stride) { return j * stride + (t ^ (2 * j)); }

// One warp, one particle, one region: sum over the warp's 32*R points of (weights * (V_data - V_fit))^2, identical
// in every lane on return.  cf [P][8], pt [kPartDoubles], mk [MW+1], fc [kFarTerms] (16-byte aligned) and ew are the
// particle's constants for this region; the R points of thread t of the tile sit at stage_slot_uv/wt(t, j, stride) of
// suv / swt; w_first is the abscissa of its first point and xi0 that point's position inside the region.  The exact path
// (peaks too narrow for the recurrences) reads the particle's parameters xs and the stored abscissae sw_first[0..n_valid).
// KK = 1 (fit_im, reference semantics) also returns through *ss_im the same sum for the imaginary parts:
// I_data = u sin(phi) + v cos(phi) against the last peak's Kramers-Kronig counterpart (closed form, nmrfit_math.cuh).
template <int R, int TB, int KK = 0>
__device__ __forceinline__ double eval_region(const double* __restrict__ cf, const double* __restrict__ pt,
                                              const unsigned* __restrict__ mk, const double* __restrict__ fc,
                                              const double2 ew, int MW, int P, int lane, double w_first, double xi0,
                                              const double2* __restrict__ suv, const double* __restrict__ swt, int t,
                                              int stride, const double* __restrict__ tab,
                                              const double* __restrict__ xs, const double* __restrict__ sw_first,
                                              int n_valid, double h, double w_ulp, double* ss_im = nullptr) {
    constexpr double H = 16.0 * R;                         // half a region, in points
    double acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.0;
    for (int wd = 0; wd < MW; ++wd)
    for (unsigned m = mk[wd]; m; m &= m - 1) {             // peaks near this warp's region
        const int k = wd * 32 + __ffs(m) - 1;
        const double2 c01 = *reinterpret_cast<const double2*>(cf + k * 8);
        const double2 c23 = *reinterpret_cast<const double2*>(cf + k * 8 + 2);
        const double2 c45 = *reinterpret_cast<const double2*>(cf + k * 8 + 4);
        const double2 c67 = *reinterpret_cast<const double2*>(cf + k * 8 + 6);
        SpanCoef c;
        c.loc = c01.x; c.kL = c01.y; c.kG = c23.x; c.aL = c23.y;
        c.aG = c45.x; c.dT = c45.y; c.thr = c67.x; c.c2 = c67.y;
        peak_span<R, TB>(w_first - c.loc, c, tab, acc);
    }
    if (mk[MW]) {                                          // all far peaks at once
        double C[kFarTerms];
#pragma unroll
        for (int n = 0; n < kFarTerms; n += 2) {
            const double2 t = *reinterpret_cast<const double2*>(fc + n);
            C[n] = t.x; C[n + 1] = t.y;
        }
        far_eval<R>(C, xi0, 1.0 / H, acc);
    }
    if (pt[67] != 0.0) {                                   // rare: peaks too narrow for the uniform-axis shortcuts
        for (int k = 0; k < P; ++k) {
            if (!(cf[k * 8 + 6] < 0.0)) continue;
            const SpanCoef c = make_span_coef(xs[2], xs[4 + 3 * k], xs[5 + 3 * k], xs[6 + 3 * k], h, w_ulp, R);
            peak_exact<R, TB>(sw_first, n_valid, w_first, h, c, tab, acc);
        }
    }
    // residual against the phase-rotated data; the rotation advances by p1/N per point
    const double2 el = *reinterpret_cast<const double2*>(pt + 2 * lane);
    const double cd = pt[64], sd = pt[65], py = pt[66];
    double cr = fma(ew.x, el.x, -(ew.y * el.y));
    double ci = fma(ew.y, el.x, ew.x * el.y);
    double ss = 0.0, ssi = 0.0;
    // fit_im is True: I_fit is OVERWRITTEN peak by peak (equations.py:198-199), so only the LAST peak's Kramers-Kronig
    // curve is ever compared with the data.  Its constants are its span coefficients - rebuilt from the parameters
    // when the peak is on the exact path (the shared copy is nulled then).
    double kloc = 0.0, kkL = 0.0, kkG = 0.0, kaL = 0.0, kaG = 0.0;
    bool kexact = false;
    if (KK) {
        const double* cl = cf + (P - 1) * 8;
        kexact = cl[6] < 0.0;
        if (kexact) {
            const int k = P - 1;
            const SpanCoef c = make_span_coef(xs[2], xs[4 + 3 * k], xs[5 + 3 * k], xs[6 + 3 * k], h, w_ulp, R);
            kloc = c.loc; kkL = c.kL; kkG = c.kG; kaL = c.aL; kaG = c.aG * kTwoOverSqrtPi;
        } else {
            kloc = cl[0]; kkL = cl[1]; kkG = cl[2]; kaL = cl[3]; kaG = cl[4] * kTwoOverSqrtPi;
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const double2 uv = suv[stage_slot_uv(t, j, stride)];
        const double wt = swt[stage_slot_wt(t, j, stride)];
        const double vd = fma(uv.x, cr, -fma(uv.y, ci, py));     // V_data - P*yoff
        const double res = wt * (vd - acc[j]);
        ss = fma(res, res, ss);
        if (KK) {
            const double idat = fma(uv.x, ci, uv.y * cr);
            // abscissa: w_first + j*h on the uniform axis; the stored value when the peak is too narrow for that
            const double wj = (kexact && j < n_valid) ? sw_first[j] : (j == 0 ? w_first : fma((double)j, h, w_first));
            const double d = wj - kloc;
            const double tt = d * kkL;
            const double rq = rcp_pos(fma(tt, tt, 1.0));
            const double daw = dawson(d * kkG, NMRFIT_DAW_TAB, NMRFIT_DAW_TAIL);
            const double ifit = fma(kaL * tt, rq, kaG * daw);
            const double ri = wt * (idat - ifit);
            ssi = fma(ri, ri, ssi);
        }
        if (j + 1 < R) {
            const double c2 = fma(cr, cd, -(ci * sd));
            ci = fma(ci, cd, cr * sd);
            cr = c2;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (KK) ssi += __shfl_xor_sync(0xffffffffu, ssi, o);
    }
    if (KK) *ss_im = ssi;
    return ss;
}

}  // namespace nmrfit
