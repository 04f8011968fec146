// Host build of nmrfit_b200/csrc/nmrfit_math.cuh for accuracy tests without a GPU.
// Compiled by tests/test_host_math.py with g++ into a shared object; the product
// never links or loads it.
#define NMRFIT_HOST_MATH 1
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#include "../nmrfit_b200/csrc/nmrfit_math.cuh"

extern "C" {

void h_exp_neg(int tb, const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        switch (tb) {
            case 0: out[i] = nmrfit::exp_neg<0>(x[i], nullptr); break;
            case 6: out[i] = nmrfit::exp_neg<6>(x[i], NMRFIT_EXP2_TAB6); break;
            case 8: out[i] = nmrfit::exp_neg<8>(x[i], NMRFIT_EXP2_TAB8); break;
            default: out[i] = nmrfit::exp_neg<10>(x[i], NMRFIT_EXP2_TAB10); break;
        }
    }
}

void h_rcp_pos(const double* q, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = nmrfit::rcp_pos(q[i]);
}

void h_dawson(const double* s, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = nmrfit::dawson(s[i], NMRFIT_DAW_TAB, NMRFIT_DAW_TAIL);
}

// body of one peak at the points w (same association as the kernels use)
void h_voigt_body(const double* w, int n, double r, double width, double loc, double a, int tb, double* out) {
    nmrfit::PeakCoef c = nmrfit::make_coef(r, width, loc, a);
    for (int i = 0; i < n; ++i) {
        double d = w[i] - c.loc, d2 = d * d;
        double q = NMRFIT_FMA(d2, c.kL2, 1.0);
        double acc = c.aL * nmrfit::rcp_pos(q);
        double e = tb == 0 ? nmrfit::exp_neg<0>(d2 * c.nkG2, nullptr) : nmrfit::exp_neg<6>(d2 * c.nkG2, NMRFIT_EXP2_TAB6);
        out[i] = NMRFIT_FMA(c.aG, e, acc);
    }
}

void h_philox(unsigned long long seed, unsigned long long lo, unsigned long long hi, double* out2) {
    nmrfit::Philox2 p = nmrfit::philox_uniform2(seed, lo, hi);
    out2[0] = p.a; out2[1] = p.b;
}

}
