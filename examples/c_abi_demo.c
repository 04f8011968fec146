/* Plain-C consumer of libnmrfit_b200.so: no Python, no torch.  Builds a synthetic two-peak spectrum, evaluates
 * the objective for a few parameter vectors on the GPU, checks them against a straightforward C restatement of
 * equations.objective (equations.py:152-212, real-only), then runs a small swarm through nmrfit_pso_* and prints
 * the fitted parameters - once on one context and once sharded over TWO contexts of this process through
 * nmrfit_comm_* (the library-level counterpart of the reference's processes=N; both contexts sit on GPU 0 here, a
 * multi-GPU consumer creates one per device), which must give the identical result.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -ldl -lm
 *   ./c_abi_demo nmrfit_b200/csrc/libnmrfit_b200.so
 *
 * Exit code 0 = parity within 1e-11 and the swarm improved on its starting point.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "nmrfit_b200.h"

#define NPTS 1500
#define NPEAKS 2
#define ND (4 + 3 * NPEAKS)

static double voigt(double w, double r, double width, double loc, double a) {
    const double pi = 3.14159265358979323846, ln2 = 0.69314718055994530942;
    double t = (w - loc) / (0.5 * width);
    double lor = (2.0 / (pi * width)) / (1.0 + t * t);
    double s = (w - loc) / (width / (2.0 * sqrt(ln2)));
    double gau = (2.0 / width) * sqrt(ln2 / pi) * exp(-s * s);
    return a * (r * lor + (1.0 - r) * gau);
}

static double objective_c(const double* x, const double* w, const double* u, const double* v, const double* wt) {
    double ss = 0.0;
    for (int i = 0; i < NPTS; ++i) {
        double phi = x[0] + (x[1] * (double)i) / (double)NPTS;
        double vd = u[i] * cos(phi) - v[i] * sin(phi);
        double fit = 0.0;
        for (int k = 0; k < NPEAKS; ++k) fit += x[3] + voigt(w[i], x[2], x[4 + 3 * k], x[5 + 3 * k], x[6 + 3 * k]);
        double res = wt[i] * (vd - fit);
        ss += res * res;
    }
    return sqrt(ss / NPTS);
}

#define LOAD(name) do { *(void**)(&p_##name) = dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; } } while (0)

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "nmrfit_b200/csrc/libnmrfit_b200.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int (*p_nmrfit_ctx_create)(nmrfit_ctx**, int, int, int, int, int);
    void (*p_nmrfit_ctx_destroy)(nmrfit_ctx*);
    int (*p_nmrfit_ctx_set_spectrum)(nmrfit_ctx*, int, const double*, const double*, const double*, const double*);
    int (*p_nmrfit_objective_batch_host)(nmrfit_ctx*, const double*, int, int, double*);
    int (*p_nmrfit_pso_begin)(nmrfit_ctx*, const double*, const double*, const nmrfit_pso_opts*, const double*, const double*, void*);
    int (*p_nmrfit_pso_commit)(nmrfit_ctx*, const double*, int, void*);
    int (*p_nmrfit_pso_run)(nmrfit_ctx*, int, const double*, const double*, int*, void*);
    int (*p_nmrfit_pso_get_best)(nmrfit_ctx*, double*, double*, int*, int*);
    const char* (*p_nmrfit_last_error)(void);
    int (*p_nmrfit_comm_init_all)(nmrfit_ctx* const*, int);
    int (*p_nmrfit_comm_commit)(nmrfit_ctx* const*, int);
    int (*p_nmrfit_comm_run)(nmrfit_ctx* const*, int, int, const double* const*, const double* const*, int*, int*);
    int (*p_nmrfit_ctx_set_fused)(nmrfit_ctx*, int);
    LOAD(nmrfit_comm_init_all); LOAD(nmrfit_comm_commit); LOAD(nmrfit_comm_run); LOAD(nmrfit_ctx_set_fused);
    LOAD(nmrfit_ctx_create); LOAD(nmrfit_ctx_destroy); LOAD(nmrfit_ctx_set_spectrum); LOAD(nmrfit_objective_batch_host);
    LOAD(nmrfit_pso_begin); LOAD(nmrfit_pso_commit); LOAD(nmrfit_pso_run); LOAD(nmrfit_pso_get_best); LOAD(nmrfit_last_error);

    /* synthetic spectrum: two peaks, phased by (0.2, 0.03), unit weights */
    static double w[NPTS], u[NPTS], v[NPTS], wt[NPTS];
    const double truth[ND] = {0.2, 0.03, 0.6, 0.0, 0.004, 3.35, 0.02, 0.005, 3.45, 0.01};
    for (int i = 0; i < NPTS; ++i) {
        w[i] = 3.23 + (3.60 - 3.23) * (double)i / (double)(NPTS - 1);
        double V = 0.0;
        for (int k = 0; k < NPEAKS; ++k) V += voigt(w[i], truth[2], truth[4 + 3 * k], truth[5 + 3 * k], truth[6 + 3 * k]);
        double phi = truth[0] + (truth[1] * (double)i) / (double)NPTS;
        u[i] = V * cos(phi);              /* inverse rotation of (V, 0) */
        v[i] = -V * sin(phi);
        wt[i] = 1.0;
    }
    nmrfit_ctx* ctx = NULL;
    if (p_nmrfit_ctx_create(&ctx, 0, 1, NPTS, NPEAKS, NMRFIT_FP64)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    if (p_nmrfit_ctx_set_spectrum(ctx, 0, w, u, v, wt)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }

    /* objective parity on three parameter vectors */
    double xs[3][ND], f[3];
    for (int s = 0; s < 3; ++s)
        for (int d = 0; d < ND; ++d) xs[s][d] = truth[d] * (1.0 + 0.05 * s * ((d % 2) ? 1 : -1));
    if (p_nmrfit_objective_batch_host(ctx, &xs[0][0], 3, NMRFIT_REAL_ONLY, f)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    double worst = 0.0;
    for (int s = 1; s < 3; ++s) {      /* s = 0 is the truth: residual ~1e-17, relative error meaningless */
        double want = objective_c(xs[s], w, u, v, wt);
        double rel = fabs(f[s] / want - 1.0);
        if (rel > worst) worst = rel;
        printf("objective[%d] gpu %.15e  c %.15e  rel %.2e\n", s, f[s], want, rel);
    }
    printf("objective[0] (truth) gpu %.3e\n", f[0]);

    /* a small swarm with the reference's constants, device random numbers */
    double lb[ND], ub[ND];
    for (int d = 0; d < ND; ++d) { lb[d] = truth[d] - 0.5 * fabs(truth[d]) - 1e-3; ub[d] = truth[d] + 0.5 * fabs(truth[d]) + 1e-3; }
    lb[2] = 0.0; ub[2] = 1.0;
    nmrfit_pso_opts o;
    o.swarmsize = 64; o.maxiter = 150; o.omega = -0.2134; o.phip = -0.3344; o.phig = 2.3259;
    o.minstep = 1e-8; o.minfunc = 1e-8; o.fit_im = NMRFIT_REAL_ONLY; o.bounds_per_spectrum = 0; o.seed = 5; o.particle_offset = 0;
    o.spectrum_offset = 0;
    double x0[ND], f0, x1[ND], f1;
    int gens = 0, stop = 0, running = 1;
    if (p_nmrfit_pso_begin(ctx, lb, ub, &o, NULL, NULL, NULL) || p_nmrfit_pso_commit(ctx, NULL, 1, NULL) ||
        p_nmrfit_pso_get_best(ctx, x0, &f0, NULL, NULL)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    for (int done = 0; running && done < o.maxiter; done += 50)
        if (p_nmrfit_pso_run(ctx, 50, NULL, NULL, &running, NULL)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    if (p_nmrfit_pso_get_best(ctx, x1, &f1, &gens, &stop)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    printf("swarm: f %.3e -> %.3e in %d generations (stop %d); width0 %.5f loc0 %.5f area0 %.5f\n", f0, f1, gens, stop,
           x1[4], x1[5], x1[6]);
    p_nmrfit_ctx_destroy(ctx);

    /* the same swarm, particles split 40 + 24 over two contexts (device random numbers are keyed by the global particle
     * index, so the split does not change the trajectory); reference run: per-step kernels on one context */
    double xa[ND], fa, xb[ND], fb, xr[ND], fr;
    int lost = 0, same = 1;
    {
        nmrfit_ctx* one = NULL;
        if (p_nmrfit_ctx_create(&one, 0, 1, NPTS, NPEAKS, NMRFIT_FP64) || p_nmrfit_ctx_set_spectrum(one, 0, w, u, v, wt) ||
            p_nmrfit_ctx_set_fused(one, NMRFIT_FUSED_OFF) || p_nmrfit_pso_begin(one, lb, ub, &o, NULL, NULL, NULL) ||
            p_nmrfit_pso_commit(one, NULL, 1, NULL) || p_nmrfit_pso_run(one, 60, NULL, NULL, &running, NULL) ||
            p_nmrfit_pso_get_best(one, xr, &fr, NULL, NULL)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
        p_nmrfit_ctx_destroy(one);
    }
    nmrfit_ctx* pair[2] = {NULL, NULL};
    const int count[2] = {40, 24};
    for (int r = 0; r < 2; ++r)
        if (p_nmrfit_ctx_create(&pair[r], 0, 1, NPTS, NPEAKS, NMRFIT_FP64) || p_nmrfit_ctx_set_spectrum(pair[r], 0, w, u, v, wt) ||
            p_nmrfit_ctx_set_fused(pair[r], NMRFIT_FUSED_OFF)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    if (p_nmrfit_comm_init_all(pair, 2)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    for (int r = 0; r < 2; ++r) {
        nmrfit_pso_opts os = o;
        os.swarmsize = count[r];
        os.particle_offset = r ? count[0] : 0;
        if (p_nmrfit_pso_begin(pair[r], lb, ub, &os, NULL, NULL, NULL)) { fprintf(stderr, "%s\n", p_nmrfit_last_error()); return 3; }
    }
    if (p_nmrfit_comm_commit(pair, 2) || p_nmrfit_comm_run(pair, 2, 60, NULL, NULL, &running, &lost) ||
        p_nmrfit_pso_get_best(pair[0], xa, &fa, NULL, NULL) || p_nmrfit_pso_get_best(pair[1], xb, &fb, NULL, NULL)) {
        fprintf(stderr, "%s\n", p_nmrfit_last_error());
        return 3;
    }
    for (int d = 0; d < ND; ++d) same = same && xa[d] == xb[d] && xa[d] == xr[d];
    same = same && fa == fb && fa == fr && !lost;
    printf("communicator: f %.15e on both contexts, one context %.15e: %s\n", fa, fr, same ? "identical" : "DIFFERENT");
    p_nmrfit_ctx_destroy(pair[0]);
    p_nmrfit_ctx_destroy(pair[1]);
    dlclose(lib);
    if (!same) { fprintf(stderr, "the sharded swarm differs from the unsharded one\n"); return 1; }
    if (!(worst < 1e-11)) { fprintf(stderr, "parity %.3e exceeds 1e-11\n", worst); return 1; }
    if (!(f1 < f0)) { fprintf(stderr, "swarm did not improve\n"); return 1; }
    return 0;
}
