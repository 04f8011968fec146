// K7: swarm generations fused into ONE cooperative launch - the fit loop of a small swarm.
//
// pyswarm.pso (called from the reference at utils.py:176-182) is a host loop: move every particle, call the
// objective once per particle, update personal and swarm bests, test for convergence, repeat.  The per-step
// kernels (pso.cu + objective_uniform.cu) already keep that loop on the device, but a generation is still
// three launches, and for the swarm sizes a single nmrfit.fit uses (100-204 particles, 4k-16k points) every one
// of them is shorter than its own launch latency: ~25 us per generation for ~1 us of arithmetic.
//
// Here one CTA owns one particle for the whole run of generations.  Its state (x, v, p, fp, and replicas of the
// swarm best g, fg and of the box) lives in shared memory; when the spectrum fits it is staged in shared memory
// once per launch.  A generation is: move -> per-particle constants -> objective over every region of the axis ->
// personal best -> publish (fp, p) -> ONE barrier among the CTAs of the same spectrum -> every CTA finds the
// swarm's argmin and applies pyswarm's update/stop rules redundantly (identical inputs, identical result), so
// no second barrier and no broadcast are needed.  Published records are double-buffered by generation parity:
// a CTA can run at most one generation ahead of the slowest one.
//
// The arithmetic is the per-step kernels' own (uniform_eval.cuh, swarm_common.cuh) and the squared residual is
// summed in the same order (warp tree, then the warps of a point tile, then the tiles), so a fused run is
// bit-identical to the same generations stepped kernel by kernel (tests/test_gpu_fused.py).
//
// Longer axes (more than 16 regions of 256 points): a THREAD-BLOCK CLUSTER of up to 8 CTAs owns a particle.  Each
// CTA keeps its contiguous run of the spectrum in its own shared memory, evaluates its regions, and the regions'
// sums are all-gathered through DISTRIBUTED SHARED MEMORY (every CTA stores its sums into every peer's array,
// cluster.sync()), after which all CTAs of the cluster continue redundantly exactly as the single CTA does - same
// summation order, so still bit-identical to the per-step kernels.  An 80-register build of the cluster variant
// (3 CTAs per SM, a few spills) is used only when the 2-per-SM build cannot make the clusters co-resident.
//
// Requires every CTA of the grid to be co-resident: launched cooperatively (cudaLaunchKernelEx with the cooperative
// attribute, plus the cluster dimension), and the host falls back to the per-step kernels when the device cannot
// hold n_spectra * swarmsize * cluster CTAs or a CTA's run of the axis would exceed 32 regions.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <algorithm>
#include <cstdint>
#include "nmrfit_internal.h"
#include "nmrfit_math.cuh"
#include "uniform_common.cuh"
#include "uniform_eval.cuh"
#include "swarm_common.cuh"

namespace nmrfit {

namespace {

// shared-memory carve-up in doubles; every offset is even (16-byte alignment)
struct FusedSmem {
    int tab, uv, wt, cs, farpk, part, far, anchor, mask, wpart, state, red, misc, pairs, total;
    // NRP: region slots of the whole axis (tile sums of every region end up in every CTA); NRL: regions this CTA owns
    __host__ __device__ FusedSmem(int P, int D, int threads, int R, int slots, int NRP, int NRL, bool want_pairs, int sub) {
        const int De = (D + 1) & ~1;
        int o = 0;
        tab = o;    o += 64;
        uv = o;     o += slots * threads * R * 2;
        wt = o;     o += slots * threads * R;
        cs = o;     o += P * 8;
        farpk = o;  o += P * 4;
        part = o;   o += kPartDoubles;
        far = o;    o += NRL * sub * kFarPoly;              // per far-field cell (uniform_eval.cuh)
        anchor = o; o += NRL * 2;
        mask = o;   o += ((NRL * mask_words_per_region(P, sub) + 3) / 4) * 2;
        wpart = o;  o += (NRP + 1) & ~1;
        state = o;  o += 9 * De;         // x, v, p, g, lb, ub, best_x, p_min, spare
        red = o;    o += 64;             // per-warp argmin values and indices
        misc = o;   o += 8;
        // (cell, peak) series scratch of the pair-parallel prepare (one pair per thread, in rounds)
        pairs = want_pairs ? o : -1;
        if (pairs >= 0) o += ((NRL * sub * P * kPairDoubles + 1) & ~1);
        total = o;
    }
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// first-occurrence argmin rule of np.argmin on (value, index) pairs
__device__ __forceinline__ void take_min(double& bf, int& bi, double f, int i) {
    if (f < bf || (f == bf && i < bi)) { bf = f; bi = i; }
}

}  // namespace

// Optional phase timing (FusedArgs::timing != null): CTA 0's thread 0 adds the SM clock cycles each phase of a
// generation took - move, constants, objective, tile sums, publish, barrier, argmin, commit.
#define FUSED_MARK(slot)                                                        \
    do {                                                                        \
        if (a.timing && blockIdx.x == 0 && tid == 0) {                          \
            const long long now_ = clock64();                                   \
            a.timing[slot] += now_ - tmark;                                     \
            tmark = now_;                                                       \
        }                                                                       \
    } while (0)

template <int THREADS, int R, int TB, bool CL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
swarm_fused_kernel(FusedArgs a) {
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) double smem[];
    const SwarmState& s = a.s;
    const int S = s.S, D = s.D, De = (D + 1) & ~1;
    // a.cluster CTAs (one thread-block cluster) share a particle: each owns a contiguous run of supertiles
    const int G = CL ? a.cluster : 1, crank = CL ? (int)(blockIdx.x % G) : 0;
    const int pidx = blockIdx.x / G;
    const int b = pidx / S, sl = pidx % S;
    if (s.stop[b]) return;                                 // this spectrum's swarm has already stopped (whole cluster)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, N = a.N, MW = (P + 31) / 32;
    const int VW = a.vw, NRP = a.n_vtiles * VW;            // warps per point tile of the per-step kernels; region slots
    const int NR = (N + 32 * R - 1) / (32 * R);
    const int n_super = (NRP + NW - 1) / NW;
    const int per = (n_super + G - 1) / G;                 // supertiles per CTA
    const int st_lo = min(crank * per, n_super), st_hi = min(st_lo + per, n_super);
    const int r_lo = st_lo * NW, r_hi = min(st_hi * NW, NRP);
    const bool resident = a.slots >= per;
    const int SUB = a.sub;
    const FusedSmem L(P, D, THREADS, R, a.slots, NRP, per * NW, a.pairs != 0, SUB);
    double* tab = smem + L.tab;
    double2* suv = reinterpret_cast<double2*>(smem + L.uv);
    double* swt = smem + L.wt;
    double* cs = smem + L.cs;
    double* farpk = smem + L.farpk;
    double* part = smem + L.part;
    double* farc = smem + L.far;
    double* anchor = smem + L.anchor;
    unsigned* mask = reinterpret_cast<unsigned*>(smem + L.mask);
    double* wpart = smem + L.wpart;
    double* xs = smem + L.state;
    double* vs = xs + De;
    double* ps = vs + De;
    double* gs = ps + De;
    double* lbs = gs + De;
    double* ubs = lbs + De;
    double* bxs = ubs + De;                                // what pso() returns (p_min on an early stop)
    double* pm = bxs + De;                                 // the generation's best personal-best position
    double* redf = smem + L.red;
    int* redi = reinterpret_cast<int*>(smem + L.red + 32);
    double* misc = smem + L.misc;                          // 0 fx, 1 fp, 2 fg, 3 best_f, 4 improved, 5 action, 6 fmin
    double* pairs = L.pairs >= 0 ? smem + L.pairs : nullptr;

    const double* sw = a.spec + (size_t)b * 4 * N;
    const double h = a.grid_h[2 * b], w_ulp = a.grid_h[2 * b + 1];
    const size_t bs = (size_t)b * S + sl;
    constexpr double H = 16.0 * R;
    const double xi0 = cell_xi0<R>(lane, SUB);
    const LaneCell lcell = lane_cell(lane, SUB, P);
    const double inv_H = (double)SUB / H;

    // ---- load the particle and the swarm's shared state
    for (int d = tid; d < D; d += THREADS) {
        xs[d] = s.x[bs * D + d];
        vs[d] = s.v[bs * D + d];
        ps[d] = s.p[bs * D + d];
        gs[d] = s.g[(size_t)b * D + d];
        lbs[d] = s.lb[(size_t)b * D + d];
        ubs[d] = s.ub[(size_t)b * D + d];
        bxs[d] = s.best_x[(size_t)b * D + d];
    }
    if (tid == 0) {
        misc[0] = s.fx[bs];
        misc[1] = s.fp[bs];
        misc[2] = s.fg[b];
        misc[3] = s.best_f[b];
    }
    for (int i = tid; i < 64; i += THREADS) tab[i] = NMRFIT_EXP2_TAB6[i];
    int it = s.it[b];
    int stop = 0;

    auto stage = [&](int st, int slot) {                   // (u, v, weights) of supertile st -> slot
        const int base = st * THREADS * R;
        for (int e = tid; e < THREADS * R; e += THREADS) {
            const int i = base + e;
            const bool ok = i < N;
            const int t = e / R, j = e % R, o = slot * THREADS * R;
            const double wgt = ok ? sw[3 * N + i] : 0.0;      // zero weight: padding adds nothing
            suv[o + stage_slot_uv(t, j, THREADS)] = stage_point(ok ? sw[N + i] : 0.0, ok ? sw[2 * N + i] : 0.0, wgt);
            swt[o + stage_slot_wt(t, j, THREADS)] = wgt;
        }
    };
    if (resident)
        for (int st = st_lo; st < st_hi; ++st) stage(st, st - st_lo);
    __syncthreads();

    long long tmark = clock64();
    for (int k = 0; k < a.n_gen && !stop; ++k) {
        const int par = k & 1;
        // ---- move (pyswarm: v = omega v + phip rp (p - x) + phig rg (g - x); x += v; clamp)
        for (int d = tid; d < D; d += THREADS) {
            double rp, rg;
            if (a.rp) {
                const size_t idx = ((size_t)k * s.B * S + bs) * D + d;
                rp = a.rp[idx];
                rg = a.rg[idx];
            } else {
                const Philox2 u = philox_uniform2(s.seed, elem_counter(s, b, sl, d), (unsigned long long)(a.gen0 + k));
                rp = u.a;
                rg = u.b;
            }
            double x = xs[d], v = vs[d];
            move_element(s.omega, s.phip, s.phig, rp, rg, ps[d], gs[d], lbs[d], ubs[d], x, v);
            xs[d] = x;
            vs[d] = v;
        }
        __syncthreads();
        FUSED_MARK(0);

        // ---- objective (equations.py:152-212) of the moved particle
        prepare_particle<R>(xs, sw, h, w_ulp, N, P, NR, NRP, tid, THREADS, cs, nullptr, part, farc, anchor, mask, farpk, pairs,
                            r_lo, r_hi, SUB);
        __syncthreads();
        FUSED_MARK(1);
        for (int st = st_lo; st < st_hi; ++st) {
            if (!resident) {
                __syncthreads();
                stage(st, 0);
                __syncthreads();
            }
            const int rgn = st * NW + warp;
            if (rgn < NRP) {
                const int slot = resident ? st - st_lo : 0;
                const int rl = rgn - r_lo;                 // index into this CTA's region constants
                const int i_first = (st * THREADS + tid) * R;
                const double w_first = i_first < N ? __ldg(sw + i_first) : fma((double)i_first, h, __ldg(sw));
                const double2 ew = *reinterpret_cast<const double2*>(anchor + 2 * rl);
                const double ss = eval_region<R, TB>(cs, part, mask + (size_t)rl * mask_words_per_region(P, SUB),
                                                     farc + (size_t)rl * SUB * kFarPoly, ew, MW, P, lane, lcell, w_first, xi0,
                                                     inv_H, suv + slot * THREADS * R, swt + slot * THREADS * R, tid, THREADS,
                                                     tab, xs, sw + i_first, N - i_first, h, w_ulp);
                if (!CL) {
                    if (lane == 0) wpart[rgn] = ss;
                } else if (lane < G) {
                    // all-gather through distributed shared memory: lane q stores into CTA q of the cluster
                    cooperative_groups::this_cluster().map_shared_rank(wpart, lane)[rgn] = ss;
                }
            }
        }
        if (!CL) __syncthreads();
        else cooperative_groups::this_cluster().sync();    // every CTA of the particle now holds every region's sum
        FUSED_MARK(2);
        if (tid == 0) {
            // same order as the per-step path: the warps of a point tile, then the tiles, then sqrt(mean)
            double total = 0.0;
            for (int t = 0; t < a.n_vtiles; ++t) {
                double tt = 0.0;
                for (int wi = 0; wi < VW; ++wi) tt += wpart[t * VW + wi];
                total += tt;
            }
            const double fx = sqrt(total / (double)N);
            const bool better = fx < misc[1];
            misc[0] = fx;
            if (better) misc[1] = fx;
            misc[4] = better ? 1.0 : 0.0;
            if (crank == 0) a.rec_f[((size_t)par * s.B + b) * S + sl] = misc[1];
        }
        __syncthreads();
        FUSED_MARK(3);
        {
            const bool better = misc[4] != 0.0;
            double* rx = a.rec_x + (((size_t)par * s.B + b) * S + sl) * D;
            for (int d = tid; d < D; d += THREADS) {
                if (better) ps[d] = xs[d];
                if (crank == 0) rx[d] = ps[d];
            }
        }

        // ---- barrier among the S CTAs of this spectrum
        __syncthreads();
        FUSED_MARK(4);
        if (tid == 0) {
            // release: this CTA's record (ordered before by the bar.sync above) is visible to whoever acquires the count
            red_release_add(a.barrier + b, 1u);
            const unsigned target = (unsigned)(S * G) * (unsigned)(k + 1);
            // The cooperative launch guarantees that every CTA is resident, so this wait ends within microseconds -
            // unless a CTA of the spectrum died.  Bounded in wall time: on expiry raise the context's error flag and
            // leave (exited threads count as arrived at every later barrier, so nobody is left hanging).
            unsigned spins = 0;
            unsigned long long t0 = 0;
            misc[7] = 0.0;
            while (ld_acquire(a.barrier + b) < target) {
                if ((++spins & 1023u) == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > (unsigned long long)a.max_wait_ns) { misc[7] = 1.0; break; }
                }
            }
        }
        __syncthreads();
        if (misc[7] != 0.0) {
            if (tid == 0) atomicExch(a.error, 1);
            return;
        }
        FUSED_MARK(5);

        // ---- swarm best: argmin over the personal bests, first index wins (np.argmin)
        {
            const double* rf = a.rec_f + ((size_t)par * s.B + b) * S;
            double bf = CUDART_INF;
            int bi = 0x7fffffff;
            for (int i = tid; i < S; i += THREADS) take_min(bf, bi, __ldcg(rf + i), i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double of = __shfl_xor_sync(0xffffffffu, bf, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                take_min(bf, bi, of, oi);
            }
            if (lane == 0) { redf[warp] = bf; redi[warp] = bi; }
        }
        __syncthreads();
        if (warp == 0) {
            double bf = lane < NW ? redf[lane] : CUDART_INF;
            int bi = lane < NW ? redi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double of = __shfl_xor_sync(0xffffffffu, bf, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                take_min(bf, bi, of, oi);
            }
            const int win = bi == 0x7fffffff ? 0 : bi;
            double* sq = pm + De;                          // (g - p_min)^2, element by element
            if (bf < misc[2]) {                            // only a new swarm best needs its position
                const double* rx = a.rec_x + (((size_t)par * s.B + b) * S + win) * D;
                for (int d = lane; d < D; d += 32) {
                    const double pd = __ldcg(rx + d);
                    const double df = __dsub_rn(gs[d], pd);
                    pm[d] = pd;
                    sq[d] = __dmul_rn(df, df);
                }
            }
            __syncwarp();
            if (lane == 0) {
                // pyswarm: if fp[i_min] < fg: test |fg - fp[i_min]| <= minfunc, then ||g - p_min|| <= minstep, else adopt
                const double fmin = bf, fg = misc[2];
                int action = 0;
                if (fmin < fg) {
                    const double step = sqrt(numpy_pairwise_sum<3>(sq, D));   // np.sum's order, as the per-step commit
                    if (fabs(__dsub_rn(fg, fmin)) <= s.minfunc) action = 2 + kStopMinFunc;
                    else if (step <= s.minstep) action = 2 + kStopMinStep;
                    else action = 1;
                }
                misc[5] = (double)action;
                misc[6] = fmin;
            }
        }
        __syncthreads();
        FUSED_MARK(6);
        {
            const int action = (int)misc[5];
            it += 1;
            if (action >= 2) stop = action - 2;
            else if (it >= a.maxiter) stop = kStopMaxIter;
            if (action != 0) {
                for (int d = tid; d < D; d += THREADS) {
                    bxs[d] = pm[d];
                    if (action == 1) gs[d] = pm[d];
                }
                if (tid == 0) {
                    misc[3] = misc[6];
                    if (action == 1) misc[2] = misc[6];
                }
            }
        }
        __syncthreads();
        FUSED_MARK(7);
    }

    // ---- write the state back for the host (and for further generations by either path)
    if (crank != 0) return;                                // the cluster's CTAs hold identical state: one writes it back
    for (int d = tid; d < D; d += THREADS) {
        s.x[bs * D + d] = xs[d];
        s.v[bs * D + d] = vs[d];
        s.p[bs * D + d] = ps[d];
        if (sl == 0) {
            s.g[(size_t)b * D + d] = gs[d];
            s.best_x[(size_t)b * D + d] = bxs[d];
        }
    }
    if (tid == 0) {
        s.fx[bs] = misc[0];
        s.fp[bs] = misc[1];
        if (sl == 0) {
            s.fg[b] = misc[2];
            s.best_f[b] = misc[3];
            s.it[b] = it;
            s.stop[b] = stop;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------

// dense = 3 CTAs of 256 threads per SM (80 registers, a few spills) instead of 2: only when that is what makes the
// clusters co-resident
static const void* fused_kernel(int threads, int r, int cluster, bool dense) {
    if (cluster > 1) {
        if (threads != 256) return nullptr;
        if (dense) return r == 8 ? (const void*)swarm_fused_kernel<256, 8, 6, true, 3> : r == 4 ? (const void*)swarm_fused_kernel<256, 4, 6, true, 3> : nullptr;
        return r == 8 ? (const void*)swarm_fused_kernel<256, 8, 6, true, 2> : r == 4 ? (const void*)swarm_fused_kernel<256, 4, 6, true, 2> : nullptr;
    }
    if (threads == 512 && r == 8) return (const void*)swarm_fused_kernel<512, 8, 6, false, 1>;
    if (threads == 512 && r == 4) return (const void*)swarm_fused_kernel<512, 4, 6, false, 1>;
    if (threads == 256 && r == 8) return (const void*)swarm_fused_kernel<256, 8, 6, false, 2>;
    if (threads == 256 && r == 4) return (const void*)swarm_fused_kernel<256, 4, 6, false, 2>;
    return nullptr;
}

static void fill_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int grid, int threads, size_t smem, int cluster,
                        cudaStream_t st) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)grid);
    cfg->blockDim = dim3((unsigned)threads);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
    at[0].id = cudaLaunchAttributeCooperative;             // every CTA co-resident, or the launch fails
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)cluster;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg->attrs = at;
    cfg->numAttrs = cluster > 1 ? 2 : 1;
}

// Can `particles * cluster` CTAs of `threads` threads be co-resident?  Tries the whole share of the spectrum resident
// in shared memory first, then one supertile at a time; with the pair-parallel constants pass, then without its scratch.
static cudaError_t plan_one(const FusedArgs& a, int D, int particles, int threads, int r, int cluster, bool dense,
                            int device, FusedPlan* plan) {
    const void* kern = fused_kernel(threads, r, cluster, dense);
    if (!kern) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    const int NW = threads / 32, NRP = a.n_vtiles * a.vw, n_super = (NRP + NW - 1) / NW;
    const int per = (n_super + cluster - 1) / cluster;
    const int grid = particles * cluster;
    // keeping the spectrum resident matters more than the pair-parallel constants pass (whose scratch grows with the
    // number of far-field cells): resident + pairs, resident, then one supertile at a time with / without pairs
    for (int attempt = 0; attempt < 4; ++attempt) {
        const int slots = attempt < 2 ? per : 1;
        const bool pairs = (attempt & 1) == 0;
        if (attempt >= 2 && per == 1) continue;
        const size_t bytes = (size_t)FusedSmem(a.P, D, threads, r, slots, NRP, per * NW, pairs, a.sub).total * sizeof(double);
        if (bytes > 200 * 1024) continue;
        long long capacity = 0;
        if (cluster > 1) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute at[2];
            fill_config(&cfg, at, grid, threads, bytes, cluster, nullptr);
            int clusters = 0;
            e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); clusters = 0; }
            capacity = (long long)clusters * cluster;
        } else {
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, bytes);
            if (e != cudaSuccess) return e;
            capacity = (long long)per_sm * sms;
        }
        if (capacity >= grid) {
            plan->ok = true; plan->threads = threads; plan->r = r; plan->slots = slots; plan->cluster = cluster;
            plan->dense = dense; plan->pairs = pairs; plan->smem = bytes;
            return cudaSuccess;
        }
    }
    return cudaSuccess;
}

cudaError_t swarm_fused_plan(const FusedArgs& a, int D, int B, int S, const ObjTune& t, int device, bool force,
                             FusedPlan* plan) {
    plan->ok = false;
    if (t.tb != 6 || (t.r != 4 && t.r != 8) || (t.threads != 128 && t.threads != 256)) return cudaSuccess;
    const long long particles = (long long)B * S;
    if (particles > 4096) return cudaSuccess;
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    const int NRP = a.n_vtiles * a.vw;
    // Short axes (<= 16 regions): one CTA per particle - 16 warps when every particle has an SM to itself.
    if (NRP <= 16) {
        // (a cluster of two 8-warp CTAs instead of one 16-warp CTA was measured here: no gain - the phase is bound by
        // each warp's own dependent chain, not by the SM's FP64 rate)
        if (particles <= sms && NRP > 8) {
            e = plan_one(a, D, (int)particles, 512, t.r, 1, false, device, plan);
            if (e != cudaSuccess || plan->ok) return e;
        }
        return plan_one(a, D, (int)particles, 256, t.r, 1, false, device, plan);
    }
    // Longer axes: a thread-block cluster per particle, as many CTAs as stay co-resident (<= 8, the portable limit),
    // each with its run of supertiles; past ~32 regions per CTA the three-launch per-step path, which spreads a
    // particle's tiles over many CTAs, is faster (profiles/r01k_fused_probe.json).
    const int n_super = (NRP + 7) / 8;
    for (int cluster = 8; cluster >= 1; cluster /= 2) {
        if (cluster > n_super) continue;
        const int per = (n_super + cluster - 1) / cluster;
        if (!force && per * 8 > 32) break;                 // smaller clusters only get longer runs
        for (bool dense : {false, true}) {
            if (dense && cluster == 1) continue;
            e = plan_one(a, D, (int)particles, 256, t.r, cluster, dense, device, plan);
            if (e != cudaSuccess || plan->ok) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_swarm_fused(FusedArgs a, const FusedPlan& plan, int B, int S, cudaStream_t st) {
    a.slots = plan.slots;
    a.cluster = plan.cluster;
    a.pairs = plan.pairs ? 1 : 0;
    const void* kern = fused_kernel(plan.threads, plan.r, plan.cluster, plan.dense);
    if (!kern) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(a.barrier, 0, sizeof(unsigned) * B, st);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[2];
    fill_config(&cfg, at, B * S * plan.cluster, plan.threads, plan.smem, plan.cluster, st);
    void* args[] = {&a};
    e = cudaLaunchKernelExC(&cfg, kern, args);
    count_launches(1);
    return e;
}

}  // namespace nmrfit
