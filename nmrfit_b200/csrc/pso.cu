// K2/K3: the particle-swarm bookkeeping that pyswarm.pso does in numpy on the host
// (called from the reference at nmrfit/utils.py:176-182; algorithm restated in
// oracle/pso_oracle.py), kept on the device so that a generation is
// move -> objective -> personal best -> swarm best without a host round trip.
//
// The position/velocity update reproduces numpy's evaluation order
//   v = ((omega*v) + ((phip*rp)*(p-x))) + ((phig*rg)*(g-x));  x = x + v
// with explicitly rounded multiplies/adds (no FMA contraction), so that with
// host-supplied random numbers the trajectory is bit-identical to the CPU one as
// long as the objective values compare the same way.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "nmrfit_internal.h"
#include "nmrfit_math.cuh"
#include "swarm_common.cuh"

namespace nmrfit {

// generation 0, part 1: x = lb + r*(ub - lb); fp = inf   (pyswarm: x = rand(S,D); x = lb + x*(ub-lb))
__global__ void swarm_init_kernel(SwarmState s, const double* __restrict__ r_pos) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)s.B * s.S * s.D;
    if (idx >= total) return;
    int d = idx % s.D;
    size_t bs = idx / s.D;
    int sl = bs % s.S, b = bs / s.S;
    double r = r_pos ? r_pos[idx] : philox_uniform2(s.seed, elem_counter(s, b, sl, d), 0ull).a;
    double lb = s.lb[b * s.D + d], ub = s.ub[b * s.D + d];
    s.x[idx] = __dadd_rn(lb, __dmul_rn(r, __dsub_rn(ub, lb)));
    s.p[idx] = 0.0;
    if (d == 0) s.fp[bs] = CUDART_INF;
}

// generation 0, part 2 (after the first evaluation): v = vlow + r*(vhigh - vlow),
// vhigh = |ub - lb|, vlow = -vhigh
__global__ void swarm_init_velocity_kernel(SwarmState s, const double* __restrict__ r_vel) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)s.B * s.S * s.D;
    if (idx >= total) return;
    int d = idx % s.D;
    size_t bs = idx / s.D;
    int sl = bs % s.S, b = bs / s.S;
    double r = r_vel ? r_vel[idx] : philox_uniform2(s.seed, elem_counter(s, b, sl, d), 0ull).b;
    double vhigh = fabs(__dsub_rn(s.ub[b * s.D + d], s.lb[b * s.D + d]));
    double vlow = -vhigh;
    s.v[idx] = __dadd_rn(vlow, __dmul_rn(r, __dsub_rn(vhigh, vlow)));
}

__global__ void swarm_move_kernel(SwarmState s, const double* __restrict__ rp_in, const double* __restrict__ rg_in,
                                  int generation) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)s.B * s.S * s.D;
    if (idx >= total) return;
    int d = idx % s.D;
    size_t bs = idx / s.D;
    int sl = bs % s.S, b = bs / s.S;
    if (s.stop[b]) return;
    double rp, rg;
    if (rp_in) {
        rp = rp_in[idx];
        rg = rg_in[idx];
    } else {
        Philox2 u = philox_uniform2(s.seed, elem_counter(s, b, sl, d), (unsigned long long)generation);
        rp = u.a;
        rg = u.b;
    }
    double x = s.x[idx], v = s.v[idx], p = s.p[idx], g = s.g[b * s.D + d];
    move_element(s.omega, s.phip, s.phig, rp, rg, p, g, s.lb[b * s.D + d], s.ub[b * s.D + d], x, v);
    s.v[idx] = v;
    s.x[idx] = x;
}

// swarm-best update and the minfunc/minstep stop tests for spectrum b, by one CTA of `nthreads` threads.
// `recs` holds one record per rank ([n_ranks][B][D+2]); every rank runs this redundantly on identical input,
// so g/fg stay bit-identical everywhere.
// Record of rank r for this spectrum: recs + r * rank_stride + b_offset (global [n_ranks][B][D+2] layout:
// rank_stride = B * (D+2), b_offset = b * (D+2); a per-spectrum shared-memory copy: D+2 and 0).
__device__ __forceinline__ void commit_spectrum(const SwarmState& s, const double* recs, int n_ranks,
                                                int initial, int maxiter, int b, int tid, int nthreads,
                                                size_t rank_stride, size_t b_offset) {
    const int W = s.D + 2;
    (void)W;
    __shared__ int s_win;
    __shared__ double s_step;
    __shared__ int s_action;   // 0 nothing, 1 adopt as g, 2 stop (return p_min)
    if (tid == 0) {
        int win = 0;
        double wf = recs[b_offset], wi = recs[b_offset + 1];
        for (int r = 1; r < n_ranks; ++r) {
            const double* q = recs + (size_t)r * rank_stride + b_offset;
            if (q[0] < wf || (q[0] == wf && q[1] < wi)) { wf = q[0]; wi = q[1]; win = r; }
        }
        s_win = win;
    }
    __syncthreads();
    const double* q = recs + (size_t)s_win * rank_stride + b_offset;
    const double fmin = q[0];
    double* g = s.g + (size_t)b * s.D;
    if (initial) {
        // pyswarm: fg = fp[i_min]; g = p[i_min] (else g = x[0] when nothing is finite)
        for (int d = tid; d < s.D; d += nthreads) { g[d] = q[2 + d]; s.best_x[(size_t)b * s.D + d] = q[2 + d]; }
        if (tid == 0) { s.fg[b] = fmin; s.best_f[b] = fmin; s.it[b] = 0; }
        return;
    }
    const double fg = s.fg[b];
    // stepsize = np.sqrt(np.sum((g - p_min)**2)): the squares element by element, then numpy's pairwise order
    __shared__ double s_sq[kMaxParams];
    if (fmin < fg)
        for (int d = tid; d < s.D; d += nthreads) {
            const double df = __dsub_rn(g[d], q[2 + d]);
            s_sq[d] = __dmul_rn(df, df);
        }
    __syncthreads();
    if (tid == 0) {
        int action = 0;
        if (fmin < fg) {
            s_step = sqrt(numpy_pairwise_sum<3>(s_sq, s.D));
            if (fabs(__dsub_rn(fg, fmin)) <= s.minfunc) { action = 2; s.stop[b] = kStopMinFunc; }
            else if (s_step <= s.minstep) { action = 2; s.stop[b] = kStopMinStep; }
            else action = 1;
        }
        s_action = action;
        s.it[b] += 1;
        if (action == 0 || action == 1) {
            if (s.it[b] >= maxiter) s.stop[b] = kStopMaxIter;
        }
    }
    __syncthreads();
    if (s_action == 0) return;
    for (int d = tid; d < s.D; d += nthreads) {
        s.best_x[(size_t)b * s.D + d] = q[2 + d];
        if (s_action == 1) g[d] = q[2 + d];
    }
    if (tid == 0) {
        s.best_f[b] = fmin;
        if (s_action == 1) s.fg[b] = fmin;
    }
}

__global__ void __launch_bounds__(128) swarm_commit_kernel(SwarmState s, const double* __restrict__ recs, int n_ranks,
                                                           int initial, int maxiter) {
    const int b = blockIdx.x;
    if (s.stop[b]) return;
    commit_spectrum(s, recs, n_ranks, initial, maxiter, b, threadIdx.x, 128, (size_t)s.B * (s.D + 2), (size_t)b * (s.D + 2));
}

// ---- record exchange over peer memory (particle sharding without a collective call) -------------------------
// Every rank's context owns a window [2 parities][n_ranks][B][D+2] of records plus [n_ranks][B] 64-bit tokens,
// mapped into every peer (CUDA IPC between processes; plain pointers between contexts of one process).  One CTA per
// spectrum stores this rank's local best record into slot `rank` of EVERY peer's window (NVLink stores), fences at
// system scope, stores the generation token into every peer's token array, waits until the tokens of all ranks have
// arrived in its own window, and applies the same commit as every other rank on the same n_ranks records - the
// all-gather + commit of the NCCL path without a collective call.  It runs inside the finish kernel's last CTA of the
// spectrum (a sharded generation is then the same three launches as an unsharded one) or, for generation 0, as a kernel
// of its own.  Windows are double-buffered by generation parity: a rank cannot be two generations ahead, because its next
// commit needs the slowest rank's next token.  The wait is bounded in WALL TIME (%globaltimer): on expiry the CTA
// raises *error, marks the spectrum stopped (kStopPeerLost) so that this rank does not step it any further - every
// other rank then runs into the same timeout one generation later - and returns without committing.
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Called by all `nthreads` threads of one CTA for spectrum b; srec: shared [n_ranks][D+2].  Returns false on timeout.
__device__ __forceinline__ bool exchange_records(const SwarmState& s, const PeerArgs& pa, int b, int tid, int nthreads,
                                                 double* srec) {
    const int W = s.D + 2, R = pa.n_ranks;
    const int par = (int)(pa.token & 1);
    const double* mine = s.rec + (size_t)b * W;
    for (int q = 0; q < R; ++q) {
        double* dst = pa.recs[q] + (((size_t)par * R + pa.rank) * s.B + b) * W;
        for (int d = tid; d < W; d += nthreads) dst[d] = mine[d];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < R) st_release_sys(pa.tokens[tid] + (size_t)pa.rank * s.B + b, pa.token);
    __shared__ int s_timeout;
    if (tid == 0) s_timeout = 0;
    __syncthreads();
    if (tid < R) {
        const long long* mytok = pa.tokens[pa.rank] + (size_t)tid * s.B + b;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(mytok) < pa.token) {
            if ((++spins & 255u) == 0 && global_timer_ns() - t0 > (unsigned long long)pa.max_wait_ns) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (tid == 0) {
            atomicExch(pa.error, 1);
            s.stop[b] = kStopPeerLost;
        }
        return false;
    }
    const double* win = pa.recs[pa.rank] + ((size_t)par * R * s.B + b) * W;
    for (int i = tid; i < R * W; i += nthreads) {
        const int q = i / W, d = i - q * W;
        srec[i] = __ldcv(win + (size_t)q * s.B * W + d);   // written by peers: never from a stale L1 line
    }
    __syncthreads();
    return true;
}

// ---- finish: everything after the objective's tile sums, in ONE launch --------------------------------------
// Per particle (one warp each): fixed-order sum of its tile partials -> fx = sqrt(mean) (equations.py:202,
// 205-209), personal-best update (value and position), and a candidate (fp, index) for the swarm's argmin.
// The CTAs of a spectrum combine their candidates through a ticket: the last one to finish reduces them
// (argmin with the lowest-index tie-break is order-independent, so this is deterministic), writes the local
// best record and - when `commit` - applies pyswarm's swarm-best update and stop tests right there.
constexpr int kFinWarps = 8;

__global__ void __launch_bounds__(kFinWarps * 32)
swarm_finish_kernel(SwarmState s, const double* __restrict__ partials, int n_tiles, int nsum, int N,
                    double* rec, double* __restrict__ scratch, unsigned* __restrict__ tickets, int commit,
                    int maxiter, int nw, int L, PeerArgs pa) {
    extern __shared__ __align__(16) double srec[];         // commit == 2: the ranks' records of this spectrum
    const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
    if (s.stop[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // L lanes per particle (a power of two, at least the number of tiles up to a whole warp): a warp takes 32 / L
    // particles - with a whole warp per particle a two-tile axis left 30 lanes idle and 65,536 particles were 8,192
    // CTAs, each a chain of dependent memory round trips, plus as many candidates for the last CTA to scan
    const int ppw = 32 / L, sl = lane & (L - 1), lane0 = lane & ~(L - 1);
    const int i = (blk * kFinWarps + warp) * ppw + lane / L;     // this lane group's particle
    __shared__ double sf[kFinWarps];
    __shared__ int si[kFinWarps];
    __shared__ int s_last;
    double cf = CUDART_INF;
    int ci = 0x7fffffff;
    {
        const bool valid = i < s.S;
        const size_t bs = (size_t)b * s.S + (valid ? i : 0);
        // partial sums per tile (nw == 1) or per region, [n_tiles][nw] (the streamed kernel; nw = 4 or 8 divides 32):
        // the regions of a tile are added first, then the tiles
        const double* p = partials + bs * n_tiles * nw * nsum;
        double sv = 0.0, sim = 0.0;
        for (int t0 = 0; t0 < n_tiles; t0 += L) {          // a lane per tile: its regions in order (in parallel over the
            const int t = t0 + sl;                         // lanes), then the tiles in order - every lane sums all of
            double a0 = 0.0, a1 = 0.0;                     // them, the same sequential order as objective_finalize_kernel
            if (valid && t < n_tiles) {
                for (int w = 0; w < nw; ++w) {
                    a0 += p[(t * nw + w) * nsum];
                    if (nsum == 2) a1 += p[(t * nw + w) * nsum + 1];
                }
            }
            const int cnt = min(L, n_tiles - t0);
            for (int k = 0; k < cnt; ++k) {
                sv += __shfl_sync(0xffffffffu, a0, lane0 + k);
                if (nsum == 2) sim += __shfl_sync(0xffffffffu, a1, lane0 + k);
            }
        }
        if (valid) {
            double fx = sqrt(sv / (double)N);
            if (nsum == 2) fx = (fx + sqrt(sim / (double)N)) / 2.0;
            double fp = s.fp[bs];
            if (fx < fp) {
                fp = fx;
                for (int d = sl; d < s.D; d += L) s.p[bs * s.D + d] = s.x[bs * s.D + d];
            }
            if (sl == 0) { s.fx[bs] = fx; s.fp[bs] = fp; }
            cf = fp;
            ci = i;
        }
    }
    // the warp's best particle (lowest index wins a tie)
    for (int o = L; o < 32; o <<= 1) {
        const double of = __shfl_xor_sync(0xffffffffu, cf, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
        if (of < cf || (of == cf && oi < ci)) { cf = of; ci = oi; }
    }
    if (lane == 0) { sf[warp] = cf; si[warp] = ci; }
    __syncthreads();
    if (tid == 0) {
        double bf = CUDART_INF;
        int bi = 0x7fffffff;
        for (int k = 0; k < kFinWarps; ++k)
            if (sf[k] < bf || (sf[k] == bf && si[k] < bi)) { bf = sf[k]; bi = si[k]; }
        double* sc = scratch + ((size_t)b * nblk + blk) * 2;
        sc[0] = bf;
        sc[1] = (double)bi;
        __threadfence();                                   // candidate and the p rows above are visible before the ticket
        const unsigned t = atomicAdd(tickets + b, 1u);
        s_last = t == (unsigned)nblk - 1u;
        if (s_last) tickets[b] = 0u;                       // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last CTA of this spectrum: argmin over the CTA candidates, first index wins (np.argmin)
    double bf = CUDART_INF;
    int bi = 0x7fffffff;
    for (int k = tid; k < nblk; k += kFinWarps * 32) {
        const double* sc = scratch + ((size_t)b * nblk + k) * 2;
        const double of = __ldcg(sc);
        const int oi = (int)__ldcg(sc + 1);
        if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xffffffffu, bf, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
    }
    __syncthreads();                                       // sf/si are reused
    if (lane == 0) { sf[warp] = bf; si[warp] = bi; }
    __syncthreads();
    bf = sf[0]; bi = si[0];
    for (int k = 1; k < kFinWarps; ++k)
        if (sf[k] < bf || (sf[k] == bf && si[k] < bi)) { bf = sf[k]; bi = si[k]; }
    int best = bi;
    const double* src = s.p;
    if (best == 0x7fffffff) { best = 0; src = s.x; }       // nothing finite yet: pyswarm falls back to x[0]
    double* r = rec + (size_t)b * (s.D + 2);
    if (tid == 0) { r[0] = bf; r[1] = (double)(s.index0 + best); }
    for (int d = tid; d < s.D; d += kFinWarps * 32) r[2 + d] = __ldcg(src + ((size_t)b * s.S + best) * s.D + d);
    if (!commit) return;
    __syncthreads();                                       // the record is complete (same CTA wrote it)
    if (commit == 2) {
        // particle sharding: exchange the ranks' records over peer memory, then the same commit on all of them
        __threadfence();
        if (!exchange_records(s, pa, b, tid, kFinWarps * 32, srec)) return;
        commit_spectrum(s, srec, pa.n_ranks, 0, maxiter, b, tid, kFinWarps * 32, (size_t)(s.D + 2), 0);
        return;
    }
    commit_spectrum(s, rec, 1, 0, maxiter, b, tid, kFinWarps * 32, (size_t)s.B * (s.D + 2), (size_t)b * (s.D + 2));
}

__global__ void __launch_bounds__(128)
swarm_exchange_commit_kernel(SwarmState s, PeerArgs pa, int initial, int maxiter) {
    extern __shared__ __align__(16) double srec[];         // [n_ranks][D+2] of this spectrum
    const int b = blockIdx.x, tid = threadIdx.x;
    if (s.stop[b]) return;                                 // identical on every rank
    if (!exchange_records(s, pa, b, tid, 128, srec)) return;
    commit_spectrum(s, srec, pa.n_ranks, initial, maxiter, b, tid, 128, (size_t)(s.D + 2), 0);
}


cudaError_t launch_swarm_exchange_commit(const SwarmState& s, const PeerArgs& pa, int initial, int maxiter,
                                         cudaStream_t st) {
    const size_t smem = (size_t)pa.n_ranks * (s.D + 2) * sizeof(double);
    swarm_exchange_commit_kernel<<<s.B, 128, smem, st>>>(s, pa, initial, maxiter);
    count_launches(1);
    return cudaGetLastError();
}

static inline unsigned blocks_for(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

cudaError_t launch_swarm_init(const SwarmState& s, const double* r_pos, const double*, cudaStream_t st) {
    size_t total = (size_t)s.B * s.S * s.D;
    swarm_init_kernel<<<blocks_for(total, 256), 256, 0, st>>>(s, r_pos);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_swarm_init_velocity(const SwarmState& s, const double* r_vel, cudaStream_t st) {
    size_t total = (size_t)s.B * s.S * s.D;
    swarm_init_velocity_kernel<<<blocks_for(total, 256), 256, 0, st>>>(s, r_vel);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_swarm_move(const SwarmState& s, const double* rp, const double* rg, int generation,
                              cudaStream_t st) {
    size_t total = (size_t)s.B * s.S * s.D;
    swarm_move_kernel<<<blocks_for(total, 256), 256, 0, st>>>(s, rp, rg, generation);
    count_launches(1);
    return cudaGetLastError();
}

size_t swarm_finish_scratch_doubles(int B, int S) { return (size_t)B * ((S + kFinWarps - 1) / kFinWarps) * 2; }

cudaError_t launch_swarm_finish(const SwarmState& s, const double* partials, int n_tiles, int nsum, int N, double* rec,
                                double* scratch, unsigned* tickets, int commit, int maxiter, cudaStream_t st, int nw,
                                const PeerArgs* pa) {
    int L = 1;                                             // lanes per particle: a power of two >= n_tiles, at most a warp
    while (L < 32 && L < n_tiles) L <<= 1;
    const int per_cta = kFinWarps * (32 / L);
    dim3 grid((s.S + per_cta - 1) / per_cta, s.B);
    const size_t smem = commit == 2 ? (size_t)pa->n_ranks * (s.D + 2) * sizeof(double) : 0;
    swarm_finish_kernel<<<grid, kFinWarps * 32, smem, st>>>(s, partials, n_tiles, nsum, N, rec, scratch, tickets, commit,
                                                            maxiter, nw, L, pa ? *pa : PeerArgs{});
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_swarm_commit(const SwarmState& s, const double* recs, int n_ranks, int initial, int maxiter,
                                cudaStream_t st) {
    swarm_commit_kernel<<<s.B, 128, 0, st>>>(s, recs, n_ranks, initial, maxiter);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace nmrfit
