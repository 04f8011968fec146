// Element-level swarm arithmetic shared by the per-step swarm kernels (pso.cu) and the fused swarm
// kernel (swarm_fused.cu): both move a particle with exactly the same rounded operations.
#pragma once
#include <cuda_runtime.h>
#include "nmrfit_internal.h"
#include "nmrfit_math.cuh"

namespace nmrfit {

enum { kStopRunning = 0, kStopMinFunc = 1, kStopMinStep = 2, kStopMaxIter = 3, kStopPeerLost = 4 };
constexpr int kMaxParams = 4 + 3 * 256;                   // nmrfit_ctx_create admits up to 256 peaks

__device__ __forceinline__ unsigned long long elem_counter(const SwarmState& s, int b, int sl, int d) {
    // global (sharding-independent) element number
    return ((unsigned long long)(s.spec0 + b) << 40) ^ ((unsigned long long)(s.index0 + sl) * (unsigned long long)s.D + d);
}

// pyswarm's update in numpy's evaluation order, every operation rounded on its own (no FMA contraction):
//   v = ((omega*v) + ((phip*rp)*(p-x))) + ((phig*rg)*(g-x));  x = x + v;  x clamped to [lb, ub]
__device__ __forceinline__ void move_element(double omega, double phip, double phig, double rp, double rg, double p,
                                             double g, double lb, double ub, double& x, double& v) {
    double t1 = __dmul_rn(omega, v);
    double t2 = __dmul_rn(__dmul_rn(phip, rp), __dsub_rn(p, x));
    double t3 = __dmul_rn(__dmul_rn(phig, rg), __dsub_rn(g, x));
    v = __dadd_rn(__dadd_rn(t1, t2), t3);
    x = __dadd_rn(x, v);
    x = x < lb ? lb : (x > ub ? ub : x);
}

}  // namespace nmrfit
