#!/usr/bin/env python
"""Randomised differential test on the GPU: random shapes (points, peaks, particles, spectra), objective against
the CPU oracle, fused swarm kernel against the per-step kernels (bitwise), device weights against the oracle.

    python tests/fuzz_parity.py [--cases 40] [--seed 0]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, swarm, synth, utils      # noqa: E402
from oracle import nmrfit_oracle as orc                  # checker only  # noqa: E402

PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)


def swarm_run(specs, los, ups, S, iters, mode, rs_seed, host_rng):
    B, N = len(specs), len(specs[0][0])
    D = len(los[0])
    rs = np.random.RandomState(rs_seed)
    with _cabi.Context(B, N, (D - 4) // 3) as ctx:
        ctx.set_fused(mode)
        ctx.set_spectra(*[np.stack([sp[k] for sp in specs]) for k in range(4)])
        opts = swarm._make_opts(S, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-6, False, 99)
        r = (lambda *sh: rs.rand(*sh)) if host_rng else (lambda *sh: None)
        ctx.pso_begin(np.array(los), np.array(ups), opts, r(B, S, D), r(B, S, D))
        ctx.pso_commit()
        rp, rg = r(iters, B, S, D), r(iters, B, S, D)
        done = 0
        for n in (3, iters):
            n = min(n, iters - done)
            if n <= 0:
                break
            ctx.pso_run(n, None if rp is None else rp[done:done + n], None if rg is None else rg[done:done + n])
            done += n
        return ctx.pso_best(), ctx.pso_state(), ctx.fused_launches()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=40)
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    worst = {'objective_rel': 0.0, 'objective_im_rel': 0.0}
    fails = []
    for case in range(args.cases):
        P = int(rng.choice([6, 6, 12, 18, 24, 36]))
        N = int(rng.choice([rng.integers(40, 400), rng.integers(400, 3000), rng.integers(3000, 9000), 4096, 2048]))
        S = int(rng.integers(1, 60))
        B = int(rng.choice([1, 1, 2, 3]))
        desc = dict(case=case, P=P, N=N, S=S, B=B)
        specs, los, ups = [], [], []
        for b in range(B):
            data, true = synth.multiplet(max(N, 64), P, seed=int(rng.integers(1, 10 ** 6)))
            w, u, v = data.w[:N], data.u[:N], data.v[:N]
            if rng.random() < 0.25:
                w, u, v = w[::-1].copy(), u[::-1].copy(), v[::-1].copy()      # descending axis
            wts = utils.compute_weights(w, data.peaks)
            lo, up = data.generate_solution_bounds()
            specs.append((w, u, v, wts)); los.append(lo); ups.append(up)
        # objective vs oracle (first spectrum), real only and fit_im
        xs = synth.particles(los[0], ups[0], 5, seed=case)
        w, u, v, wts = specs[0]
        with _cabi.Context(1, N, P) as ctx:
            ctx.set_spectrum(0, w, u, v, wts)
            for mode, key in ((_cabi.REAL_ONLY, 'objective_rel'), (_cabi.IM_REFERENCE, 'objective_im_rel')):
                got = ctx.objective_host(xs, mode)
                want = np.array([orc.objective(x, w, u, v, wts, mode == _cabi.IM_REFERENCE) for x in xs])
                rel = float(np.max(np.abs(got / want - 1)))
                worst[key] = max(worst[key], rel)
                if not rel < 1e-10:
                    fails.append(dict(desc, what=key, rel=rel))
        # device weights vs host
        peaks_list = []
        for b in range(B):
            data, _ = synth.multiplet(max(N, 64), P, seed=1 + b)
            peaks_list.append(data.peaks)
        with _cabi.Context(B, N, P) as ctx:
            W = np.stack([sp[0] for sp in specs])
            ctx.set_spectra(W, W, W)
            got = ctx.compute_weights(*utils.peak_windows(peaks_list))
            for b in range(B):
                if not np.array_equal(got[b], orc.compute_weights(specs[b][0], peaks_list[b])):
                    fails.append(dict(desc, what='weights', b=b))
        # fused vs per-step, bitwise
        iters = int(rng.integers(4, 14))
        host_rng = bool(rng.random() < 0.5)
        try:
            fa = swarm_run(specs, los, ups, S, iters, _cabi.FUSED_REQUIRE, case, host_rng)
            st = swarm_run(specs, los, ups, S, iters, _cabi.FUSED_OFF, case, host_rng)
            same = all(np.array_equal(a, b) for a, b in zip(fa[0], st[0])) and \
                all(np.array_equal(fa[1][k], st[1][k]) for k in ('x', 'v', 'p', 'fx', 'fp'))
            if not same or fa[2] < 1:
                fails.append(dict(desc, what='fused_vs_per_step', fused_launches=fa[2]))
        except _cabi.NmrfitError as e:
            fails.append(dict(desc, what='fused_error', msg=str(e)[:200]))
    print(json.dumps({'cases': args.cases, 'seed': args.seed, 'worst': worst, 'failures': fails}, indent=1))
    return 1 if fails else 0


if __name__ == '__main__':
    sys.exit(main())
