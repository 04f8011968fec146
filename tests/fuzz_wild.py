#!/usr/bin/env python
"""Adversarial parameter vectors: widths from far below the grid spacing to wider than the window, centres inside,
at the edges and far outside the window, any Lorentzian fraction, large phases.  Uniform-axis and any-axis kernels
against the CPU oracle (north-star bar 1e-9 relative)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, synth, utils      # noqa: E402
from oracle import nmrfit_oracle as orc          # checker only  # noqa: E402


def main():
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    worst = {}
    bad = []
    for case in range(24):
        P = int(rng.choice([6, 12, 24]))
        N = int(rng.choice([257, 1000, 2048, 4096, 6000, 16384]))
        data, true = synth.multiplet(max(N, 64), P, seed=case + 1)
        w, u, v = data.w[:N], data.u[:N], data.v[:N]
        wts = utils.compute_weights(w, data.peaks)
        S = 24
        xs = np.tile(true, (S, 1))
        span = w[-1] - w[0]
        h = abs(w[1] - w[0])
        for s in range(S):
            kind = s % 6
            k = rng.integers(0, P)
            if kind == 0:      # widths log-uniform over 7 decades, all peaks
                xs[s, 4::3] = 10.0 ** rng.uniform(-6, 1, P)
            elif kind == 1:    # one peak right at the recurrence / exact-path switch (R*|hG| ~ 4, R = 8)
                xs[s, 4 + 3 * k] = 8 * h * 2 * np.sqrt(np.log(2)) / 4.0 * rng.uniform(0.98, 1.02)
            elif kind == 2:    # centres at the edges and outside the window
                xs[s, 5::3] = w[0] + span * rng.uniform(-1.5, 2.5, P)
            elif kind == 3:    # pure Lorentzian / pure Gaussian, offsets, big areas
                xs[s, 2] = rng.choice([0.0, 1.0]); xs[s, 3] = rng.uniform(-1, 1); xs[s, 6::3] *= 10.0 ** rng.uniform(-3, 3, P)
            elif kind == 4:    # large phases
                xs[s, 0] = rng.uniform(-50, 50); xs[s, 1] = rng.uniform(-200, 200)
            else:              # everything at once
                xs[s, 4::3] = 10.0 ** rng.uniform(-5, 0, P); xs[s, 5::3] = w[0] + span * rng.uniform(-0.2, 1.2, P)
                xs[s, 0] = rng.uniform(-4, 4); xs[s, 1] = rng.uniform(-4, 4); xs[s, 2] = rng.uniform(0, 1)
        want = orc.objective_swarm(xs, w, u, v, wts)
        want_im = np.array([orc.objective(x, w, u, v, wts, True) for x in xs])
        with _cabi.Context(1, N, P) as ctx:
            ctx.set_spectrum(0, w, u, v, wts)
            for algo, name in ((_cabi.ALGO_UNIFORM, 'uniform'), (_cabi.ALGO_GENERAL, 'general')):
                ctx.set_algorithm(algo)
                for mode, ref, tag in ((_cabi.REAL_ONLY, want, 'real'), (_cabi.IM_REFERENCE, want_im, 'fit_im')):
                    got = ctx.objective_host(xs, mode)
                    rel = np.abs(got / ref - 1)
                    key = name + '_' + tag
                    worst[key] = max(worst.get(key, 0.0), float(np.nanmax(rel)))
                    for s in np.where(~(rel < 1e-9))[0]:
                        bad.append(dict(case=case, P=P, N=N, s=int(s), kind=int(s % 6), key=key, got=float(got[s]), want=float(ref[s])))
    print(json.dumps({'worst': worst, 'n_bad': len(bad), 'bad': bad[:20]}, indent=1))
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
