#!/usr/bin/env python
"""Objective kernel time by fit_im mode (real only / reference semantics / sum over peaks) on BASELINE shapes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, synth, utils   # noqa: E402


def main():
    import torch
    out = {}
    for name, (P, N, S, seed) in {'c1x4096': (6, 4096, 4096, 1000), 'c2': (12, 32768, 4096, 2000)}.items():
        data, _ = synth.multiplet(N, P, seed=seed)
        lo, up = (np.array(a) for a in data.generate_solution_bounds())
        wts = utils.compute_weights(data.w, data.peaks)
        xs = torch.from_numpy(synth.particles(lo, up, S, seed=7)).cuda()
        f = torch.empty(S, dtype=torch.float64, device='cuda')
        row = {}
        with _cabi.Context(1, N, P) as ctx:
            ctx.set_spectrum(0, data.w, data.u, data.v, wts)
            for mode, tag in ((_cabi.REAL_ONLY, 'real_only'), (_cabi.IM_REFERENCE, 'fit_im_reference'), (_cabi.IM_SUM, 'fit_im_sum')):
                for _ in range(3):
                    ctx.objective_device(xs, S, f, fit_im=mode)
                ctx.profile(True)
                for _ in range(20):
                    ctx.objective_device(xs, S, f, fit_im=mode)
                ms, n = ctx.profile_read()
                ctx.profile(False)
                row[tag] = {'kernel_ms': ms / n, 'evals_per_s': S / (ms / n * 1e-3), 'algorithm': ctx.get_algorithm(mode)}
        out[name] = row
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
