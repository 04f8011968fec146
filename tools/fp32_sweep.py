#!/usr/bin/env python
"""BASELINE config 5b: tolerance sweep of the opt-in FP32 objective against the FP64 kernel.

    python tools/fp32_sweep.py > gpurun_out/fp32_sweep.json

For each shape: particles drawn uniformly inside the solution bounds plus particles concentrated around the
generating parameters (where the residual is small and FP32 cancellation is worst).  Reports the relative
error distribution of the FP32 objective against FP64 and both kernels' times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, synth, utils      # noqa: E402

SHAPES = {'c1': (6, 4096, 1000), 'c2': (12, 32768, 2000), 'c3': (6, 16384, 3000), 'c4': (24, 65536, 4000)}


def main():
    import torch
    out = {}
    for name, (P, N, seed) in SHAPES.items():
        data, true = synth.multiplet(N, P, seed=seed)
        wts = utils.compute_weights(data.w, data.peaks)
        lo, up = np.array(data.generate_solution_bounds())
        S = 2048
        rng = np.random.default_rng(5)
        sets = {'uniform_in_bounds': synth.particles(lo, up, S, seed=7)}
        for scale in (1e-1, 1e-2, 1e-3, 0.0):
            sets['around_truth_%g' % scale] = np.clip(true + scale * (up - lo) * rng.uniform(-0.5, 0.5, (S, len(true))), lo, up)
        res = {}
        with _cabi.Context(1, N, P) as c64, _cabi.Context(1, N, P, precision=_cabi.FP32) as c32:
            for c in (c64, c32):
                c.set_spectrum(0, data.w, data.u, data.v, wts)
            for tag, xs in sets.items():
                xd = torch.from_numpy(np.ascontiguousarray(xs)).cuda()
                vals, ms = {}, {}
                for key, c in (('fp64', c64), ('fp32', c32)):
                    f = torch.empty(S, dtype=torch.float64, device='cuda')
                    for _ in range(3):
                        c.objective_device(xd, S, f)
                    c.profile(True)
                    for _ in range(10):
                        c.objective_device(xd, S, f)
                    t, n = c.profile_read()
                    c.profile(False)
                    vals[key], ms[key] = f.cpu().numpy(), t / n
                rel = np.abs(vals['fp32'] / vals['fp64'] - 1)
                res[tag] = {'max_rel_err': float(rel.max()), 'p99_rel_err': float(np.quantile(rel, 0.99)),
                            'median_rel_err': float(np.median(rel)), 'min_objective': float(vals['fp64'].min()),
                            'fp64_ms': ms['fp64'], 'fp32_ms': ms['fp32'], 'within_1e-5': bool(rel.max() < 1e-5)}
        out[name] = dict(n_peaks=P, n_points=N, particles=S, sets=res)
        print(name, json.dumps(res), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
