#!/usr/bin/env python
"""Launch-geometry sweep of the objective kernel on a GPU (development aid).

    python tools/sweep.py [c2|c1|c4 ...]

For each workload and each (threads, points/thread, exp-table bits, particles/CTA) it times
the objective kernel alone with CUDA events (nmrfit_ctx_profile) and prints evals/s,
peak-points/s and the fraction of the measured DFMA peak under the canonical flop model.
"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from nmrfit_b200 import _cabi, synth           # noqa: E402


def main():
    import torch
    names = sys.argv[1:] or ['c2', 'c1']
    burst, sustained = _cabi.fp64_peak(0, iters=4096, repeats=20)
    print('fp64 peak: burst %.2f TFLOP/s, sustained %.2f TFLOP/s' % (burst, sustained), flush=True)
    rows = []
    for name in names:
        P, N, S, _ = bench.WORKLOADS[name]
        data, weights, lo, up, true = bench.make_inputs(name)
        xs = torch.from_numpy(synth.particles(lo, up, S, seed=7)).cuda()
        f = torch.empty(S, dtype=torch.float64, device='cuda')
        with _cabi.Context(1, N, P) as ctx:
            ctx.set_spectrum(0, data.w, data.u, data.v, weights)
            ref = None
            sps = (4, 8, 16, 32) if name != 'c1' else (1, 2, 4)
            general = [(_cabi.ALGO_GENERAL,) + t for t in itertools.product((128, 256), (4,), (6, 10), sps)]
            uniform = [(_cabi.ALGO_UNIFORM,) + t for t in itertools.product((128, 256), (4, 8, 16), (6, 10), sps)]
            for algo, threads, r, tb, sp in general + uniform:
                ctx.set_algorithm(algo)
                ctx.set_tuning(threads, r, tb, sp)
                for _ in range(2):
                    ctx.objective_device(xs, S, f)
                ctx.profile(True)
                reps = 12 if name != "c1" else 50
                for _ in range(reps):
                    ctx.objective_device(xs, S, f)
                ms, n = ctx.profile_read()
                ctx.profile(False)
                ms /= n
                got = f.cpu().numpy()
                if ref is None:
                    ref = got
                err = float(np.max(np.abs(got - ref) / np.abs(ref)))
                evals = S / (ms * 1e-3)
                frac = evals * bench.flop_per_eval(N, P) / 1e12 / sustained
                rows.append(dict(workload=name, algo=algo, threads=threads, r=r, tb=tb, sp=sp, ms=ms, evals_per_s=evals,
                                 peak_points_per_s=evals * N * P, frac=frac, rel_dev=err))
                print('%s %s T=%3d R=%2d TB=%2d SP=%2d  %8.4f ms  %.3e evals/s  %.3e pp/s  frac %.3f  dev %.1e'
                      % (name, 'uni' if algo == _cabi.ALGO_UNIFORM else 'gen', threads, r, tb, sp, ms, evals,
                         evals * N * P, frac, err), flush=True)
    best = {}
    for row in rows:
        if row['workload'] not in best or row['ms'] < best[row['workload']]['ms']:
            best[row['workload']] = row
    print('BEST', json.dumps(best))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(dict(burst=burst, sustained=sustained, rows=rows), open(os.path.join(ROOT, 'gpurun_out', 'sweep.json'), 'w'))


if __name__ == '__main__':
    main()
