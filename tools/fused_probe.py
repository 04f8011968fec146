#!/usr/bin/env python
"""Device time per swarm generation: fused swarm kernel vs per-step kernels (CUDA events), plus the host-side
costs of one fit (context, spectrum upload, begin, best)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, swarm, synth, utils   # noqa: E402

PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)


def main():
    import torch
    out = {}
    shapes = [(4096, 6, 100), (4096, 6, 148), (4096, 6, 204), (4096, 6, 296), (2048, 6, 100), (8192, 6, 100),
              (8192, 6, 148), (8192, 6, 204), (16384, 6, 100), (16384, 6, 204), (32768, 12, 36), (32768, 12, 148)]
    for N, P, S in shapes:
        data, _ = synth.multiplet(N, P, seed=1000)
        lo, up = (np.array(a) for a in data.generate_solution_bounds())
        wts = utils.compute_weights(data.w, data.peaks)
        row = {}
        t0 = time.perf_counter()
        ctx = _cabi.Context(1, N, P)
        row['ctx_create_ms'] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        row['set_spectrum_ms'] = 1e3 * (time.perf_counter() - t0)
        for mode, tag in ((_cabi.FUSED_AUTO, 'fused'), (_cabi.FUSED_OFF, 'per_step')):
            ctx.set_fused(mode)
            opts = swarm._make_opts(S, 10 ** 9, PSO['omega'], PSO['phip'], PSO['phig'], -1.0, -1.0, False, 7)
            t0 = time.perf_counter()
            ctx.pso_begin(lo, up, opts)
            ctx.pso_commit()
            torch.cuda.synchronize()
            row[tag + '_begin_ms'] = 1e3 * (time.perf_counter() - t0)
            ctx.pso_run(8)
            for n in (1, 16, 128):
                ts = []
                for _ in range(5):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    ctx.pso_run(n)
                    ts.append(time.perf_counter() - t0)
                row['%s_run%d_us_per_gen' % (tag, n)] = 1e6 * float(np.median(ts)) / n
            t0 = time.perf_counter()
            ctx.pso_best()
            row[tag + '_best_ms'] = 1e3 * (time.perf_counter() - t0)
        # phase breakdown of the fused kernel (CTA 0), cycles per generation
        ctx.set_fused(_cabi.FUSED_AUTO)
        ctx.pso_begin(lo, up, swarm._make_opts(S, 10 ** 9, PSO['omega'], PSO['phip'], PSO['phig'], -1.0, -1.0, False, 7))
        ctx.pso_commit()
        ctx.pso_run(8)
        before = ctx.fused_launches()
        ctx.fused_timing(True)
        ctx.pso_run(256)
        cyc = ctx.fused_timing(False, read=True)
        if ctx.fused_launches() > before:
            names = ('move', 'constants', 'objective', 'tile_sums', 'publish', 'barrier', 'argmin', 'commit')
            row['fused_phase_cycles_per_gen'] = {n: float(c) / 256 for n, c in zip(names, cyc)}
        row['fused_launches'] = ctx.fused_launches()
        t0 = time.perf_counter()
        ctx.close()
        row['ctx_close_ms'] = 1e3 * (time.perf_counter() - t0)
        out['N%d_P%d_S%d' % (N, P, S)] = row
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
