#!/usr/bin/env python
"""Wall time of one nmrfit_b200.fit through the public API: fused swarm kernel vs per-step kernels.

    python tools/fit_latency.py
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nmrfit_b200                                   # noqa: E402
from nmrfit_b200 import synth                        # noqa: E402


def main():
    out = {}
    sink = io.StringIO()
    for name, (N, P, opts) in {
        'c1_S100_maxiter100': (4096, 6, {'swarmsize': 100, 'maxiter': 100}),
        'defaults_S204': (4096, 6, {}),
        'n16384_S204_maxiter200': (16384, 6, {'maxiter': 200}),
        'p12_n32768_S148_maxiter100': (32768, 12, {'swarmsize': 148, 'maxiter': 100}),
    }.items():
        data, _ = synth.multiplet(N, P, seed=1000)
        lo, up = data.generate_solution_bounds()
        row = {}
        for fused in ('auto', 'off'):
            for chunk in (16, 64):
                times = []
                for rep in range(4):
                    t0 = time.perf_counter()
                    with contextlib.redirect_stdout(sink):
                        f = nmrfit_b200.fit(data, lo, up, summary=False,
                                            options=dict(opts, rng='device', seed=rep, fused=fused, chunk=chunk))
                    times.append(time.perf_counter() - t0)
                row['%s_chunk%d' % (fused, chunk)] = {'ms': 1e3 * float(np.median(times[1:])),
                                                      'generations': f.fit_info['generations'],
                                                      'us_per_generation': 1e6 * float(np.median(times[1:])) / max(1, f.fit_info['generations'])}
        out[name] = row
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
