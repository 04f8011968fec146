#!/usr/bin/env python
"""Objective-kernel timing over launch geometries / kernel variants on one GPU (development aid).

    python tools/kernel_probe.py [metric c2 c3 c4] [--tunes "0,0,0,0;256,8,6,8;..."] [--reps 10] [--out gpurun_out/probe.json]

For each workload and each tuning (threads, points per thread, exp-table bits, particles per CTA; 0 = library default)
(optionally followed by kernel variant, pipeline stages and far-field cells per region) it times the prepare pass and the evaluation kernel separately with CUDA events on the launching stream
(nmrfit_ctx_profile_read_split), checks the values against the first tuning's, and prints peak-points/s."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from nmrfit_b200 import _cabi, synth, utils    # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument('names', nargs='*', default=['metric', 'c2', 'c3', 'c4'])
    ap.add_argument('--tunes', default='0,0,0,0')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--c3-spectra', type=int, default=128)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'probe.json'))
    args = ap.parse_args()
    tunes = [tuple(int(t) for t in s.split(',')) for s in args.tunes.split(';') if s]
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for name in args.names:
        wl = bench.WORKLOADS[name]
        P, N, S = wl['P'], wl['N'], wl['S']
        B = min(wl.get('B', 1), args.c3_spectra)
        ctx = _cabi.Context(B, N, P)
        xs = []
        for b in range(B):
            d, _ = synth.multiplet(N, P, seed=wl['seed'] + b)
            ctx.set_spectrum(b, d.w, d.u, d.v, utils.compute_weights(d.w, d.peaks))
            lo, up = d.generate_solution_bounds()
            xs.append(synth.particles(lo, up, S, seed=7 + b))
        xs = torch.from_numpy(np.array(xs)).cuda()
        f = torch.empty((B, S), dtype=torch.float64, device='cuda')
        ref = None
        for tune in tunes:
            try:
                ctx.set_tuning(*tune[:4])
                ctx.set_variant(*(tune[4:6] if len(tune) > 4 else (-1, 0)), tune[7] if len(tune) > 7 else 0)
                ctx.set_far_cells(tune[6] if len(tune) > 6 else 0)
                for _ in range(3):
                    ctx.objective_device(xs, S, f)
                torch.cuda.synchronize()
                ctx.profile(True)
                for _ in range(args.reps):
                    flush.fill_(1)
                    ctx.objective_device(xs, S, f)
                prep, ev, n = ctx.profile_read_split()
                ctx.profile(False)
                got = f.cpu().numpy()
                if ref is None:
                    ref = got
                dev = float(np.max(np.abs(got - ref) / np.abs(ref)))
                row = dict(workload=name, tune=tune, picked=ctx.get_tuning(S), variant=ctx.get_variant(S), prepare_ms=prep / n, eval_ms=ev / n,
                           peak_points_per_s=B * S * N * P / ((prep + ev) / n * 1e-3), evals_per_s=B * S / ((prep + ev) / n * 1e-3),
                           rel_dev_vs_first=dev)
                rows.append(row)
                print('%-7s tune=%-25s v=%s sp=%2d  prepare %8.4f ms  eval %8.4f ms  %.3e pp/s  %.3e evals/s  dev %.1e'
                      % (name, tune, row['variant'], row['picked']['particles_per_cta'], row['prepare_ms'], row['eval_ms'],
                         row['peak_points_per_s'], row['evals_per_s'], dev), flush=True)
            except Exception as e:                          # a tuning the shape does not admit
                print('%-7s tune=%-16s ERR %s' % (name, tune, str(e)[:120]), flush=True)
        ctx.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
