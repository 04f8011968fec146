#!/usr/bin/env python
"""Multi-GPU check of the particle-sharded swarm (one process per GPU, NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--shape c1|c2] [--swarm S] [--maxiter M]

Every rank runs `pso_sharded` on its block of particles; rank 0 also runs the same swarm unsharded on its own
GPU and asserts that the sharded result (best position, best value, generation count, stop reason) is
bit-identical - the per-generation exchange is one all-gather of (f, global index, x[D]) per rank and every
rank applies the same first-index argmin.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from nmrfit_b200 import swarm, synth, utils
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='c1')
    ap.add_argument('--swarm', type=int, default=0)
    ap.add_argument('--maxiter', type=int, default=40)
    ap.add_argument('--mode', default='particles', choices=['particles', 'spectra'])
    ap.add_argument('--exchange', default='nccl', choices=['nccl', 'p2p'])
    args = ap.parse_args()
    P, N, S, seed = {'c1': (6, 4096, 256, 1000), 'c2': (12, 32768, 4096, 2000), 'c4': (24, 65536, 8192, 4000)}[args.shape]
    S = args.swarm or S
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    if args.mode == 'spectra':
        return spectra_mode(args, rank, world, local)
    data, true = synth.multiplet(N, P, seed=seed)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    kw = dict(swarmsize=S, maxiter=args.maxiter, omega=-0.2134, phip=-0.3344, phig=2.3259, seed=77)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    x, f, info = swarm.pso_sharded(data.w, data.u, data.v, wts, lo, up, device=local, exchange=args.exchange, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # identical on every rank
    blob = torch.tensor(np.concatenate([x, [f, info['generations'], info['stop']]]), device='cuda')
    gathered = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(gathered, blob)
    same_on_all = all(torch.equal(g, gathered[0]) for g in gathered)
    ok = same_on_all
    line = dict(world=world, shape=args.shape, exchange=args.exchange, swarmsize=S, generations=info['generations'], stop=info['stop'], f=f,
                seconds=dt, evals_per_s=S * (info['generations'] + 1) / dt, identical_on_all_ranks=same_on_all)
    if rank == 0:
        x1, f1, info1 = swarm.pso_single(data.w, data.u, data.v, wts, lo, up, rng='device', quiet=True, device=local, **kw)
        line['bit_identical_to_one_gpu'] = bool(np.array_equal(x, x1) and f == f1 and
                                                info['generations'] == info1['generations'] and info['stop'] == info1['stop'])
        ok = ok and line['bit_identical_to_one_gpu']
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def spectra_mode(args, rank, world, local):
    """Spectra sharding (BASELINE config 3's flow): fit_batch_sharded over the ranks against fit_batch of the whole
    batch on rank 0 - bit-identical, because device random numbers are keyed by the global spectrum index."""
    import contextlib
    import io
    import torch
    import torch.distributed as dist
    import nmrfit_b200
    from nmrfit_b200 import synth
    B = 11                                                 # not a multiple of the rank count on purpose
    datas, los, ups = [], [], []
    for b in range(B):
        d, _ = synth.multiplet(2048, 6, seed=900 + b)
        lo, up = d.generate_solution_bounds()
        datas.append(d); los.append(lo); ups.append(up)
    opts = {'swarmsize': args.swarm or 48, 'maxiter': args.maxiter, 'rng': 'device', 'seed': 5, 'device': local}
    with contextlib.redirect_stdout(io.StringIO()):
        x, f, it, stop = nmrfit_b200.fit_batch_sharded(datas, los, ups, options=opts)
    blob = torch.tensor(np.concatenate([x.ravel(), f, it, stop]), device='cuda')
    gathered = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(gathered, blob)
    same_on_all = all(torch.equal(g, gathered[0]) for g in gathered)
    ok = same_on_all
    if rank == 0:
        with contextlib.redirect_stdout(io.StringIO()):
            fits = nmrfit_b200.fit_batch(datas, los, ups, options=opts)
        same = all(np.array_equal(fits[b].params, x[b]) and fits[b].error == f[b] and
                   fits[b].fit_info['generations'] == it[b] for b in range(B))
        ok = ok and same
        print(json.dumps(dict(world=world, mode='spectra', spectra=B, identical_on_all_ranks=same_on_all,
                              bit_identical_to_one_gpu=bool(same), generations=[int(i) for i in it])), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
