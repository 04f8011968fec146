"""world_size-2 checks of the multi-rank host logic on CPU (gloo): record exchange order,
the first-index tie-break across ranks, and that a particle-sharded swarm (each rank
advancing its own block, one record all-gather per generation) reproduces the unsharded
oracle run exactly.  The objective here is the CPU oracle - this tests the sharding
protocol, not the kernels (those are covered on the GPU by tests/test_gpu_pso.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sharded_pso(rank, world, g, S, iters, seed):
    """pyswarm's loop with the particles split over ranks; uses the package's exchange helpers."""
    from nmrfit_b200 import swarm
    from oracle import nmrfit_oracle as orc
    lb, ub = g['lower'], g['upper']
    D = lb.size
    rs = np.random.RandomState(seed)                    # every rank draws the full stream, uses its slice
    off, cnt = swarm.shard_range(S, rank, world)
    sl = slice(off, off + cnt)

    def evaluate(xs):
        return orc.objective_swarm(xs, g['w'], g['u'], g['v'], g['weights'])

    def exchange(fp, p):
        i = int(np.argmin(fp))
        rec = torch.from_numpy(np.r_[fp[i], float(off + i), p[i]])
        allrec = swarm.gather_records(rec).numpy()
        assert allrec.shape == (world, D + 2)
        return swarm.select_record(allrec)

    x = lb + rs.rand(S, D)[sl] * (ub - lb)
    vhigh = np.abs(ub - lb)
    v = -vhigh + rs.rand(S, D)[sl] * (vhigh - -vhigh)
    fp = evaluate(x)
    p = x.copy()
    win = exchange(fp, p)
    fg, gbest = win[0], win[2:].copy()
    it = 1
    while it <= iters:
        rp, rg = rs.uniform(size=(S, D))[sl], rs.uniform(size=(S, D))[sl]
        v = PSO['omega'] * v + PSO['phip'] * rp * (p - x) + PSO['phig'] * rg * (gbest - x)
        x = x + v
        lo_m, hi_m = x < lb, x > ub
        x = x * (~np.logical_or(lo_m, hi_m)) + lb * lo_m + ub * hi_m
        fx = evaluate(x)
        better = fx < fp
        p[better] = x[better]
        fp[better] = fx[better]
        win = exchange(fp, p)
        if win[0] < fg:
            step = np.sqrt(np.sum((gbest - win[2:])**2))
            if np.abs(fg - win[0]) <= 1e-8 or step <= 1e-8:
                return win[2:].copy(), win[0], it
            gbest, fg = win[2:].copy(), win[0]
        it += 1
    return gbest, fg, iters


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nmrfit_b200 import swarm
        from conftest import load_golden
        # 1. gather order and tie-break
        rec = torch.tensor([0.5, float(10 * (world - rank)), float(rank)], dtype=torch.float64)
        allrec = swarm.gather_records(rec)
        assert allrec.shape == (world, 3) and [float(r[2]) for r in allrec] == list(range(world))
        assert swarm.select_record(allrec.numpy())[2] == world - 1          # equal f -> lowest global index wins
        # 2. sharded swarm == unsharded oracle
        g = load_golden(case)
        x, f, it = _sharded_pso(rank, world, g, 21, 8, seed=17)
        out[rank] = (x, f, it)
        # 3. spectra sharding: final gather of per-rank results, rank order
        mine = torch.full((2, 4), float(rank), dtype=torch.float64)
        parts = swarm.gather_records(mine.reshape(-1)).reshape(world, 2, 4)
        assert [float(parts[r, 0, 0]) for r in range(world)] == list(range(world))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_protocol_equals_unsharded(world):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from conftest import load_golden
    from oracle import nmrfit_oracle as orc
    from oracle import pso_oracle
    case = 'fit_lite_1024x6'
    g = load_golden(case)
    ref_x, ref_f, info = pso_oracle.pso(orc.objective, g['lower'], g['upper'],
                                        args=(g['w'], g['u'], g['v'], g['weights'], False), swarmsize=21, maxiter=8,
                                        rng=np.random.RandomState(17), quiet=True, **PSO)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
        assert sorted(out.keys()) == list(range(world))
        for r in range(world):
            x, f, it = out[r]
            assert np.array_equal(x, ref_x) and f == ref_f and it == info['it']
