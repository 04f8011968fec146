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


def _sharded_batch_worker(rank, world, port, B, D, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from nmrfit_b200 import core

    class Fake:
        pass

    def fake_fit_batch(datas, lowers, uppers, expon, dynamic_weighting, fit_im, summary, options):
        fits = []
        for k, d in enumerate(datas):
            f = Fake()
            f.params = np.full(D, float(d)) + np.arange(D) * 1e-3       # recognisable per spectrum
            f.error = float(d) * 10.0
            f.fit_info = dict(generations=int(d) + 1, stop=int(d) % 3, seed=options['seed'], seeds=options.get('seeds'))
            fits.append(f)
        out.put((rank, len(datas), options['spectrum_offset'], options.get('seeds')))
        return fits

    datas = list(range(100, 100 + B))                     # the "spectra" are just tags here
    lo = [[0.0] * D] * B
    x, f, it, stop = core.fit_batch_sharded(datas, lo, lo, options={'seed': 7, 'seeds': list(range(B))}, fit_fn=fake_fit_batch)
    assert x.shape == (B, D) and np.array_equal(x[:, 0], np.arange(100, 100 + B, dtype=float))
    assert np.array_equal(f, 10.0 * np.arange(100, 100 + B)) and np.array_equal(it, np.arange(101, 101 + B))
    assert np.array_equal(stop, np.arange(100, 100 + B) % 3)
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [5, 2, 1])
def test_fit_batch_sharded_splits_and_gathers(B):
    """Spectra-sharded batch fit: contiguous blocks per rank (uneven and even empty), per-rank seed offsets, one final
    all-gather that hands every rank every result in spectrum order."""
    world, D = 2, 7
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_batch_worker, args=(r, world, port, B, D, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    seen = sorted(out.get(timeout=5) for _ in range(min(world, B)))
    counts = [s[1] for s in seen]
    assert sum(counts) == B and max(counts) - min(counts) <= 1 if B >= world else counts == [1]
    assert seen[0][2] == 0 and seen[0][3] == list(range(counts[0]))      # rank 0: spectrum offset 0, its slice of seeds
    if len(seen) > 1:
        assert seen[1][2] == counts[0] and seen[1][3] == list(range(counts[0], B))
