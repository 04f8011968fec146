"""The on-device swarm against the CPU oracle loop and the reference's golden fits."""
import io
import contextlib

import numpy as np
import pytest

from conftest import load_golden, peaks_from_golden, relerr
import nmrfit_b200
from nmrfit_b200 import _cabi, swarm, synth, utils
from oracle import nmrfit_oracle as orc
from oracle import pso_oracle

pytestmark = pytest.mark.gpu
PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)


def golden_data(g):
    d = nmrfit_b200.containers.Data(g['w'], g['u'], g['v'])
    d.set_peaks(peaks_from_golden(g))
    return d


@pytest.mark.parametrize('case', ['fit_lite_1024x6', 'fit_c1_4096x6', 'fit_default_2048x6'])
def test_fit_reproduces_reference_with_same_seed(case, capsys):
    """nmrfit.fit(data, lb, ub) end to end: same legacy RNG stream -> the reference's fitted
    parameters to 1e-6 relative (north-star bar); in practice the trajectory is in lock-step
    and the agreement is ~1e-12."""
    g = load_golden(case)
    np.random.seed(int(g['seed']))
    f = nmrfit_b200.fit(golden_data(g), list(g['lower']), list(g['upper']), summary=True,
                        options={'swarmsize': int(g['swarmsize']), 'maxiter': int(g['maxiter'])})
    out = capsys.readouterr().out
    assert 'Fit Summary:' in out and 'Stopping search:' in out
    assert np.array_equal(f.weights, g['weights'])
    assert relerr(f.params, g['params']) < 1e-6
    assert abs(f.error / g['error'] - 1) < 1e-9
    assert f.fit_info['generations'] == g['generations']
    assert f.fit_info['stop'] == (g['stop'] if g['stop'] else _cabi.STOP_MAXITER)
    assert np.random.rand() == g['next_rand']          # the global stream is left where pyswarm leaves it
    assert abs(f.calculate_area_fraction() / g['area_fraction'] - 1) < 1e-6


def test_lockstep_trace_against_oracle():
    g = load_golden('fit_lite_1024x6')
    np.random.seed(11)
    ref_tr = []
    pso_oracle.pso(orc.objective, g['lower'], g['upper'], args=(g['w'], g['u'], g['v'], g['weights'], False),
                   swarmsize=31, maxiter=25, trace=ref_tr, quiet=True, **PSO)
    np.random.seed(11)
    tr = []
    swarm.pso_single(g['w'], g['u'], g['v'], g['weights'], g['lower'], g['upper'], swarmsize=31, maxiter=25,
                     trace=tr, quiet=True, **PSO)
    assert [t[0] for t in tr] == [t[0] for t in ref_tr]
    for (_, gx, gf), (_, rx, rf) in zip(tr, ref_tr):
        assert np.array_equal(gx, rx)                   # positions are bit-identical in lock-step
        assert abs(gf / rf - 1) < 1e-11


def test_swarm_state_after_one_generation_matches_numpy():
    g = load_golden('fit_lite_1024x6')
    S, D = 13, 22
    rs = np.random.RandomState(3)
    r_pos, r_vel, rp, rg = rs.rand(S, D), rs.rand(S, D), rs.rand(S, D), rs.rand(S, D)
    lb, ub = g['lower'], g['upper']
    with _cabi.Context(1, g['w'].size, 6) as ctx:
        ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
        opts = swarm._make_opts(S, 10, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0)
        ctx.pso_begin(lb, ub, opts, r_pos, r_vel)
        ctx.pso_commit()
        st = ctx.pso_state()
        x0 = lb + r_pos * (ub - lb)
        v0 = -np.abs(ub - lb) + r_vel * (np.abs(ub - lb) - -np.abs(ub - lb))
        assert np.array_equal(st['x'][0], x0) and np.array_equal(st['v'][0], v0) and np.array_equal(st['p'][0], x0)
        f0 = orc.objective_swarm(x0, g['w'], g['u'], g['v'], g['weights'])
        assert relerr(st['fx'][0], f0) < 1e-11 and np.array_equal(st['fx'], st['fp'])
        gbest = x0[np.argmin(f0)]
        ctx.pso_advance(rp, rg)
        ctx.pso_commit()
        st = ctx.pso_state()
        v1 = PSO['omega'] * v0 + PSO['phip'] * rp * (x0 - x0) + PSO['phig'] * rg * (gbest - x0)
        x1 = x0 + v1
        lo_m, hi_m = x1 < lb, x1 > ub
        x1 = x1 * (~np.logical_or(lo_m, hi_m)) + lb * lo_m + ub * hi_m
        assert np.array_equal(st['v'][0], v1) and np.array_equal(st['x'][0], x1)
        assert (lo_m | hi_m).any()                       # the clamp was exercised


def test_device_rng_fit_converges_and_is_reproducible():
    data, true = synth.multiplet(2048, 6, seed=2)
    lo, up = data.generate_solution_bounds()
    res = []
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            f = nmrfit_b200.fit(data, lo, up, summary=False, options={'rng': 'device', 'seed': 99, 'swarmsize': 204})
        res.append(f)
    assert np.array_equal(res[0].params, res[1].params) and res[0].error == res[1].error
    assert res[0].error < 1e-2 and res[0].fit_info['stop'] in (_cabi.STOP_MINFUNC, _cabi.STOP_MINSTEP)
    lo_a, up_a = np.array(lo), np.array(up)
    assert np.all(res[0].params >= lo_a) and np.all(res[0].params <= up_a)
    with contextlib.redirect_stdout(io.StringIO()):
        other = nmrfit_b200.fit(data, lo, up, summary=False, options={'rng': 'device', 'seed': 100})
    assert not np.array_equal(other.params, res[0].params)


def test_batch_of_spectra_equals_individual_fits():
    """fit_batch with per-spectrum legacy streams == one reference-style fit per spectrum."""
    B, N, S, iters = 4, 768, 20, 15
    datas, los, ups = [], [], []
    for b in range(B):
        d, _ = synth.multiplet(N, 6, seed=300 + b)
        lo, up = d.generate_solution_bounds()
        datas.append(d); los.append(lo); ups.append(up)
    seeds = [5, 6, 7, 8]
    opts = {'swarmsize': S, 'maxiter': iters, 'rng': 'host', 'seeds': seeds}
    fits = nmrfit_b200.fit_batch(datas, los, ups, options=opts)
    for b in range(B):
        np.random.seed(seeds[b])
        wts = orc.compute_weights(datas[b].w, datas[b].peaks)
        x, f, info = pso_oracle.pso(orc.objective, los[b], ups[b], args=(datas[b].w, datas[b].u, datas[b].v, wts, False),
                                    swarmsize=S, maxiter=iters, quiet=True, **PSO)
        assert relerr(fits[b].params, x) < 1e-6 and abs(fits[b].error / f - 1) < 1e-9
        assert fits[b].fit_info['generations'] == info['it']


def test_early_stop_freezes_only_the_stopped_spectrum():
    """Two spectra, loose minfunc: both stop at different generations; results equal the
    per-spectrum runs (a stopped swarm must not move while its neighbour continues)."""
    B, N, S = 2, 512, 30
    datas, los, ups = [], [], []
    for b in range(B):
        d, _ = synth.multiplet(N, 6, seed=400 + b)
        lo, up = d.generate_solution_bounds()
        datas.append(d); los.append(lo); ups.append(up)
    opts = {'swarmsize': S, 'maxiter': 300, 'rng': 'host', 'seeds': [1, 2], 'minfunc': 1e-4}
    fits = nmrfit_b200.fit_batch(datas, los, ups, options=opts)
    gens = [f.fit_info['generations'] for f in fits]
    for b in range(B):
        np.random.seed([1, 2][b])
        wts = orc.compute_weights(datas[b].w, datas[b].peaks)
        x, f, info = pso_oracle.pso(orc.objective, los[b], ups[b], args=(datas[b].w, datas[b].u, datas[b].v, wts, False),
                                    swarmsize=S, maxiter=300, minfunc=1e-4, quiet=True, **PSO)
        assert info['it'] == gens[b] and info['stop'] == fits[b].fit_info['stop']
        assert relerr(fits[b].params, x) < 1e-6
    assert gens[0] != gens[1]


def test_sharded_swarm_emulated_on_one_gpu_equals_unsharded():
    """Particle sharding: 3 'ranks' (contexts) on one GPU, records concatenated on the device as the
    all-gather would deliver them.  g/fg and the result are bit-identical to the single-context run."""
    import ctypes
    g = load_golden('fit_lite_1024x6')
    S, D, iters, ranks = 30, 22, 12, 3
    rs = np.random.RandomState(8)
    r_pos, r_vel = rs.rand(S, D), rs.rand(S, D)
    gens = [(rs.rand(S, D), rs.rand(S, D)) for _ in range(iters)]
    lb, ub = g['lower'], g['upper']

    def run_single():
        with _cabi.Context(1, g['w'].size, 6) as ctx:
            ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            ctx.pso_begin(lb, ub, swarm._make_opts(S, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0), r_pos, r_vel)
            ctx.pso_commit()
            for rp, rg in gens:
                ctx.pso_advance(rp, rg)
                ctx.pso_commit()
            return ctx.pso_best()

    def run_sharded():
        import torch
        ctxs, recs = [], []
        for r in range(ranks):
            off, cnt = swarm.shard_range(S, r, ranks)
            ctx = _cabi.Context(1, g['w'].size, 6)
            ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            o = swarm._make_opts(cnt, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0, offset=off)
            ctx.pso_begin(lb, ub, o, r_pos[off:off + cnt], r_vel[off:off + cnt])
            ptr, n = ctx.pso_record()
            recs.append(torch.as_tensor(swarm._DeviceArray(ptr, n), device='cuda:0'))
            ctxs.append((ctx, off, cnt))

        def exchange_and_commit():
            torch.cuda.synchronize()
            allrec = torch.stack(recs).contiguous()
            for ctx, _, _ in ctxs:
                ctx.pso_commit(allrec, ranks)
            torch.cuda.synchronize()
        exchange_and_commit()
        for rp, rg in gens:
            for ctx, off, cnt in ctxs:
                ctx.pso_advance(rp[off:off + cnt], rg[off:off + cnt])
            exchange_and_commit()
        outs = [ctx.pso_best() for ctx, _, _ in ctxs]
        for ctx, _, _ in ctxs:
            ctx.close()
        return outs

    x, f, it, stop = run_single()
    for xs, fs, its, stops in run_sharded():
        assert np.array_equal(xs, x) and np.array_equal(fs, f) and np.array_equal(its, it) and np.array_equal(stops, stop)


def test_philox_trajectory_independent_of_sharding():
    import torch
    g = load_golden('fit_lite_1024x6')
    S, iters = 24, 6
    lb, ub = g['lower'], g['upper']
    results = []
    for ranks in (1, 2, 4):
        ctxs, recs = [], []
        for r in range(ranks):
            off, cnt = swarm.shard_range(S, r, ranks)
            ctx = _cabi.Context(1, g['w'].size, 6)
            ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            ctx.pso_begin(lb, ub, swarm._make_opts(cnt, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 77, offset=off))
            ptr, n = ctx.pso_record()
            recs.append(torch.as_tensor(swarm._DeviceArray(ptr, n), device='cuda:0'))
            ctxs.append(ctx)
        for k in range(iters + 1):
            if k:
                for ctx in ctxs:
                    ctx.pso_advance()
            torch.cuda.synchronize()
            allrec = torch.stack(recs).contiguous()
            for ctx in ctxs:
                ctx.pso_commit(allrec, ranks)
            torch.cuda.synchronize()
        results.append(ctxs[0].pso_best())
        states = np.concatenate([c.pso_state()['x'][0] for c in ctxs])
        results[-1] = results[-1] + (states,)
        for c in ctxs:
            c.close()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)


def test_pso_argument_errors():
    g = load_golden('fit_lite_1024x6')
    with pytest.raises(AssertionError, match='greater than lower-bound'):
        swarm.pso_single(g['w'], g['u'], g['v'], g['weights'], g['upper'], g['lower'])
    with _cabi.Context(1, g['w'].size, 6) as ctx:
        with pytest.raises(_cabi.NmrfitError, match='pso_begin'):
            ctx._swarmsize = 4
            ctx.pso_advance()


def test_pso_step_equals_advance_plus_commit():
    g = load_golden('fit_lite_1024x6')
    S, D, iters = 17, 22, 6
    rs = np.random.RandomState(9)
    r_pos, r_vel = rs.rand(S, D), rs.rand(S, D)
    rp, rg = rs.rand(iters, S, D), rs.rand(iters, S, D)
    outs = []
    for use_step in (True, False):
        with _cabi.Context(1, g['w'].size, 6) as ctx:
            ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            ctx.pso_begin(g['lower'], g['upper'], swarm._make_opts(S, 100, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0), r_pos, r_vel)
            ctx.pso_commit()
            for k in range(iters):
                if use_step:
                    ctx.pso_step(rp[k], rg[k])
                else:
                    ctx.pso_advance(rp[k], rg[k])
                    ctx.pso_commit()
            outs.append((ctx.pso_best(), ctx.pso_state()))
    (ba, sa), (bb, sb) = outs
    assert all(np.array_equal(x, y) for x, y in zip(ba, bb))
    assert all(np.array_equal(sa[k], sb[k]) for k in ('x', 'v', 'p', 'fx', 'fp'))


@pytest.mark.parametrize('ranks', [2, 3])
def test_record_exchange_over_peer_memory_equals_unsharded(ranks):
    """Particle sharding with the exchange + commit kernel (records and generation tokens stored straight into every
    peer's window; here the peers are contexts of one process on one GPU, each on its own stream): same trajectory,
    bit for bit, as the unsharded swarm - and as the all-gather path, which the test above pins the same way."""
    import torch
    g = load_golden('fit_c1_4096x6')
    S, D, iters = 45, 22, 9
    rs = np.random.RandomState(4)
    r_pos, r_vel = rs.rand(S, D), rs.rand(S, D)
    rp, rg = rs.rand(iters, S, D), rs.rand(iters, S, D)
    lb, ub = g['lower'], g['upper']

    def opts(cnt, off):
        return swarm._make_opts(cnt, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0, offset=off)

    with _cabi.Context(1, g['w'].size, 6) as ctx:
        ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
        ctx.pso_begin(lb, ub, opts(S, 0), r_pos, r_vel)
        ctx.pso_commit()
        ctx.pso_run(iters, rp, rg)
        want = ctx.pso_best()
        want_x = ctx.pso_state()['x'][0]

    ctxs = [_cabi.Context(1, g['w'].size, 6) for _ in range(ranks)]
    streams = [torch.cuda.Stream() for _ in range(ranks)]
    try:
        shards = [swarm.shard_range(S, r, ranks) for r in range(ranks)]
        bases = []
        for r, c in enumerate(ctxs):
            c.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            c.set_fused(_cabi.FUSED_OFF)
            bases.append(c.peer_export(ranks, r)[1])
        for c in ctxs:
            c.peer_open(local_bases=bases)
        for r, c in enumerate(ctxs):
            off, cnt = shards[r]
            c.pso_begin(lb, ub, opts(cnt, off), r_pos[off:off + cnt], r_vel[off:off + cnt], stream=streams[r].cuda_stream)
        for r, c in enumerate(ctxs):
            c.pso_commit_peers(stream=streams[r].cuda_stream)
        for k in range(iters):
            for r, c in enumerate(ctxs):
                off, cnt = shards[r]
                c.pso_step_peers(np.ascontiguousarray(rp[k, off:off + cnt]), np.ascontiguousarray(rg[k, off:off + cnt]),
                                 stream=streams[r].cuda_stream)
        torch.cuda.synchronize()
        assert all(c.peer_error() == 0 for c in ctxs)
        outs = [c.pso_best() for c in ctxs]
        for o in outs:
            assert all(np.array_equal(a, b) for a, b in zip(o, want))
        got_x = np.concatenate([c.pso_state()['x'][0] for c in ctxs])
        assert np.array_equal(got_x, want_x)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize('ranks', [2, 4])
def test_c_level_communicator_equals_unsharded(ranks):
    """nmrfit_comm_init_all / _commit / _run: the contexts of one process (one per GPU when the box has that many, else
    all on this GPU) as the ranks of a particle-sharded swarm, driven through the C ABI alone - no torch.distributed, no
    NCCL.  Same trajectory, bit for bit, as the unsharded swarm."""
    import torch
    g = load_golden('fit_c1_4096x6')
    S, D, iters = 45, 22, 9
    rs = np.random.RandomState(4)
    r_pos, r_vel = rs.rand(S, D), rs.rand(S, D)
    rp, rg = rs.rand(iters, S, D), rs.rand(iters, S, D)
    lb, ub = g['lower'], g['upper']

    def opts(cnt, off):
        return swarm._make_opts(cnt, iters, PSO['omega'], PSO['phip'], PSO['phig'], 1e-8, 1e-8, False, 0, offset=off)

    with _cabi.Context(1, g['w'].size, 6) as ctx:
        ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
        ctx.pso_begin(lb, ub, opts(S, 0), r_pos, r_vel)
        ctx.pso_commit()
        ctx.pso_run(iters, rp, rg)
        want = ctx.pso_best()
        want_x = ctx.pso_state()['x'][0]

    n_dev = torch.cuda.device_count()
    ctxs = [_cabi.Context(1, g['w'].size, 6, device=(r % n_dev) if n_dev >= ranks else 0) for r in range(ranks)]
    try:
        shards = [swarm.shard_range(S, r, ranks) for r in range(ranks)]
        for c in ctxs:
            c.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
            c.set_fused(_cabi.FUSED_OFF)
        comm = _cabi.Communicator(ctxs)
        for r, c in enumerate(ctxs):
            off, cnt = shards[r]
            c.pso_begin(lb, ub, opts(cnt, off), r_pos[off:off + cnt], r_vel[off:off + cnt])
        comm.commit()
        for k0 in range(0, iters, 4):                      # in chunks, as a caller polling the stop flags would
            k1 = min(iters, k0 + 4)
            running, lost = comm.run(k1 - k0,
                                     [np.ascontiguousarray(rp[k0:k1, off:off + cnt]) for off, cnt in shards],
                                     [np.ascontiguousarray(rg[k0:k1, off:off + cnt]) for off, cnt in shards])
            assert lost == 0 and running in (0, 1)
        for c in ctxs:
            assert all(np.array_equal(a, b) for a, b in zip(c.pso_best(), want))
        got_x = np.concatenate([c.pso_state()['x'][0] for c in ctxs])
        assert np.array_equal(got_x, want_x)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize('pos0', [624, 0, 1, 7, 311, 622, 623])
def test_legacy_stream_continued_on_the_device_is_numpys(pos0):
    """csrc/mt19937.cu: given np.random's state the device produces the numbers np.random.rand would, for every
    position inside a 624-word block (a double's two words may straddle blocks), and hands back the advanced state."""
    import torch
    np.random.seed(12345)
    if pos0 != 624:
        np.random.random_sample(3 * 312)                    # land on a block boundary (three whole blocks) ...
        kind, key, pos, hg, cg = np.random.get_state()
        np.random.set_state((kind, key, pos0, hg, cg))     # ... then anywhere inside the block
    state = np.random.get_state()
    S, D, pairs = 37, 22, 5
    want = np.random.random_sample((2 * pairs, S * D))
    after_host = np.random.get_state()
    np.random.set_state(state)
    with _cabi.Context(1, 64, 6) as ctx:
        a, b = ctx.legacy_uniform_pairs(pairs, S * D)
        got_a = torch.as_tensor(swarm._DeviceArray(a, pairs * S * D), device='cuda').cpu().numpy().reshape(pairs, S * D)
        got_b = torch.as_tensor(swarm._DeviceArray(b, pairs * S * D), device='cuda').cpu().numpy().reshape(pairs, S * D)
    assert np.array_equal(got_a, want[0::2]) and np.array_equal(got_b, want[1::2])
    after_dev = np.random.get_state()
    assert after_dev[2] == after_host[2] and np.array_equal(after_dev[1], after_host[1])
    assert np.random.rand() == (np.random.set_state(after_host) or np.random.rand())


def test_fit_with_device_continued_stream_equals_host_drawn_arrays():
    """The default parity mode (rng='host': legacy stream continued on the device) and rng='host_arrays' (drawn on the
    host, copied) are the same fit - parameters, error, generations, and where np.random is left - including an early
    stop in the middle of a chunk."""
    import contextlib
    import io
    import nmrfit_b200
    data, true = synth.multiplet(2048, 6, seed=2)
    lo, up = data.generate_solution_bounds()
    out = []
    for rng in ('host', 'host_arrays'):
        np.random.seed(4)
        with contextlib.redirect_stdout(io.StringIO()):
            fit = nmrfit_b200.fit(data, lo, up, summary=False, options={'rng': rng})      # defaults: stops on minfunc
        out.append((fit.params.copy(), fit.error, fit.fit_info['generations'], fit.fit_info['stop'], np.random.rand()))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1:] == out[1][1:]
    assert out[0][3] in (_cabi.STOP_MINFUNC, _cabi.STOP_MINSTEP) and 0 < out[0][2] < 2000
