"""Host-side driver of the on-device particle swarm.

Replaces ``pyswarm.pso`` as the reference calls it (utils.py:176-182).  The swarm
state (positions, velocities, personal/global bests) lives in the library's
context on the GPU; this module only sequences generations, feeds random numbers
when the reference's host RNG stream is to be reproduced, and polls the stop flags.

Three shapes of the same loop:
  ``pso_single``   one spectrum, one GPU (what ``nmrfit.fit`` needs);
  ``pso_batch``    B independent spectra, one swarm each, one GPU - spectra shard
                   across ranks with no communication (BASELINE config 3);
  ``pso_sharded``  one spectrum, the particles of one swarm sharded over the ranks
                   of a ``torch.distributed`` group; per generation each rank
                   contributes its best record (f, global index, x[D]) to one
                   all-gather and every rank applies the same first-index argmin
                   (BASELINE config 4).
"""
import numpy as np

from . import _cabi

STOP_TEXT = {
    _cabi.STOP_MINFUNC: 'Stopping search: Swarm best objective change less than {minfunc}',
    _cabi.STOP_MINSTEP: 'Stopping search: Swarm best position change less than {minstep}',
    _cabi.STOP_MAXITER: 'Stopping search: maximum iterations reached --> {maxiter}',
    _cabi.RUNNING: 'Stopping search: maximum iterations reached --> {maxiter}',
    _cabi.STOP_PEER_LOST: 'Stopping search: a peer rank never delivered its best record (exchange timed out)',
}


def _precision(p):
    if p in ('fp64', 'f64', _cabi.FP64, None):
        return _cabi.FP64
    if p in ('fp32', 'f32', _cabi.FP32):
        return _cabi.FP32
    raise ValueError("precision must be 'fp64' or 'fp32'")


def _fit_im_mode(fit_im):
    from .equations import _fit_im_mode as f
    return f(fit_im)


def _fused_mode(fused):
    """'auto' (default): small swarms run their generations in one cooperative launch (csrc/swarm_fused.cu);
    'off': always one set of launches per generation; 'require': fail when the fused kernel cannot run."""
    modes = {'auto': _cabi.FUSED_AUTO, None: _cabi.FUSED_AUTO, 'off': _cabi.FUSED_OFF, False: _cabi.FUSED_OFF,
             'require': _cabi.FUSED_REQUIRE, True: _cabi.FUSED_AUTO}
    if fused not in modes:
        raise ValueError("fused must be 'auto', 'off' or 'require'")
    return modes[fused]


def _check_bounds(lower, upper):
    lb = np.array(lower, dtype=np.float64)
    ub = np.array(upper, dtype=np.float64)
    # pyswarm's own argument checks, same messages
    assert lb.shape == ub.shape, 'Lower- and upper-bounds must be the same length'
    assert np.all(ub > lb), 'All upper-bound values must be greater than lower-bound values'
    return lb, ub


def _make_opts(S, maxiter, omega, phip, phig, minstep, minfunc, fit_im, seed, offset=0, spectrum_offset=0):
    o = _cabi.PsoOpts()
    o.swarmsize, o.maxiter = int(S), int(maxiter)
    o.omega, o.phip, o.phig = float(omega), float(phip), float(phig)
    o.minstep, o.minfunc = float(minstep), float(minfunc)
    o.fit_im = _fit_im_mode(fit_im)
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    o.particle_offset = int(offset)
    o.spectrum_offset = int(spectrum_offset)
    return o


def _draw_generations(rnd, n, S, D):
    """rp then rg per generation, from the legacy stream - pyswarm's order.  One call for the whole chunk: the
    legacy generator fills an array sequentially, so uniform(size=(n, 2, S, D)) consumes the stream exactly as
    n x (uniform(size=(S, D)), uniform(size=(S, D))) would."""
    r = rnd.uniform(size=(n, 2, S, D))
    return np.ascontiguousarray(r[:, 0]), np.ascontiguousarray(r[:, 1])


def pso_single(w, u, v, weights, lower, upper, fit_im=False, swarmsize=100, maxiter=100, omega=0.5, phip=0.5,
               phig=0.5, minstep=1e-8, minfunc=1e-8, rng='host', seed=0, precision='fp64', chunk=16, device=None,
               quiet=False, trace=None, tuning=None, fused='auto'):
    """Minimise the nmrfit objective for one spectrum.  Returns (x_best, f_best, info).

    rng='host' consumes ``np.random`` exactly as pyswarm does (rand(S,D) for the
    positions, rand(S,D) for the velocities, then uniform(size=(S,D)) twice per
    generation), and leaves the global stream where pyswarm would have left it.
    """
    lb, ub = _check_bounds(lower, upper)
    D = lb.size
    if D < 7 or (D - 4) % 3:
        raise ValueError('bounds must have 4 + 3*n_peaks entries')
    w = _cabi.as_f64(w)
    S = int(swarmsize)
    # 'host': numpy's global legacy stream, CONTINUED ON THE DEVICE from np.random's own state (csrc/mt19937.cu) - the
    # same numbers pyswarm would draw, without drawing and shipping them; 'host_arrays': drawn on the host and copied
    host = rng in ('host', 'host_arrays')
    on_device = rng == 'host'
    if rng not in ('host', 'host_arrays', 'device'):
        raise ValueError("rng must be 'host', 'host_arrays' or 'device'")
    with _cabi.pooled_context(1, w.size, (D - 4) // 3, device=device, precision=_precision(precision)) as ctx:
        if tuning:
            ctx.set_tuning(**tuning)
        ctx.set_fused(_fused_mode(fused))
        ctx.set_spectrum(0, w, u, v, weights)
        opts = _make_opts(S, maxiter, omega, phip, phig, minstep, minfunc, fit_im, seed)
        # the generations go through in chunks: the stop flag is polled less and less often; with host numbers an early
        # stop rewinds the stream and redraws what was used, so the chunks stay moderate there
        sizes, c, left = [], (max(1, int(chunk)) if trace is None else 1), maxiter
        while left > 0:
            sizes.append(min(c, left))
            left -= sizes[-1]
            if trace is None:
                c = min(2 * c, 256 if not host else 64)
        first = None
        if on_device:
            # one draw for the initial positions / velocities AND the first chunk's (rp, rg): pair 0, then pairs 1..n
            state0 = np.random.get_state()
            n0 = sizes[0] if sizes else 0
            r_pos, r_vel = ctx.legacy_uniform_pairs_end(ctx.legacy_uniform_pairs_begin(1 + n0, S * D))
            first = (r_pos + 8 * S * D, r_vel + 8 * S * D)
        else:
            r_pos = np.random.rand(S, D) if host else None
            r_vel = np.random.rand(S, D) if host else None
        ctx.pso_begin(lb, ub, opts, r_pos, r_vel)
        ctx.pso_commit()
        if trace is not None:
            x, f, it, stop = ctx.pso_best()
            trace.append((0, x[0].copy(), float(f[0])))
        done = 0
        # rng='host': the legacy stream continues on the device, one chunk AHEAD of the swarm - while chunk k runs, the
        # numbers of chunk k + 1 are generated on a stream of their own (legacy_uniform_pairs_begin / _end)
        pending = ctx.legacy_uniform_pairs_begin(sizes[1], S * D) if on_device and len(sizes) > 1 else None
        try:
            for k, n in enumerate(sizes):
                extra = 0                                   # pairs drawn together with this chunk's (the initial pair)
                if host:
                    if on_device and k == 0:
                        state, extra = state0, 1
                        rp, rg = first
                    elif on_device:
                        state = np.random.get_state()
                        rp, rg = ctx.legacy_uniform_pairs_end(pending)          # np.random: now after chunk k
                        pending = (ctx.legacy_uniform_pairs_begin(sizes[k + 1], S * D) if k + 1 < len(sizes) else None)
                    else:
                        state = np.random.get_state()
                        rp, rg = _draw_generations(np.random, n, S, D)
                else:
                    rp = rg = None
                running = ctx.pso_run(n, rp, rg)
                if trace is not None:
                    x, f, it, stop = ctx.pso_best()
                    trace.append((int(it[0]), x[0].copy(), float(f[0])))
                if running == 0:
                    if host:
                        # leave the legacy stream where pyswarm would have: it stops drawing
                        # at the generation that tripped the test
                        if pending is not None:
                            ctx.legacy_uniform_pairs_end(pending, advance=False)    # drawn ahead, not needed
                            pending = None
                        x, f, it, stop = ctx.pso_best()
                        used = int(it[0]) - done
                        if used < n:
                            np.random.set_state(state)
                            if used + extra > 0:
                                if on_device:
                                    ctx.legacy_uniform_pairs(used + extra, S * D)
                                else:
                                    _draw_generations(np.random, used, S, D)
                    break
                done += n
        finally:
            if pending is not None:
                ctx.legacy_uniform_pairs_end(pending, advance=False)
        x, f, it, stop = ctx.pso_best()
    info = dict(generations=int(it[0]), stop=int(stop[0]), evaluations=S * (int(it[0]) + 1))
    if not quiet:
        print(STOP_TEXT[info['stop']].format(minfunc=minfunc, minstep=minstep, maxiter=maxiter))
    return x[0].copy(), float(f[0]), info


def pso_batch(spectra, lowers, uppers, fit_im=False, swarmsize=100, maxiter=100, omega=0.5, phip=0.5, phig=0.5,
              minstep=1e-8, minfunc=1e-8, rng='device', seeds=None, seed=0, precision='fp64', chunk=16, device=None,
              tuning=None, fused='auto', ctx=None, spectrum_offset=0):
    """B independent fits in one context.  ``spectra``: sequence of (w, u, v, weights),
    all of one length; ``lowers``/``uppers``: [B][D].  With ``ctx`` (a context that already
    holds the B spectra) ``spectra`` is ignored.

    rng='host' gives spectrum b its own legacy stream ``RandomState(seeds[b])``, drawn
    in pyswarm's order - i.e. the result equals running the reference fit b after
    ``np.random.seed(seeds[b])``.  rng='device' uses Philox keyed by ``seed``.
    Returns (x_best [B][D], f_best [B], generations [B], stop [B]).
    """
    lb = np.array(lowers, dtype=np.float64)
    ub = np.array(uppers, dtype=np.float64)
    B = lb.shape[0] if ctx is not None else len(spectra)
    assert lb.shape == ub.shape and lb.ndim == 2 and lb.shape[0] == B, 'bounds must be [B][D]'
    assert np.all(ub > lb), 'All upper-bound values must be greater than lower-bound values'
    D = lb.shape[1]
    if ctx is not None:
        return _pso_batch_run(ctx, lb, ub, fit_im, swarmsize, maxiter, omega, phip, phig, minstep, minfunc, rng,
                              seeds, seed, chunk, tuning, fused, spectrum_offset)
    N = len(spectra[0][0])
    with _cabi.pooled_context(B, N, (D - 4) // 3, device=device, precision=_precision(precision)) as own:
        own.set_spectra(*[np.stack([sp[k] for sp in spectra]) for k in range(4)])
        return _pso_batch_run(own, lb, ub, fit_im, swarmsize, maxiter, omega, phip, phig, minstep, minfunc, rng,
                              seeds, seed, chunk, tuning, fused, spectrum_offset)


def _pso_batch_run(ctx, lb, ub, fit_im, swarmsize, maxiter, omega, phip, phig, minstep, minfunc, rng, seeds, seed,
                   chunk, tuning, fused, spectrum_offset=0):
    B, D = lb.shape
    S = int(swarmsize)
    host = rng == 'host'
    if host:
        if seeds is None or len(seeds) != B:
            raise ValueError("rng='host' needs one seed per spectrum")
        streams = [np.random.RandomState(int(s)) for s in seeds]
    if tuning:
        ctx.set_tuning(**tuning)
    ctx.set_fused(_fused_mode(fused))
    opts = _make_opts(S, maxiter, omega, phip, phig, minstep, minfunc, fit_im, seed, spectrum_offset=spectrum_offset)
    r_pos = r_vel = None
    if host:
        r_pos = np.empty((B, S, D))
        r_vel = np.empty((B, S, D))
        for b, st in enumerate(streams):
            r_pos[b] = st.rand(S, D)
            r_vel[b] = st.rand(S, D)
    ctx.pso_begin(lb, ub, opts, r_pos, r_vel)
    ctx.pso_commit()
    done = 0
    chunk = max(1, int(chunk))
    while done < maxiter:
        n = min(chunk, maxiter - done)
        if not host:
            chunk = min(2 * chunk, 256)                    # device RNG: poll the stop flags less and less often
        rp = rg = None
        if host:
            rp = np.empty((n, B, S, D))
            rg = np.empty((n, B, S, D))
            for b, st in enumerate(streams):
                rp[:, b], rg[:, b] = _draw_generations(st, n, S, D)
        if ctx.pso_run(n, rp, rg) == 0:
            break
        done += n
    return ctx.pso_best()


# ---- particle sharding over torch.distributed ------------------------------------------------------

class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias library-owned memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


def shard_range(total, rank, world):
    """Contiguous rank-major split of ``total`` items: (offset, count)."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def gather_records(rec, group=None):
    """All-gather one record tensor per rank into [world, n], rank order (works for
    NCCL device tensors and for gloo CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world, rec.numel()), dtype=rec.dtype, device=rec.device)
    if rec.is_cuda:
        dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    else:
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def select_record(recs):
    """The winning record among ranks: smallest f, ties to the lowest global particle
    index (np.argmin's first-occurrence rule).  Host mirror of the device commit rule,
    used by the CPU tests of the exchange logic.  recs: [world, D+2]."""
    recs = np.asarray(recs)
    order = np.lexsort((recs[:, 1], recs[:, 0]))
    return recs[order[0]]


def pso_sharded(w, u, v, weights, lower, upper, fit_im=False, swarmsize=100, maxiter=100, omega=0.5, phip=0.5,
                phig=0.5, minstep=1e-8, minfunc=1e-8, seed=0, precision='fp64', check_every=8, device=None,
                group=None, host_random=None, tuning=None, exchange='nccl'):
    """One swarm of ``swarmsize`` particles sharded over the ranks of ``group``.

    Every rank holds the spectrum and a contiguous block of particles.  Random numbers
    come from device Philox keyed by the GLOBAL particle index, so the trajectory does
    not depend on the number of ranks; ``host_random`` (dict with 'pos', 'vel' [S,D] and a
    callable 'gen'(k) -> (rp, rg) [S,D]) substitutes host numbers for lock-step tests.
    ``exchange``: 'nccl' - one ``all_gather`` of the best records per generation, then the commit kernel;
    'p2p' - the exchange + commit kernel stores the records straight into every peer's window over NVLink (CUDA
    IPC handles are exchanged once through the group) and no collective runs per generation.  Same result, bit for bit.
    Returns (x_best, f_best, info) - identical on every rank.
    """
    if exchange not in ('nccl', 'p2p'):
        raise ValueError("exchange must be 'nccl' or 'p2p'")
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lb, ub = _check_bounds(lower, upper)
    D = lb.size
    off, cnt = shard_range(swarmsize, rank, world)
    if cnt < 1:
        raise ValueError('swarmsize must be at least the number of ranks')
    dev = _cabi.default_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream().cuda_stream
    w = _cabi.as_f64(w)
    p2p = exchange == 'p2p'
    # a context with an exchange window is wired to its peers: it does not come from (or go back to) the pool
    manager = _cabi.Context(1, w.size, (D - 4) // 3, device=dev, precision=_precision(precision)) if p2p else \
        _cabi.pooled_context(1, w.size, (D - 4) // 3, device=dev, precision=_precision(precision))
    with manager as ctx:
        if tuning:
            ctx.set_tuning(**tuning)
        ctx.set_spectrum(0, w, u, v, weights)
        if p2p:
            handle, _ = ctx.peer_export(world, rank)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            ctx.peer_open(ipc_handles=handles)
            dist.barrier(group)                            # every window is mapped before anyone stores into it
        opts = _make_opts(cnt, maxiter, omega, phip, phig, minstep, minfunc, fit_im, seed, offset=off)
        sl = slice(off, off + cnt)
        r_pos = host_random['pos'][sl] if host_random else None
        r_vel = host_random['vel'][sl] if host_random else None
        ctx.pso_begin(lb, ub, opts, r_pos, r_vel, stream=stream)
        if p2p:
            ctx.pso_commit_peers(stream=stream)
        else:
            ptr, n = ctx.pso_record()
            rec = torch.as_tensor(_DeviceArray(ptr, n), device='cuda:%d' % dev)
            recs = gather_records(rec, group)
            ctx.pso_commit(recs, world, stream=stream)
        gen = 0
        stop = np.zeros(1, dtype=np.int32)
        lost = 0
        while gen < maxiter:
            n = min(check_every, maxiter - gen)
            if p2p:
                # a chunk of generations in ONE library call: three launches per generation, the record exchange inside
                # the finish kernel, no per-generation host work
                rp = rg = None
                if host_random:
                    drawn = [host_random['gen'](gen + 1 + k) for k in range(n)]
                    rp = np.ascontiguousarray(np.stack([d[0][sl] for d in drawn]))
                    rg = np.ascontiguousarray(np.stack([d[1][sl] for d in drawn]))
                _, lost = ctx.pso_run_peers(n, rp, rg, stream=stream)
                gen += n
            else:
                for _ in range(n):
                    gen += 1
                    rp, rg = host_random['gen'](gen) if host_random else (None, None)
                    rp, rg = (rp[sl], rg[sl]) if host_random else (None, None)
                    ctx.pso_advance(rp, rg, stream=stream)
                    recs = gather_records(rec, group)
                    ctx.pso_commit(recs, world, stream=stream)
            x, f, it, stop = ctx.pso_best()
            if p2p:
                # a timeout on ANY rank is everybody's failure: agree on it before anyone raises or leaves
                flag = torch.tensor([int(lost or ctx.peer_error() or stop[0] == _cabi.STOP_PEER_LOST)], device='cuda:%d' % dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
                if int(flag.item()):
                    dist.barrier(group)                    # nobody unmaps a window a peer may still store into
                    raise _cabi.NmrfitError(-3, 'record exchange over peer memory: a rank\'s wait for its peers expired '
                                                '(the failure is reported on every rank)')
            if stop[0] != _cabi.RUNNING:
                break
        x, f, it, stop = ctx.pso_best()
        if p2p:
            dist.barrier(group)                            # nobody unmaps a window a peer may still store into
    info = dict(generations=int(it[0]), stop=int(stop[0]), evaluations=int(swarmsize) * (int(it[0]) + 1),
                world=world, local_particles=cnt)
    return x[0].copy(), float(f[0]), info
