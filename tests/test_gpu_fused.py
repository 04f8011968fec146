"""The fused swarm kernel (csrc/swarm_fused.cu: every generation of a small swarm in one cooperative launch)
against the per-step kernels - bit for bit - and against the CPU oracle loop."""
import numpy as np
import pytest

from conftest import load_golden
from nmrfit_b200 import _cabi, swarm, synth, utils
from oracle import nmrfit_oracle as orc
from oracle import pso_oracle

pytestmark = pytest.mark.gpu
PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)


def _run(spectra, lbs, ubs, S, iters, mode, host_rng=True, seed=5, minfunc=1e-8, minstep=1e-8, chunks=(None,)):
    """Run `iters` generations (in the given chunking) and return (best, state, fused launches)."""
    B, N = len(spectra), len(spectra[0][0])
    D = np.asarray(lbs).shape[-1]
    rs = np.random.RandomState(seed)
    with _cabi.Context(B, N, (D - 4) // 3) as ctx:
        ctx.set_fused(mode)
        for b, sp in enumerate(spectra):
            ctx.set_spectrum(b, *sp)
        opts = swarm._make_opts(S, iters, PSO['omega'], PSO['phip'], PSO['phig'], minstep, minfunc, False, 1234)
        opts.bounds_per_spectrum = 1
        r_pos = rs.rand(B, S, D) if host_rng else None
        r_vel = rs.rand(B, S, D) if host_rng else None
        ctx.pso_begin(np.asarray(lbs), np.asarray(ubs), opts, r_pos, r_vel)
        ctx.pso_commit()
        # all uniforms drawn up front, so that the chunking does not change which numbers a generation sees
        rp_all = rs.rand(iters, B, S, D) if host_rng else None
        rg_all = rs.rand(iters, B, S, D) if host_rng else None
        done = 0
        sizes = list(chunks)
        while done < iters:
            n = sizes.pop(0) if sizes else None
            n = iters - done if n is None else min(n, iters - done)
            rp = rp_all[done:done + n] if host_rng else None
            rg = rg_all[done:done + n] if host_rng else None
            if ctx.pso_run(n, rp, rg) == 0:
                break
            done += n
        return ctx.pso_best(), ctx.pso_state(), ctx.fused_launches()


def _spectrum(N, P, seed):
    data, true = synth.multiplet(N, P, seed=seed)
    lo, up = data.generate_solution_bounds()
    return (data.w, data.u, data.v, utils.compute_weights(data.w, data.peaks)), np.array(lo), np.array(up)


def _assert_identical(a, b):
    (xa, fa, ia, sa), sta, _ = a
    (xb, fb, ib, sb), stb, _ = b
    assert np.array_equal(ia, ib) and np.array_equal(sa, sb)
    assert np.array_equal(xa, xb) and np.array_equal(fa, fb)
    for k in ('x', 'v', 'p', 'fx', 'fp'):
        assert np.array_equal(sta[k], stb[k]), k


@pytest.mark.parametrize('N,P,S,iters', [
    (4096, 6, 100, 30),        # BASELINE configs[0]: whole spectrum resident in shared memory, 16-warp CTAs
    (4096, 6, 204, 12),        # the reference's default swarm: more CTAs than SMs, two per SM
    (1000, 6, 31, 20),         # ragged axis, 4 points per thread
    (16384, 6, 64, 6),         # a cluster of 4 CTAs per particle, two resident supertiles each
    (8192, 6, 30, 8),          # cluster of 4, one supertile each
    (32768, 12, 10, 4),        # cluster of 8, two supertiles each
    (16384, 6, 204, 5),        # cluster of 2: four supertiles each, restaged tile by tile
    (5000, 6, 20, 6),          # ragged: the last CTA of the cluster owns a partial run
    (3000, 12, 40, 10),        # ragged, 12 peaks
    (2500, 36, 12, 5),         # more than 32 peaks: two near-peak mask words per region
    (300, 6, 1, 4),            # a swarm of one
])
def test_fused_equals_per_step_kernels_bitwise(N, P, S, iters):
    sp, lo, up = _spectrum(N, P, seed=11)
    fused = _run([sp], [lo], [up], S, iters, _cabi.FUSED_REQUIRE)
    steps = _run([sp], [lo], [up], S, iters, _cabi.FUSED_OFF)
    assert fused[2] == 1 and steps[2] == 0
    _assert_identical(fused, steps)
    assert fused[0][2][0] == iters                      # ran every generation


def test_fused_chunked_and_mixed_with_per_step_generations():
    """State written back by a fused launch continues seamlessly - in another fused launch or per step."""
    sp, lo, up = _spectrum(2048, 6, seed=3)
    whole = _run([sp], [lo], [up], 50, 24, _cabi.FUSED_REQUIRE)
    chunked = _run([sp], [lo], [up], 50, 24, _cabi.FUSED_REQUIRE, chunks=(1, 7, 16))
    assert whole[2] == 1 and chunked[2] == 3
    _assert_identical(whole, chunked)


def test_fused_device_rng_equals_per_step():
    sp, lo, up = _spectrum(4096, 6, seed=4)
    fused = _run([sp], [lo], [up], 100, 20, _cabi.FUSED_REQUIRE, host_rng=False, chunks=(5, 15))
    steps = _run([sp], [lo], [up], 100, 20, _cabi.FUSED_OFF, host_rng=False)
    _assert_identical(fused, steps)


def test_fused_early_stop_matches_per_step():
    sp, lo, up = _spectrum(2048, 6, seed=6)
    for kw in (dict(minfunc=1e-3), dict(minstep=2e-2, minfunc=0.0)):
        fused = _run([sp], [lo], [up], 60, 200, _cabi.FUSED_REQUIRE, **kw)
        steps = _run([sp], [lo], [up], 60, 200, _cabi.FUSED_OFF, **kw)
        _assert_identical(fused, steps)
        (x, f, it, stop), _, _ = fused
        assert stop[0] in (_cabi.STOP_MINFUNC, _cabi.STOP_MINSTEP) and 0 < it[0] < 200


def test_fused_batch_of_independent_swarms():
    """Several spectra in one launch: one barrier per spectrum, swarms stop independently."""
    specs, los, ups = zip(*[_spectrum(1536, 6, seed=20 + b) for b in range(5)])
    fused = _run(list(specs), los, ups, 40, 120, _cabi.FUSED_REQUIRE, minfunc=1e-4, chunks=(16,) * 8)
    steps = _run(list(specs), los, ups, 40, 120, _cabi.FUSED_OFF, minfunc=1e-4, chunks=(16,) * 8)
    _assert_identical(fused, steps)
    assert len(set(fused[0][2].tolist())) > 1           # they did stop at different generations


def test_fused_lockstep_with_oracle_loop():
    g = load_golden('fit_c1_4096x6')
    np.random.seed(21)
    ref_tr = []
    pso_oracle.pso(orc.objective, g['lower'], g['upper'], args=(g['w'], g['u'], g['v'], g['weights'], False),
                   swarmsize=100, maxiter=12, trace=ref_tr, quiet=True, **PSO)
    np.random.seed(21)
    x, f, info = swarm.pso_single(g['w'], g['u'], g['v'], g['weights'], g['lower'], g['upper'], swarmsize=100,
                                  maxiter=12, quiet=True, fused='require', **PSO)
    assert np.array_equal(x, ref_tr[-1][1])             # positions are bit-identical in lock-step
    assert abs(f / ref_tr[-1][2] - 1) < 1e-11


def test_fused_require_fails_for_a_swarm_that_cannot_be_resident():
    sp, lo, up = _spectrum(2048, 6, seed=1)
    with pytest.raises(_cabi.NmrfitError, match='fused'):
        _run([sp], [lo], [up], 8192, 2, _cabi.FUSED_REQUIRE, host_rng=False)
    auto = _run([sp], [lo], [up], 8192, 2, _cabi.FUSED_AUTO, host_rng=False)
    assert auto[2] == 0
