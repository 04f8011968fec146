"""Parity AT the BASELINE.json shapes (C2, C3, C4 and the 6-peak / 4,096-point shape the metric
is quoted on), CUDA path through the C ABI against the CPU oracle on the same seeded inputs.  The oracle costs
0.8 / 8 / 2.4 / 31 ms per evaluation at these shapes, so a few hundred evaluations stay within seconds; the golden
fixtures objective_c{2,3,4}_*.npz (tests/test_gpu_objective.py, test_gpu_uniform.py) pin the same shapes to the
unmodified reference itself.  Tolerance 1e-11 relative (contract: 1e-9)."""
import numpy as np
import pytest

from conftest import relerr
from nmrfit_b200 import _cabi, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-11


def _inputs(N, P, seed):
    data, true = synth.multiplet(N, P, seed=seed)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    return data, true, wts, np.array(lo), np.array(up)


def test_metric_shape_6_peaks_4096_points_65536_particles():
    """bench.py's headline workload: every particle on both kernels, 256 of them against the oracle."""
    N, P, S = 4096, 6, 65536
    data, true, wts, lo, up = _inputs(N, P, 1000)
    xs = synth.particles(lo, up, S, seed=7)
    xs[0] = true
    with _cabi.Context(1, N, P) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        f = ctx.objective_host(xs)
        assert f.shape == (S,) and np.all(np.isfinite(f)) and f[0] == f.min()
        idx = np.r_[0, S - 1, np.random.default_rng(1).choice(S, 254, replace=False)]
        assert relerr(f[idx], orc.objective_swarm(xs[idx], data.w, data.u, data.v, wts)) < TOL
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        assert relerr(ctx.objective_host(xs), f) < TOL
        ctx.set_algorithm(_cabi.ALGO_UNIFORM)
        perm = np.random.default_rng(0).permutation(S)
        assert np.array_equal(ctx.objective_host(xs[perm]), f[perm])     # bitwise independent of the particle's slot


def test_c3_batch_of_spectra_at_shape():
    """BASELINE configs[2]: spectra of the 1,024-batch (6 peaks, 16,384 points), swarm of 204 each; 8 spectra in one
    context, 16 particles of every spectrum against the oracle, all of them against the general kernel."""
    N, P, S, B = 16384, 6, 204, 8
    with _cabi.Context(B, N, P) as ctx:
        xs, cases = [], []
        for b in range(B):
            data, true, wts, lo, up = _inputs(N, P, 3000 + 127 * b)       # spread over the batch's seeds 3000..4023
            ctx.set_spectrum(b, data.w, data.u, data.v, wts)
            x = synth.particles(lo, up, S, seed=7 + b)
            x[0] = true
            xs.append(x)
            cases.append((data, wts))
        xs = np.array(xs)
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        f = ctx.objective_host(xs)
        assert f.shape == (B, S)
        for b, (data, wts) in enumerate(cases):
            idx = np.r_[0, S - 1, np.random.default_rng(b).choice(S, 14, replace=False)]
            assert relerr(f[b, idx], orc.objective_swarm(xs[b, idx], data.w, data.u, data.v, wts)) < TOL
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        assert relerr(ctx.objective_host(xs), f) < TOL


@pytest.mark.parametrize('algo', [_cabi.ALGO_UNIFORM, _cabi.ALGO_GENERAL])
def test_c4_24_peaks_65536_points_at_shape(algo):
    """BASELINE configs[3]: 24 peaks on a 65,536-point window (256 regions of 256 points per particle): 24 particles
    against the oracle on either kernel; the uniform kernel also bitwise under permutation at 2,048 particles."""
    N, P = 65536, 24
    data, true, wts, lo, up = _inputs(N, P, 4000)
    xs = synth.particles(lo, up, 24, seed=7)
    xs[0] = true
    want = orc.objective_swarm(xs, data.w, data.u, data.v, wts)
    with _cabi.Context(1, N, P) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        ctx.set_algorithm(algo)
        assert relerr(ctx.objective_host(xs), want) < TOL
        if algo == _cabi.ALGO_UNIFORM:
            big = synth.particles(lo, up, 2048, seed=11)
            big[:24] = xs
            f = ctx.objective_host(big)
            assert relerr(f[:24], want) < TOL
            perm = np.random.default_rng(0).permutation(2048)
            assert np.array_equal(ctx.objective_host(big[perm]), f[perm])
