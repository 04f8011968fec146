"""Parity of the CUDA objective (through the C ABI) with the oracle and with the golden
vectors of the unmodified reference.  Tolerance: 1e-9 relative is the north-star bar for
the FP64 kernel; the asserts use 1e-11 so that swarm comparisons stay in lock-step."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from nmrfit_b200 import _cabi, equations, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-11
OBJ_CASES = ['c1_4096x6', 'ragged_1000x6', 'p12_2048', 'tiny_257x6', 'p24_1536',
             'c3_16384x6', 'c2_32768x12', 'c4_65536x24']      # the last three: full BASELINE shapes


@pytest.mark.parametrize('case', OBJ_CASES)
def test_objective_matches_reference_golden(case):
    g = load_golden('objective_' + case)
    f = equations.objective_batch(g['xs'], g['w'], g['u'], g['v'], g['weights'])
    assert relerr(f, g['f']) < TOL
    f1 = equations.objective_batch(g['xs'][:4], g['w'], g['u'], g['v'], np.ones_like(g['w']))
    assert relerr(f1, g['f_ones']) < TOL
    # scalar entry point, as pyswarm would call it
    assert abs(equations.objective(g['xs'][0], g['w'], g['u'], g['v'], g['weights']) / g['f'][0] - 1) < TOL


@pytest.mark.parametrize('threads,r', [(128, 2), (128, 4), (128, 8), (256, 2), (256, 4), (256, 8)])
@pytest.mark.parametrize('tb', [-1, 6, 8, 10])
def test_every_kernel_variant(threads, r, tb):
    g = load_golden('objective_ragged_1000x6')
    with _cabi.Context(1, g['w'].size, 6) as ctx:
        ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
        ctx.set_algorithm(_cabi.ALGO_GENERAL)           # the uniform-axis kernel has its own file
        for sp in (1, 3, 16):
            ctx.set_tuning(threads, r, tb, sp)
            t = ctx.get_tuning(len(g['xs']))
            assert (t['threads'], t['points_per_thread'], t['particles_per_cta']) == (threads, r, sp)
            assert relerr(ctx.objective_host(g['xs']), g['f']) < TOL


@pytest.mark.parametrize('n_points', [1, 2, 31, 33, 255, 1025, 4097])
def test_ragged_and_tiny_grids(n_points):
    data, true = synth.multiplet(max(n_points, 8), 6, seed=n_points)
    w, u, v = data.w[:n_points], data.u[:n_points], data.v[:n_points]
    wts = np.linspace(0.5, 2.0, n_points)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 5, seed=n_points)
    assert relerr(equations.objective_batch(xs, w, u, v, wts), orc.objective_swarm(xs, w, u, v, wts)) < TOL


def test_single_particle_and_many_particles():
    data, true = synth.multiplet(512, 6, seed=3)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    for S in (1, 17, 1000):
        xs = synth.particles(lo, up, S, seed=S)
        want = orc.objective_swarm(xs[:64], data.w, data.u, data.v, wts)
        got = equations.objective_batch(xs, data.w, data.u, data.v, wts)
        assert got.shape == (S,) and relerr(got[:64], want) < TOL


def test_batched_spectra_axis():
    B, N, S = 5, 700, 9
    with _cabi.Context(B, N, 6) as ctx:
        xs, want = [], []
        for b in range(B):
            data, _ = synth.multiplet(N, 6, seed=100 + b)
            wts = utils.compute_weights(data.w, data.peaks)
            ctx.set_spectrum(b, data.w, data.u, data.v, wts)
            lo, up = data.generate_solution_bounds()
            x = synth.particles(lo, up, S, seed=b)
            xs.append(x)
            want.append(orc.objective_swarm(x, data.w, data.u, data.v, wts))
        got = ctx.objective_host(np.array(xs))
        assert got.shape == (B, S) and relerr(got, np.array(want)) < TOL


def test_fit_im_modes():
    g = load_golden('objective_fit_im_96x6')
    # reference semantics (equations.py:199: last peak only), pinned by the reference's own quadrature
    f = equations.objective_batch(g['xs'], g['w'], g['u'], g['v'], g['weights'], fit_im=True)
    assert relerr(f, g['f']) < 1e-9
    want = [orc.objective(x, g['w'], g['u'], g['v'], g['weights'], True) for x in g['xs']]
    assert relerr(f, want) < TOL
    # truthy but not True -> real only
    f1 = equations.objective_batch(g['xs'], g['w'], g['u'], g['v'], g['weights'], fit_im=1)
    f0 = equations.objective_batch(g['xs'], g['w'], g['u'], g['v'], g['weights'], fit_im=False)
    assert np.array_equal(f1, f0)
    # 'sum' accumulates the Kramers-Kronig curves of all peaks (as generate_result does)
    fs = equations.objective_batch(g['xs'], g['w'], g['u'], g['v'], g['weights'], fit_im='sum')
    want = []
    for x in g['xs']:
        V, I = orc.ps2(g['u'], g['v'], x[0], x[1])
        vf = sum(orc.voigt(g['w'], x[2], x[3], *x[k:k + 3]) for k in range(4, len(x), 3))
        jf = sum(orc.kk_closed(g['w'], x[2], x[3], *x[k:k + 3]) for k in range(4, len(x), 3))
        want.append((np.sqrt(np.mean((g['weights'] * (V - vf))**2)) + np.sqrt(np.mean((g['weights'] * (I - jf))**2))) / 2)
    assert relerr(fs, want) < TOL
    # at the generating parameters the summed imaginary fit explains the data down to the weighted noise
    # floor sigma * sqrt(mean(weights^2)) (sigma = 1e-4, synth.multiplet); the reference's last-peak-only
    # imaginary fit cannot
    floor = 1e-4 * np.sqrt(np.mean(g['weights'] ** 2))
    assert fs[-1] < 1.2 * floor < f[-1]


def test_deterministic_and_tiling_independent_of_particle_tile():
    g = load_golden('objective_p12_2048')
    with _cabi.Context(1, g['w'].size, 12) as ctx:
        ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
        ctx.set_tuning(128, 4, 6, 1)
        a = ctx.objective_host(g['xs'])
        ctx.set_tuning(128, 4, 6, 7)
        b = ctx.objective_host(g['xs'])
        c = ctx.objective_host(g['xs'])
        assert np.array_equal(a, b) and np.array_equal(b, c)      # bitwise: particle tiling never changes the sum order


def test_extreme_parameters_stay_finite():
    data, true = synth.multiplet(1024, 6)
    wts = np.ones(1024)
    x = true.copy()
    xs = []
    for width in (1e-6, 1e-3, 10.0):
        y = x.copy(); y[4::3] = width; xs.append(y)
    y = x.copy(); y[5::3] = 100.0; xs.append(y)          # peaks far outside the window
    y = x.copy(); y[0] = 50.0; y[1] = -80.0; xs.append(y)  # large phases (outside +-pi)
    xs = np.array(xs)
    got = equations.objective_batch(xs, data.w, data.u, data.v, wts)
    want = orc.objective_swarm(xs, data.w, data.u, data.v, wts)
    assert np.all(np.isfinite(got)) and relerr(got, want) < 1e-9


def test_full_size_properties_c2():
    """BASELINE config 2 shape (12 peaks, 32,768 points, 4,096 particles): properties that
    need no CPU oracle at full size, plus an oracle spot-check on a few particles."""
    N, P, S = 32768, 12, 4096
    data, true = synth.multiplet(N, P, seed=2000)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, S, seed=7)
    xs[0] = true
    with _cabi.Context(1, N, P) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        f = ctx.objective_host(xs)
        assert f.shape == (S,) and np.all(np.isfinite(f)) and np.all(f > 0)
        floor = 1e-4 * np.sqrt(np.mean(wts ** 2))             # weighted noise floor (sigma = 1e-4)
        assert f[0] == f.min() and 0.8 * floor < f[0] < 1.2 * floor   # generating parameters sit on it
        idx = np.r_[0, 1, 2, S // 2, S - 1, np.random.default_rng(5).choice(S, 251, replace=False)]   # 256 particles
        assert relerr(f[idx], orc.objective_swarm(xs[idx], data.w, data.u, data.v, wts)) < TOL
        # permutation of particles permutes the result bitwise
        perm = np.random.default_rng(0).permutation(S)
        assert np.array_equal(ctx.objective_host(xs[perm]), f[perm])
        # homogeneity: scaling data, areas and yoff by c scales the objective by c (c a power of 2: exact)
        ctx.set_spectrum(0, data.w, 4 * data.u, 4 * data.v, wts)
        xs4 = xs.copy(); xs4[:, 6::3] *= 4; xs4[:, 3] *= 4
        assert np.array_equal(ctx.objective_host(xs4[:256]), 4 * f[:256])
        # zero weights -> zero objective
        ctx.set_spectrum(0, data.w, data.u, data.v, np.zeros(N))
        assert np.all(ctx.objective_host(xs[:64]) == 0)


def test_argument_errors():
    with pytest.raises(_cabi.NmrfitError):
        _cabi.Context(0, 10, 6)
    with _cabi.Context(1, 64, 6) as ctx:
        with pytest.raises(_cabi.NmrfitError, match='never set'):
            ctx.objective_host(np.zeros((2, 22)))
        with pytest.raises(ValueError):
            ctx.set_spectrum(0, np.zeros(63), np.zeros(64), np.zeros(64), np.zeros(64))
        with pytest.raises(_cabi.NmrfitError):
            ctx.set_tuning(threads=100)
    with pytest.raises(ValueError):
        equations.objective_batch(np.zeros((3, 5)), np.zeros(8), np.zeros(8), np.zeros(8), np.zeros(8))


def test_unchanged_spectrum_is_not_resent_but_a_changed_one_is():
    """The host shadow of a small context's spectrum: same arrays -> same values; an array modified IN PLACE between
    two calls must be noticed (byte comparison, not identity)."""
    data, true = synth.multiplet(1500, 6, seed=8)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 9)
    wts = utils.compute_weights(data.w, data.peaks)
    a = equations.objective_batch(xs, data.w, data.u, data.v, wts)
    b = equations.objective_batch(xs, data.w, data.u, data.v, wts)
    assert np.array_equal(a, b)
    u2 = data.u.copy()
    u2[700] += 0.25
    c = equations.objective_batch(xs, data.w, u2, data.v, wts)
    assert relerr(c, orc.objective_swarm(xs, data.w, u2, data.v, wts)) < 1e-11 and not np.array_equal(a, c)
    u2[700] -= 0.25                                        # same buffer, original content again
    d = equations.objective_batch(xs, data.w, u2, data.v, wts)
    assert np.array_equal(a, d)
    wts[10:20] *= 3.0                                      # in place
    e = equations.objective_batch(xs, data.w, u2, data.v, wts)
    assert relerr(e, orc.objective_swarm(xs, data.w, u2, data.v, wts)) < 1e-11


@pytest.mark.parametrize('case', ['c1_4096x6', 'ragged_1000x6', 'p24_1536'])
def test_page_locked_positions_are_read_in_place(case):
    """Positions in page-locked host memory are read by the prepare pass itself (no staging copy, each element crosses
    PCIe once); pageable arrays go through the copy path.  Same values either way, bit for bit - also with fit_im, whose
    evaluation re-reads the device copy the prepare pass leaves behind - and the golden tolerance holds."""
    import torch
    g = load_golden('objective_' + case)
    xs = np.ascontiguousarray(g['xs'])
    pinned = torch.empty(xs.shape, dtype=torch.float64).pin_memory().numpy()
    pinned[:] = xs
    for fit_im in (False, True):
        a = equations.objective_batch(xs, g['w'], g['u'], g['v'], g['weights'], fit_im=fit_im)
        b = equations.objective_batch(pinned, g['w'], g['u'], g['v'], g['weights'], fit_im=fit_im)
        assert np.array_equal(a, b), fit_im
    assert relerr(equations.objective_batch(pinned, g['w'], g['u'], g['v'], g['weights']), g['f']) < TOL
    # a long particle list (the sliced copy path for pageable memory) against the in-place path
    big = np.ascontiguousarray(np.tile(xs, (20000 // xs.shape[0] + 1, 1))[:20000])
    big_pinned = torch.empty(big.shape, dtype=torch.float64).pin_memory().numpy()
    big_pinned[:] = big
    assert np.array_equal(equations.objective_batch(big, g['w'], g['u'], g['v'], g['weights']),
                          equations.objective_batch(big_pinned, g['w'], g['u'], g['v'], g['weights']))
    # with page-locked positions the kernels start BEFORE the spectrum is compared with the context's copy: a spectrum
    # modified in place must still be noticed (the speculative evaluation is dropped and repeated)
    a = equations.objective_batch(pinned, g['w'], g['u'], g['v'], g['weights'])
    u2 = g['u'].copy()
    b = equations.objective_batch(pinned, g['w'], u2, g['v'], g['weights'])
    assert np.array_equal(a, b)
    u2[u2.size // 2] += 0.25
    c = equations.objective_batch(pinned, g['w'], u2, g['v'], g['weights'])
    assert not np.array_equal(a, c)
    assert relerr(c, orc.objective_swarm(xs, g['w'], u2, g['v'], g['weights'])) < TOL
    u2[u2.size // 2] -= 0.25
    assert np.array_equal(a, equations.objective_batch(pinned, g['w'], u2, g['v'], g['weights']))
