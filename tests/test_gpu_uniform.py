"""The uniform-axis objective kernel (objective_uniform.cu): parity with the reference's golden
vectors and the oracle, agreement with the general kernel, kernel selection, and the parameter
regimes that leave the recurrence (narrow peaks, coarse grids, far tails)."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from nmrfit_b200 import _cabi, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-11
OBJ_CASES = ['c1_4096x6', 'ragged_1000x6', 'p12_2048', 'tiny_257x6', 'p24_1536',
             'c3_16384x6', 'c2_32768x12', 'c4_65536x24']      # the last three: full BASELINE shapes


def _ctx(g, n_peaks, algo):
    ctx = _cabi.Context(1, g['w'].size, n_peaks)
    ctx.set_spectrum(0, g['w'], g['u'], g['v'], g['weights'])
    ctx.set_algorithm(algo)
    return ctx


@pytest.mark.parametrize('case', OBJ_CASES)
def test_golden_axes_select_the_uniform_kernel_and_match(case):
    g = load_golden('objective_' + case)
    n_peaks = (g['xs'].shape[1] - 4) // 3
    with _ctx(g, n_peaks, _cabi.ALGO_AUTO) as ctx:
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM              # np.linspace axes
        assert ctx.get_algorithm(_cabi.IM_REFERENCE) == _cabi.ALGO_UNIFORM   # fit_im (reference semantics) too
        assert ctx.get_algorithm(_cabi.IM_SUM) == _cabi.ALGO_GENERAL         # the summed variant stays general
        f_uni = ctx.objective_host(g['xs'])
        assert relerr(f_uni, g['f']) < TOL
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        f_gen = ctx.objective_host(g['xs'])
        assert relerr(f_gen, g['f']) < TOL
        assert relerr(f_uni, f_gen) < 1e-11


@pytest.mark.parametrize('threads', [128, 256])
@pytest.mark.parametrize('r', [4, 8, 16])
@pytest.mark.parametrize('tb', [-1, 6, 8, 10])
def test_every_uniform_variant(threads, r, tb):
    g = load_golden('objective_ragged_1000x6')
    with _ctx(g, 6, _cabi.ALGO_UNIFORM) as ctx:
        for sp in (1, 3, 16):
            ctx.set_tuning(threads, r, tb, sp)
            t = ctx.get_tuning(len(g['xs']))
            assert (t['threads'], t['points_per_thread'], t['particles_per_cta']) == (threads, r, sp)
            assert relerr(ctx.objective_host(g['xs']), g['f']) < TOL


def test_particle_tiling_never_changes_bits():
    g = load_golden('objective_p12_2048')
    with _ctx(g, 12, _cabi.ALGO_UNIFORM) as ctx:
        out = []
        for sp in (1, 5, 16):
            ctx.set_tuning(128, 8, 6, sp)
            out.append(ctx.objective_host(g['xs']))
        assert np.array_equal(out[0], out[1]) and np.array_equal(out[1], out[2])


def test_non_uniform_axis_falls_back_and_cannot_be_forced():
    g = load_golden('objective_tiny_257x6')
    w = g['w'].copy()
    w[100] += 3e-7                                       # one displaced point
    with _cabi.Context(1, w.size, 6) as ctx:
        ctx.set_spectrum(0, w, g['u'], g['v'], g['weights'])
        assert ctx.get_algorithm() == _cabi.ALGO_GENERAL
        want = orc.objective_swarm(g['xs'], w, g['u'], g['v'], g['weights'])
        assert relerr(ctx.objective_host(g['xs']), want) < TOL
        ctx.set_algorithm(_cabi.ALGO_UNIFORM)
        with pytest.raises(_cabi.NmrfitError, match='uniform'):
            ctx.objective_host(g['xs'])
    # quadratic axis: not uniform either
    w = np.linspace(1.0, 2.0, 300) ** 2
    with _cabi.Context(1, 300, 6) as ctx:
        ctx.set_spectrum(0, w, np.ones(300), np.ones(300), np.ones(300))
        assert ctx.get_algorithm() == _cabi.ALGO_GENERAL


@pytest.mark.parametrize('n_points', [2, 3, 31, 33, 255, 1025, 4097])
def test_ragged_and_tiny_axes(n_points):
    data, true = synth.multiplet(max(n_points, 8), 6, seed=n_points)
    w, u, v = data.w[:n_points], data.u[:n_points], data.v[:n_points]
    wts = np.linspace(0.5, 2.0, n_points)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 5, seed=n_points)
    with _cabi.Context(1, n_points, 6) as ctx:
        ctx.set_spectrum(0, w, u, v, wts)
        ctx.set_algorithm(_cabi.ALGO_UNIFORM)
        for r in (4, 16):
            ctx.set_tuning(128, r, 6, 2)
            assert relerr(ctx.objective_host(xs), orc.objective_swarm(xs, w, u, v, wts)) < TOL


def test_descending_axis():
    data, true = synth.multiplet(3000, 6, seed=9)
    w, u, v = data.w[::-1].copy(), data.u[::-1].copy(), data.v[::-1].copy()
    wts = utils.compute_weights(data.w, data.peaks)[::-1].copy()
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 9, seed=1)
    with _cabi.Context(1, 3000, 6) as ctx:
        ctx.set_spectrum(0, w, u, v, wts)
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        assert relerr(ctx.objective_host(xs), orc.objective_swarm(xs, w, u, v, wts)) < TOL


def test_widths_that_leave_the_recurrence():
    """Peaks narrower than the grid, widths so small that the axis' own rounding would matter, peaks far
    outside the window, huge widths, large phases: finite and within the contract everywhere."""
    data, true = synth.multiplet(1024, 6)
    wts = np.ones(1024)
    xs = []
    for width in (1e-6, 5e-5, 2e-4, 1e-3, 0.05, 10.0):
        y = true.copy(); y[4::3] = width; xs.append(y)
    y = true.copy(); y[5::3] = 100.0; xs.append(y)
    y = true.copy(); y[5::3] = data.w[0] - 0.5; xs.append(y)
    y = true.copy(); y[0] = 50.0; y[1] = -80.0; xs.append(y)
    y = true.copy(); y[4] = 1e-6; y[7] = 10.0; xs.append(y)       # mixed: one exact-path peak among recurrence peaks
    xs = np.array(xs)
    want = orc.objective_swarm(xs, data.w, data.u, data.v, wts)
    with _cabi.Context(1, 1024, 6) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        ctx.set_algorithm(_cabi.ALGO_UNIFORM)
        for r in (4, 8, 16):
            ctx.set_tuning(128, r, 6, 4)
            got = ctx.objective_host(xs)
            assert np.all(np.isfinite(got)) and relerr(got, want) < 1e-10


def test_batched_spectra_each_with_its_own_spacing():
    B, S = 4, 7
    with _cabi.Context(B, 900, 6) as ctx:
        xs, want = [], []
        for b in range(B):
            data, _ = synth.multiplet(900, 6, seed=200 + b)
            w = data.w * (1.0 + 0.1 * b)                         # different h per spectrum
            wts = utils.compute_weights(data.w, data.peaks)
            ctx.set_spectrum(b, w, data.u, data.v, wts)
            lo, up = data.generate_solution_bounds()
            x = synth.particles(lo, up, S, seed=b)
            x[:, 4::3] *= (1.0 + 0.1 * b); x[:, 5::3] *= (1.0 + 0.1 * b)
            xs.append(x)
            want.append(orc.objective_swarm(x, w, data.u, data.v, wts))
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        assert relerr(ctx.objective_host(np.array(xs)), np.array(want)) < TOL


def test_full_size_c2_uniform_against_general_and_oracle():
    """BASELINE config 2 shape: both kernels on all 4,096 particles, oracle on a handful."""
    N, P, S = 32768, 12, 4096
    data, true = synth.multiplet(N, P, seed=2000)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, S, seed=7)
    xs[0] = true
    with _cabi.Context(1, N, P) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        f = ctx.objective_host(xs)
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        fg = ctx.objective_host(xs)
        assert relerr(f, fg) < 1e-11
        idx = np.r_[0, 1, 2, S // 2, S - 1, np.random.default_rng(5).choice(S, 251, replace=False)]   # 256 particles
        assert relerr(f[idx], orc.objective_swarm(xs[idx], data.w, data.u, data.v, wts)) < TOL
        ctx.set_algorithm(_cabi.ALGO_UNIFORM)
        perm = np.random.default_rng(0).permutation(S)
        assert np.array_equal(ctx.objective_host(xs[perm]), f[perm])


@pytest.mark.parametrize('n_peaks', [36, 66])
def test_more_than_32_peaks_uses_several_mask_words(n_peaks):
    """The near-peak mask of a region is one 32-bit word per 32 peaks: 36 peaks -> 2 words, 66 -> 3."""
    data, true = synth.multiplet(3000, n_peaks, seed=n_peaks)
    wts = utils.compute_weights(data.w, data.peaks)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 6, seed=2)
    xs[0] = true
    want = orc.objective_swarm(xs, data.w, data.u, data.v, wts)
    with _cabi.Context(1, 3000, n_peaks) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        assert ctx.get_algorithm() == _cabi.ALGO_UNIFORM
        assert relerr(ctx.objective_host(xs), want) < TOL
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        assert relerr(ctx.objective_host(xs), want) < TOL


@pytest.mark.parametrize('n_points,n_peaks', [(96, 6), (1000, 6), (4096, 12), (2500, 36)])
def test_fit_im_reference_semantics_on_the_uniform_kernel(n_points, n_peaks):
    """fit_im is True (equations.py:198-209): the imaginary residual against the LAST peak's Kramers-Kronig curve.
    Uniform-axis kernel vs general kernel vs the oracle (closed form; the golden test pins it to the reference's quad)."""
    data, true = synth.multiplet(max(n_points, 600), n_peaks, seed=31)
    w, u, v = data.w[:n_points], data.u[:n_points], data.v[:n_points]
    wts = utils.compute_weights(data.w, data.peaks)[:n_points]
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 7, seed=5)
    xs[0] = true
    xs[1, -3] = 2e-5                                     # a last peak narrower than a grid step: stored abscissae
    want = np.array([orc.objective(x, w, u, v, wts, True) for x in xs])
    with _cabi.Context(1, n_points, n_peaks) as ctx:
        ctx.set_spectrum(0, w, u, v, wts)
        assert ctx.get_algorithm(_cabi.IM_REFERENCE) == _cabi.ALGO_UNIFORM
        assert ctx.get_algorithm(_cabi.IM_SUM) == _cabi.ALGO_GENERAL
        got = ctx.objective_host(xs, _cabi.IM_REFERENCE)
        assert relerr(got, want) < TOL
        ctx.set_algorithm(_cabi.ALGO_GENERAL)
        assert relerr(ctx.objective_host(xs, _cabi.IM_REFERENCE), got) < TOL


def test_fit_with_fit_im_true_follows_the_oracle_loop():
    import io, contextlib
    import nmrfit_b200
    from oracle import pso_oracle
    data, true = synth.multiplet(1200, 6, seed=17)
    lo, up = data.generate_solution_bounds()
    wts = utils.compute_weights(data.w, data.peaks)
    PSO = dict(omega=-0.2134, phip=-0.3344, phig=2.3259)
    np.random.seed(3)
    x_ref, f_ref, info = pso_oracle.pso(orc.objective, lo, up, args=(data.w, data.u, data.v, wts, True), swarmsize=20,
                                        maxiter=15, quiet=True, **PSO)
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        fit = nmrfit_b200.fit(data, lo, up, fit_im=True, summary=False, options={'swarmsize': 20, 'maxiter': 15})
    assert np.array_equal(fit.params, x_ref) and abs(fit.error / f_ref - 1) < 1e-10


@pytest.mark.parametrize('n_points,n_peaks,n_particles', [(4096, 6, 301), (1000, 6, 37), (32768, 12, 64), (2500, 36, 9), (257, 6, 5)])
@pytest.mark.parametrize('fit_im', [_cabi.REAL_ONLY, _cabi.IM_REFERENCE])
def test_streamed_and_one_group_kernels_agree_bit_for_bit(n_points, n_peaks, n_particles, fit_im):
    """objective_stream_kernel (a CTA keeps its tile and walks many particle groups through a TMA ring, warps rotate
    over the tile's regions) against objective_uniform_kernel (one group per CTA): the same eval_region on the same
    constants, region sums added in the same order - identical bits, for every pipeline depth and group size."""
    data, true = synth.multiplet(max(n_points, 600), n_peaks, seed=n_points)
    w, u, v = data.w[:n_points], data.u[:n_points], data.v[:n_points]
    wts = utils.compute_weights(data.w, data.peaks)[:n_points]
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, n_particles, seed=3)
    xs[0] = true
    with _cabi.Context(1, n_points, n_peaks) as ctx:
        ctx.set_spectrum(0, w, u, v, wts)
        ctx.set_variant(0)
        assert ctx.get_variant(n_particles)[0] == 0
        want = ctx.objective_host(xs, fit_im)
        for stages, sp in ((0, 0), (2, 1), (2, 3), (3, 4), (4, 2), (3, 7)):
            ctx.set_variant(1, stages, 2 if stages == 4 else 0)
            ctx.set_tuning(0, 0, 0, sp)
            assert ctx.get_variant(n_particles)[0] == 1
            assert np.array_equal(ctx.objective_host(xs, fit_im), want), (stages, sp)


@pytest.mark.parametrize('n_points,n_peaks', [(4096, 6), (1000, 6), (16384, 6), (3000, 36), (40, 6)])
def test_far_field_cells_of_any_size_meet_the_contract(n_points, n_peaks):
    """The far-field polynomial lives on cells of a region (uniform_eval.cuh): whole, halves or quarters.  Whatever the
    split, either FP64 uniform-axis kernel stays within 1e-11 of the oracle, and the two kernels agree bit for bit."""
    data, true = synth.multiplet(max(n_points, 600), n_peaks, seed=77 + n_points)
    w, u, v = data.w[:n_points], data.u[:n_points], data.v[:n_points]
    wts = utils.compute_weights(data.w, data.peaks)[:n_points]
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 21, seed=4)
    xs[0] = true
    want = orc.objective_swarm(xs, w, u, v, wts)
    with _cabi.Context(1, n_points, n_peaks) as ctx:
        ctx.set_spectrum(0, w, u, v, wts)
        for cells in (0, 1, 2, 4):
            ctx.set_far_cells(cells)
            for r in (4, 8, 16):
                ctx.set_tuning(0, r, 0, 0)
                ctx.set_variant(1)
                got = ctx.objective_host(xs)
                assert relerr(got, want) < TOL, (cells, r)
                ctx.set_variant(0)
                assert np.array_equal(ctx.objective_host(xs), got), (cells, r)
