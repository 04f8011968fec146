"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then
compare the CUDA path with the oracle and with the same fixtures."""
import numpy as np
import pytest

from conftest import load_golden, peaks_from_golden, relerr
from oracle import nmrfit_oracle as orc
from oracle import pso_oracle

OBJ_CASES = ['c1_4096x6', 'ragged_1000x6', 'p12_2048', 'tiny_257x6', 'p24_1536',
             'c3_16384x6', 'c2_32768x12', 'c4_65536x24']      # the last three: full BASELINE shapes


@pytest.mark.parametrize('case', OBJ_CASES)
def test_objective_bit_exact(case):
    g = load_golden('objective_' + case)
    f = orc.objective_swarm(g['xs'], g['w'], g['u'], g['v'], g['weights'])
    assert np.array_equal(f, g['f'])                     # same numpy ops in the same order
    f1 = orc.objective_swarm(g['xs'][:4], g['w'], g['u'], g['v'], np.ones_like(g['w']))
    assert np.array_equal(f1, g['f_ones'])


@pytest.mark.parametrize('case', OBJ_CASES)
def test_weights_and_bounds_bit_exact(case):
    g = load_golden('objective_' + case)
    peaks = peaks_from_golden(g)
    assert np.array_equal(orc.compute_weights(g['w'], peaks), g['weights'])
    lo, up = orc.solution_bounds(peaks)
    assert np.array_equal(lo, g['lower']) and np.array_equal(up, g['upper'])


def test_objective_true_parameters_is_noise_floor():
    g = load_golden('objective_c1_4096x6')
    # last particle = generating parameters; residual = the 1e-4 noise times the weights
    assert g['f'][-1] < 1e-3 < g['f'][:-1].min()


def test_fit_im_reference_semantics():
    """equations.py:199 overwrites I_fit per peak; the closed-form KK agrees with the
    reference's quadrature to quad's own tolerance."""
    g = load_golden('objective_fit_im_96x6')
    f = [orc.objective(x, g['w'], g['u'], g['v'], g['weights'], True) for x in g['xs']]
    assert relerr(f, g['f']) < 1e-9
    f_quad = orc.objective(g['xs'][0], g['w'], g['u'], g['v'], g['weights'], True, kk=orc.kk_quad_vectorized)
    assert f_quad == g['f'][0]
    # truthy-but-not-True takes the real-only path (identity test in the reference)
    assert orc.objective(g['xs'][0], g['w'], g['u'], g['v'], g['weights'], 1) == \
        orc.objective(g['xs'][0], g['w'], g['u'], g['v'], g['weights'], False)


def test_voigt_ps2_laplace_bit_exact():
    g = load_golden('voigt')
    for p, out in zip(g['pars'], g['out']):
        assert np.array_equal(orc.voigt(g['w'], *p), out)
    g = load_golden('ps2')
    for (p0, p1), fwd, inv in zip(g['phases'], g['fwd'], g['inv']):
        assert np.array_equal(np.stack(orc.ps2(g['u'], g['v'], p0, p1)), fwd)
        assert np.array_equal(np.stack(orc.ps2(g['u'], g['v'], p0, p1, inv=True)), inv)
    g = load_golden('laplace1d')
    assert np.array_equal(orc.laplace1d(g['x'].copy()), g['out'])
    assert np.array_equal(orc.laplace1d(g['x'].copy(), n=3, omega=0.5), g['out3'])


def test_kk_closed_matches_reference_quad():
    g = load_golden('kk')
    for p, out in zip(g['pars'], g['out']):
        closed = orc.kk_closed(g['w'], *p)
        near = np.ones(len(out), dtype=bool)
        if p[0] == 0.0:
            # Pure Gaussian: scipy's adaptive quad over [0, inf) does not find the narrow
            # peak from far away and returns 0 or a partial value (the true value there
            # is the -a/(pi d) tail, ~0.9 here).  That is a convergence failure of the
            # reference's numerics, not behaviour to reproduce; compare where it converges.
            near = np.abs(g['w'] - p[3]) < 7 * p[2]
            assert np.sum(out[~near] == 0.0) > 50
        assert np.max(np.abs(closed - out)[near]) < 2e-9 * np.max(np.abs(out))     # quad's default tolerance
    # one literal quad call, as the reference makes it
    p = g['pars'][0]
    assert orc.kk_quad(g['w'][5], *p) == g['out'][0][5]


def test_voigt_known_answers():
    # area parameterisation and peak height (SURVEY section 8c pins ii, iii)
    w = np.linspace(3.0, 3.8, 400001)
    for r in (0.0, 0.55, 1.0):
        y = orc.voigt(w, r, 0.0, 0.004, 3.4, 0.7)
        area = np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(w))
        # the Lorentzian tail outside +-0.4 ppm carries a/pi*W/0.4*r of the area
        assert abs(area - 0.7 * (1 - r * 0.004 / (np.pi * 0.4))) < 2e-5
        peak = 0.7 * (r * 2 / (np.pi * 0.004) + (1 - r) * (2 / 0.004) * np.sqrt(np.log(2) / np.pi))
        assert abs(orc.voigt(np.array([3.4]), r, 0.0, 0.004, 3.4, 0.7)[0] - peak) < 1e-12 * peak


def test_ps2_roundtrip():
    rng = np.random.default_rng(0)
    u, v = rng.normal(size=1000), rng.normal(size=1000)
    V, I = orc.ps2(u, v, 0.3, -1.1)
    u2, v2 = orc.ps2(V, I, 0.3, -1.1, inv=True)
    assert np.allclose(u2, u, atol=1e-14) and np.allclose(v2, v, atol=1e-14)


def test_generate_result():
    g = load_golden('generate_result_40x6')
    for tag, scale in (('s1', 1), ('s1_5', 1.5)):
        o = orc.generate_result(g['params'], g['w'], scale)
        assert np.array_equal(o['w'], g[tag + '_w'])
        assert np.array_equal(o['V'], g[tag + '_V'])
        assert np.array_equal(np.array(o['real_contribs']), g[tag + '_real'])
        scale_i = np.abs(g[tag + '_imag']).max()
        assert np.max(np.abs(np.array(o['imag_contribs']) - g[tag + '_imag'])) < 2e-9 * scale_i
        assert np.max(np.abs(o['I'] - g[tag + '_I'])) < 1e-8 * scale_i
        assert np.max(np.abs(o['u'] - g[tag + '_u'])) < 1e-8 * scale_i
        assert np.max(np.abs(o['v'] - g[tag + '_v'])) < 1e-8 * scale_i
    assert orc.area_fraction(g['areas_true']) == g['area_fraction_true']


@pytest.mark.parametrize('case', ['fit_lite_1024x6', 'fit_c1_4096x6'])
def test_fit_reproduces_reference_run(case):
    """Reference core.fit -> FitUtility.fit -> (restated) pso -> reference objective,
    against the oracle chain on the same legacy RNG stream."""
    g = load_golden(case)
    if case == 'fit_c1_4096x6':
        pytest.importorskip('numpy')   # ~6 s
    np.random.seed(int(g['seed']))
    tr = []
    x, f, info = pso_oracle.pso(orc.objective, g['lower'], g['upper'], args=(g['w'], g['u'], g['v'], g['weights'], False),
                                swarmsize=int(g['swarmsize']), maxiter=int(g['maxiter']), omega=-0.2134, phip=-0.3344,
                                phig=2.3259, trace=tr, quiet=True)
    assert np.array_equal(x, g['params']) and f == g['error']
    assert info['it'] == g['generations'] and info['stop'] == g['stop']
    assert np.array_equal(np.array([t[2] for t in tr]), g['trace_fg'])
    assert np.random.rand() == g['next_rand']


def test_pso_argument_checks_and_stops():
    with pytest.raises(AssertionError):
        pso_oracle.pso(lambda x: 0.0, [0, 1], [1, 1], quiet=True)
    rng = np.random.RandomState(3)
    # a flat function never improves on the initial best: runs to maxiter
    x, f, info = pso_oracle.pso(lambda x: 1.0, [0, 0], [1, 1], swarmsize=5, maxiter=7, rng=rng, quiet=True)
    assert info == dict(it=7, stop=pso_oracle.STOP_MAXITER) and f == 1.0
    # a smooth bowl stops early on minfunc or minstep and returns p_min, not g
    x, f, info = pso_oracle.pso(lambda x: float(np.sum(x * x)), [-1, -1], [1, 1], swarmsize=30, maxiter=500,
                                omega=0.5, phip=0.5, phig=0.5, rng=rng, quiet=True)
    assert info['stop'] in (pso_oracle.STOP_MINFUNC, pso_oracle.STOP_MINSTEP) and info['it'] < 500


@pytest.mark.parametrize('tag', 'abc')
def test_phase_estimation_oracle_matches_reference(tag):
    """brute scan, ACME score and the Nelder-Mead phase against the reference's outputs (phase.npz)."""
    g = load_golden('phase')
    u, v = g['u_' + tag], g['v_' + tag]
    assert orc.brute_phase(u, v)[0] == g['brute_' + tag][0]
    z = u + 1j * v
    assert np.array_equal([orc.acme_score(ph, z) for ph in g['acme_ph_' + tag]], g['acme_score_' + tag])
    assert np.array_equal(orc.approximate_phase(z), g['auto_' + tag])
    assert abs(g['brute_' + tag][0] - g['true_' + tag][0]) < 0.15     # and the estimate is near the truth
