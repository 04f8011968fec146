"""ps2 / voigt / Kramers-Kronig / generate_result on the GPU against the reference's golden outputs."""
import numpy as np
import pytest

from conftest import load_golden, relerr
import nmrfit_b200
from nmrfit_b200 import equations, proc_autophase, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu


def test_voigt():
    g = load_golden('voigt')
    for p, out in zip(g['pars'], g['out']):
        got = equations.voigt(g['w'], *p)
        assert np.max(np.abs(got - out)) < 1e-15 * np.max(np.abs(out)) + 1e-18
    assert equations.voigt(np.array([]), 0.5, 0, 0.004, 3.4, 1.0).size == 0


def test_ps2_forward_inverse():
    g = load_golden('ps2')
    for (p0, p1), fwd, inv in zip(g['phases'], g['fwd'], g['inv']):
        assert np.max(np.abs(np.stack(proc_autophase.ps2(g['u'], g['v'], p0, p1)) - fwd)) < 5e-15
        assert np.max(np.abs(np.stack(proc_autophase.ps2(g['u'], g['v'], p0, p1, inv=True)) - inv)) < 5e-15
    V, I = proc_autophase.ps2(g['u'], g['v'], 0.7, -0.4)
    u, v = proc_autophase.ps2(V, I, 0.7, -0.4, inv=True)
    assert np.allclose(u, g['u'], atol=1e-14) and np.allclose(v, g['v'], atol=1e-14)


def test_kk_against_reference_quadrature_and_closed_form():
    g = load_golden('kk')
    for p, out in zip(g['pars'], g['out']):
        got = equations.kk_relation_vectorized(g['w'], *p)
        want = orc.kk_closed(g['w'], *p)
        assert np.max(np.abs(got - want)) < 1e-14 * np.max(np.abs(want))
        near = np.abs(g['w'] - p[3]) < 7 * p[2] if p[0] == 0.0 else np.ones(len(out), dtype=bool)
        assert np.max(np.abs(got - out)[near]) < 2e-9 * np.max(np.abs(out))      # quad's own tolerance
    p = g['pars'][0]
    assert abs(equations.kk_relation(3.4005, *p) - 8.781235875569e-01) < 1e-11   # SURVEY Appendix A probe value
    assert np.array_equal(equations.kk_relation_parallel(g['w'], *p, pool=None), equations.kk_relation_vectorized(g['w'], *p))


def test_generate_result_matches_reference():
    g = load_golden('generate_result_40x6')
    for tag, scale in (('s1', 1), ('s1_5', 1.5)):
        data = nmrfit_b200.containers.Data(g['w'].copy(), g['u'].copy(), g['v'].copy())
        f = utils.FitUtility(data, None, None)
        f.params = g['params']
        f.generate_result(scale=scale)
        assert np.array_equal(f.w, g[tag + '_w'])
        assert len(f.real_contribs) == 6 and len(f.imag_contribs) == 6
        # elementwise: yoff and the body partly cancel in the tails, so a few ulp of the body show up
        # as ~3e-13 of the (small) sum; against the curve's own scale the agreement is at the ulp level
        real = np.array(f.real_contribs)
        assert relerr(real, g[tag + '_real']) < 2e-12
        assert np.max(np.abs(real - g[tag + '_real'])) < 1e-14 * np.abs(g[tag + '_real']).max()
        assert relerr(f.V, g[tag + '_V']) < 2e-12
        si = np.abs(g[tag + '_imag']).max()
        assert np.max(np.abs(np.array(f.imag_contribs) - g[tag + '_imag'])) < 2e-9 * si
        for name in ('I', 'u', 'v'):
            assert np.max(np.abs(getattr(f, name) - g[tag + '_' + name])) < 1e-8 * si
        # side effect on the data object (utils.py:252)
        assert data.p0 == g['params'][0] and data.p1 == g['params'][1]
        assert np.max(np.abs(data.V - g[tag + '_data_V'])) < 5e-15
        o = orc.generate_result(g['params'], g['w'], scale)
        assert relerr(f.u, o['u']) < 1e-11 and relerr(f.I, o['I']) < 1e-11


def test_generate_result_config5_shape():
    """BASELINE config 5: 24 peaks, 16,384 points, scale 16 -> 262,144-point curves."""
    data, true = synth.multiplet(16384, 24, seed=5000)
    f = utils.FitUtility(data, None, None)
    f.params = true
    f.generate_result(scale=16)
    n = 16 * 16384
    assert f.w.size == n and len(f.real_contribs) == 24 and f.real_contribs[0].shape == (n,)
    o_idx = np.random.default_rng(0).integers(0, n, 2000)
    r, yoff = true[2], true[3]
    for k in (0, 7, 23):
        width, loc, a = true[4 + 3 * k:7 + 3 * k]
        assert relerr(f.real_contribs[k][o_idx], orc.voigt(f.w[o_idx], r, yoff, width, loc, a)) < 1e-12
        want = orc.kk_closed(f.w[o_idx], r, yoff, width, loc, a)
        assert np.max(np.abs(f.imag_contribs[k][o_idx] - want)) < 1e-13 * np.abs(want).max()
    assert np.allclose(f.V, np.sum(f.real_contribs, axis=0), rtol=1e-13, atol=0)
    # the fitted curve, re-phased on its own grid, returns to (V, I)
    V2, I2 = orc.ps2(f.u, f.v, true[0], true[1])
    assert np.allclose(V2, f.V, atol=1e-13) and np.allclose(I2, f.I, atol=1e-13)


def test_kk_equation_integrand_matches_reference_form():
    """equations.kk_equation (equations.py:9-49) stays importable: the integrand [V(w-x) - V(w+x)]/x, and integrating
    it the reference's way reproduces our closed-form kk_relation."""
    import scipy.integrate
    from nmrfit_b200 import equations
    pars = (0.6, 0.003, 0.004, 3.40, 0.02)
    w0 = 3.4013
    x = np.array([1e-4, 2e-3, 0.05, 1.0])
    want = 1 / x * (orc.voigt(-x + w0, *pars) - orc.voigt(x + w0, *pars))
    got = equations.kk_equation(x, *pars, w0)
    assert np.max(np.abs(got - want)) < 1e-9 * np.max(np.abs(want))
    quad = scipy.integrate.quad(lambda t: equations.kk_equation(t, *pars, w0), 0, np.inf)[0] / np.pi
    assert abs(quad - equations.kk_relation(w0, *pars)) < 2e-8 * abs(quad)
