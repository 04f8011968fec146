"""Lineshapes and objective, mirroring ``nmrfit.equations`` (equations.py) with the
arithmetic on the GPU.

Same names, argument order and meaning as the reference:
``voigt`` (equations.py:115-149), ``kk_equation`` (:9-49), ``kk_relation`` / ``kk_relation_vectorized`` /
``kk_relation_parallel`` (:52-112, :242), ``objective`` (:152-212), ``laplace1d``
(:215-238).  ``objective_batch`` is the addition that makes the path fast: one
call evaluates a whole swarm generation.
"""
import numpy as np

from . import _cabi

_ctx_cache = {}


def _context(n_points, n_peaks, precision=_cabi.FP64):
    key = (_cabi.default_device(), int(n_points), int(n_peaks), precision)
    ctx = _ctx_cache.get(key)
    if ctx is None:
        if len(_ctx_cache) > 8:
            _ctx_cache.popitem()[1].close()
        ctx = _ctx_cache[key] = _cabi.Context(1, n_points, n_peaks, device=key[0], precision=precision)
    return ctx


def voigt(w, r, yoff, width, loc, a):
    """yoff + a*(r*Lorentzian + (1-r)*Gaussian) over ``w`` (area-parameterised)."""
    w = _cabi.as_f64(np.atleast_1d(w))
    out = np.empty_like(w)
    _cabi.check(_cabi.lib().nmrfit_voigt_host(_cabi.default_device(), _cabi.ptr(w), w.size, float(r), float(yoff),
                                              float(width), float(loc), float(a), _cabi.ptr(out)))
    return out


def kk_relation_vectorized(w, r, yoff, width, loc, a):
    """Kramers-Kronig counterpart of the Voigt body for every ``w`` (closed form;
    the reference integrates numerically with scipy ``quad`` per point)."""
    w = _cabi.as_f64(np.atleast_1d(w))
    out = np.empty_like(w)
    _cabi.check(_cabi.lib().nmrfit_kk_host(_cabi.default_device(), _cabi.ptr(w), w.size, float(r), float(yoff),
                                           float(width), float(loc), float(a), _cabi.ptr(out)))
    return out


def kk_equation(x, r, yoff, width, loc, a, w):
    """The integrand of the reference's Kramers-Kronig quadrature, [V(w - x) - V(w + x)] / x (equations.py:9-49).
    Kept importable for callers that integrate it themselves; ``kk_relation`` here does not need it (closed form).
    ``x`` may be an array; the two Voigt evaluations run on the GPU."""
    x = _cabi.as_f64(np.atleast_1d(x))
    v1 = voigt(x + w, r, yoff, width, loc, a)
    v2 = voigt(-x + w, r, yoff, width, loc, a)
    out = 1 / x * (v2 - v1)
    return out if out.size > 1 else float(out[0])


def kk_relation(w, r, yoff, width, loc, a):
    """Scalar form (equations.py:52-80)."""
    return float(kk_relation_vectorized(np.array([w], dtype=np.float64), r, yoff, width, loc, a)[0])


def kk_relation_parallel(w, r, yoff, width, loc, a, pool=None):
    """Same as the vectorised form; ``pool`` is accepted for signature compatibility
    (equations.py:83-112) and ignored - the GPU is the pool."""
    return kk_relation_vectorized(w, r, yoff, width, loc, a)


def _fit_im_mode(fit_im):
    # the reference tests ``fit_im is True`` (equations.py:184,198,205): any other
    # truthy value takes the real-only path.  The string 'sum' opts into the
    # accumulated imaginary fit.
    if fit_im is True:
        return _cabi.IM_REFERENCE
    if isinstance(fit_im, str) and fit_im == 'sum':
        return _cabi.IM_SUM
    return _cabi.REAL_ONLY


def objective_batch(xs, w, u, v, weights, fit_im=False, precision=_cabi.FP64):
    """``objective`` for every row of ``xs`` ([S, D]) in one launch -> ndarray [S]."""
    xs = _cabi.as_f64(xs)
    if xs.ndim != 2 or xs.shape[1] < 7 or (xs.shape[1] - 4) % 3:
        raise ValueError('xs must be [n_particles, 4 + 3*n_peaks]')
    w = _cabi.as_f64(w)
    ctx = _context(w.size, (xs.shape[1] - 4) // 3, precision)
    return ctx.objective_with_spectrum(xs, w, u, v, weights, _fit_im_mode(fit_im))


def objective(x, w, u, v, weights, fit_im=False):
    """Weighted RMSE between the phase-rotated data and the sum of peaks for ONE
    parameter vector (what pyswarm calls per particle in the reference)."""
    return float(objective_batch(np.asarray(x, dtype=np.float64)[None, :], w, u, v, weights, fit_im)[0])


def laplace1d(x, n=10, omega=0.33333333):
    """In-place 1-D Laplacian smoothing with pinned ends (host; runs once per fit
    on the weights, utils.py:223)."""
    for _ in range(n):
        x[1:-1] = (1. - omega) * x[1:-1] + omega * 0.5 * (x[2:] + x[:-2])
    return x
