"""A ``pyswarm``-shaped module: ``pso`` with pyswarm's signature, running the nmrfit objective's swarm on the GPU.

The reference calls ``pyswarm.pso(equations.objective, lower, upper, args=(w, u, v, weights, fit_im), swarmsize=...,
maxiter=..., omega=..., phip=..., phig=..., processes=...)`` (utils.py:176-182).  Binding this module under the name
``pyswarm`` before the reference is imported -

    import sys, nmrfit_b200.pyswarm_compat
    sys.modules['pyswarm'] = nmrfit_b200.pyswarm_compat
    import nmrfit                      # the UNMODIFIED reference package

- makes the reference's own ``nmrfit.fit`` / ``FitUtility.fit`` run its swarm through libnmrfit_b200.so with no change
to the reference at all.  Random numbers are drawn from numpy's global legacy stream in pyswarm's order, so
``np.random.seed(k)`` gives the trajectory the CPU path would have taken.

Only the nmrfit objective is accelerated; any other ``func`` raises (there is no CPU fallback here).
"""
import numpy as np

from . import swarm as _swarm


def _is_nmrfit_objective(func):
    """True only for the reference package's own ``nmrfit.equations.objective`` (or this package's mirror of it): the
    function object must be the ``objective`` attribute of an imported module named exactly ``nmrfit.equations`` /
    ``nmrfit_b200.equations`` - a look-alike from any other ``*.equations`` module is not silently replaced."""
    import sys
    mod_name = getattr(func, '__module__', None)
    if mod_name not in ('nmrfit.equations', 'nmrfit_b200.equations') or getattr(func, '__qualname__', '') != 'objective':
        return False
    mod = sys.modules.get(mod_name)
    return mod is not None and getattr(mod, 'objective', None) is func


def pso(func, lb, ub, ieqcons=[], f_ieqcons=None, args=(), kwargs={}, swarmsize=100, omega=0.5, phip=0.5, phig=0.5,
        maxiter=100, minstep=1e-8, minfunc=1e-8, debug=False, processes=1, particle_output=False):
    """pyswarm.pso for ``func = nmrfit.equations.objective``: returns ``(xopt, fopt)`` as pyswarm does."""
    if not _is_nmrfit_objective(func):
        raise NotImplementedError('nmrfit_b200.pyswarm_compat.pso only runs the nmrfit objective '
                                  '(equations.objective); use the real pyswarm for other functions')
    if ieqcons or f_ieqcons is not None:
        raise NotImplementedError('constraints are not used by nmrfit and are not supported')
    if particle_output:
        raise NotImplementedError('particle_output is not supported')
    if len(args) < 4:
        raise ValueError('args must be (w, u, v, weights[, fit_im]) as nmrfit passes them')
    w, u, v, weights = args[:4]
    fit_im = args[4] if len(args) > 4 else kwargs.get('fit_im', False)
    x, f, info = _swarm.pso_single(w, u, v, weights, lb, ub, fit_im=fit_im, swarmsize=swarmsize, maxiter=maxiter,
                                   omega=omega, phip=phip, phig=phig, minstep=minstep, minfunc=minfunc, rng='host')
    return x, f
