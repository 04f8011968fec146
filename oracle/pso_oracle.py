"""Restatement of ``pyswarm.pso`` (synchronous global-best particle swarm).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  pyswarm is not part of /root/reference and is not installable
in this image: the reference's README.md:13-17 tells users to pip-install git
master of https://github.com/tisimst/pyswarm (no version pin; requirements.txt
omits it; the ``processes=`` keyword used at nmrfit/utils.py:182 exists only on
master, not in the PyPI 0.6 release).  The reference has no test that pins a fit
result.  What follows restates the published algorithm of pyswarm/pso.py and is
anchored on the reference's only call site, nmrfit/utils.py:176-182:

    pyswarm.pso(objective, lower, upper, args=(w, u, v, weights, fit_im),
                swarmsize=204, maxiter=2000, omega=-0.2134, phip=-0.3344,
                phig=2.3259, processes=processes)     # minstep/minfunc default 1e-8

Draw order on numpy's legacy global MT19937 stream (this is what "the same
seeded host RNG stream" means for end-to-end parity):
    1. rand(S, D)            initial positions
    2. rand(S, D)            initial velocities
    3. per generation: uniform(size=(S, D)) for rp, then again for rg
"""
import numpy as np

STOP_MAXITER = 0
STOP_MINFUNC = 1
STOP_MINSTEP = 2


def pso(func, lb, ub, args=(), swarmsize=100, omega=0.5, phip=0.5, phig=0.5,
        maxiter=100, minstep=1e-8, minfunc=1e-8, evaluate=None, trace=None,
        quiet=False, rng=None):
    """Minimise ``func(x, *args)`` inside the box [lb, ub].

    ``evaluate`` (optional) maps the whole position matrix to the vector of
    objective values in one call; the default loops ``func`` over particles in
    index order, which is what pyswarm does with processes=1.  ``trace`` is an
    optional list that receives ``(it, g.copy(), fg)`` after every generation
    (generation 0 = initial swarm) for lock-step comparisons.  ``rng`` defaults
    to numpy's legacy global stream (``np.random``), as in pyswarm.

    Returns ``(x_best, f_best, info)`` where info = dict(it=, stop=).
    """
    rnd = np.random if rng is None else rng
    lb = np.array(lb, dtype=float)
    ub = np.array(ub, dtype=float)
    assert len(lb) == len(ub), 'Lower- and upper-bounds must be the same length'
    assert np.all(ub > lb), 'All upper-bound values must be greater than lower-bound values'

    if evaluate is None:
        def evaluate(xs):
            out = np.zeros(xs.shape[0])
            for i in range(xs.shape[0]):
                out[i] = func(xs[i, :], *args)
            return out

    vhigh = np.abs(ub - lb)
    vlow = -vhigh
    S, D = swarmsize, len(lb)

    x = rnd.rand(S, D)
    fp = np.ones(S) * np.inf
    p = np.zeros_like(x)
    fg = np.inf
    x = lb + x * (ub - lb)

    fx = evaluate(x)
    better = fx < fp
    p[better, :] = x[better, :].copy()
    fp[better] = fx[better]

    i_min = np.argmin(fp)
    if fp[i_min] < fg:
        fg = fp[i_min]
        g = p[i_min, :].copy()
    else:
        g = x[0, :].copy()

    v = vlow + rnd.rand(S, D) * (vhigh - vlow)
    if trace is not None:
        trace.append((0, g.copy(), fg))

    it = 1
    while it <= maxiter:
        rp = rnd.uniform(size=(S, D))
        rg = rnd.uniform(size=(S, D))
        v = omega * v + phip * rp * (p - x) + phig * rg * (g - x)
        x = x + v
        below = x < lb
        above = x > ub
        x = x * (~np.logical_or(below, above)) + lb * below + ub * above

        fx = evaluate(x)
        better = fx < fp
        p[better, :] = x[better, :].copy()
        fp[better] = fx[better]

        i_min = np.argmin(fp)
        if fp[i_min] < fg:
            p_min = p[i_min, :].copy()
            stepsize = np.sqrt(np.sum((g - p_min)**2))
            if np.abs(fg - fp[i_min]) <= minfunc:
                if not quiet:
                    print('Stopping search: Swarm best objective change less than {:}'.format(minfunc))
                if trace is not None:
                    trace.append((it, p_min.copy(), fp[i_min]))
                return p_min, fp[i_min], dict(it=it, stop=STOP_MINFUNC)
            elif stepsize <= minstep:
                if not quiet:
                    print('Stopping search: Swarm best position change less than {:}'.format(minstep))
                if trace is not None:
                    trace.append((it, p_min.copy(), fp[i_min]))
                return p_min, fp[i_min], dict(it=it, stop=STOP_MINSTEP)
            else:
                g = p_min.copy()
                fg = fp[i_min]
        if trace is not None:
            trace.append((it, g.copy(), fg))
        it += 1

    if not quiet:
        print('Stopping search: maximum iterations reached --> {:}'.format(maxiter))
    return g, fg, dict(it=maxiter, stop=STOP_MAXITER)
