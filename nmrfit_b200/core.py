"""Public entry points, mirroring ``nmrfit.core`` (core.py)."""
import numpy as np

from . import containers
from . import utils
from . import swarm as _swarm

__all__ = ['load', 'fit', 'fit_batch']


def load(path, vendor='varian'):
    """Reading spectrometer files (core.py:9-61: nmrglue + FFT) is outside the
    accelerated path; build ``containers.Data(w, u, v)`` from arrays instead."""
    raise NotImplementedError(
        'nmrfit_b200 accelerates the fit path only; load the FID with the reference package '
        '(nmrglue) and wrap the arrays in nmrfit_b200.containers.Data(w, u, v)')


def fit(data, lower, upper, expon=0.5, dynamic_weighting=True, fit_im=False, processes=1, summary=True, options={}):
    """Perform a fit of NMR spectroscopy data (same signature as core.py:64).

    Returns the ``FitUtility`` holding ``params`` and ``error``.
    """
    f = utils.FitUtility(data, lower, upper, expon, dynamic_weighting, fit_im, processes, summary, options)
    f.fit()
    return f


def fit_batch(datas, lowers, uppers, expon=0.5, dynamic_weighting=True, fit_im=False, summary=False, options={}):
    """Fit many independent spectra (equal length and peak count) at once on one GPU.

    Equivalent to ``[fit(d, lo, up, ...) for d, lo, up in zip(...)]`` but all swarms
    advance together, one launch per generation.  With options['rng'] == 'host' and
    options['seeds'] = [k_0, k_1, ...], spectrum b reproduces the reference fit run
    after ``np.random.seed(k_b)``.  Returns a list of ``FitUtility``.
    """
    fits = [utils.FitUtility(d, lo, up, expon, dynamic_weighting, fit_im, 1, summary, options)
            for d, lo, up in zip(datas, lowers, uppers)]
    spectra = []
    for f in fits:
        f.weights = f._compute_weights()
        if dynamic_weighting is False:
            f.weights = np.ones_like(f.weights)
        spectra.append((f.data.w, f.data.u, f.data.v, f.weights))
    opt = options
    x, fbest, it, stop = _swarm.pso_batch(
        spectra, lowers, uppers, fit_im=fit_im,
        swarmsize=opt.get('swarmsize', 204), maxiter=opt.get('maxiter', 2000),
        omega=opt.get('omega', -0.2134), phip=opt.get('phip', -0.3344), phig=opt.get('phig', 2.3259),
        minstep=opt.get('minstep', 1e-8), minfunc=opt.get('minfunc', 1e-8),
        rng=opt.get('rng', 'device'), seeds=opt.get('seeds'), seed=opt.get('seed', 0),
        precision=opt.get('precision', 'fp64'), chunk=opt.get('chunk', 16), device=opt.get('device'),
        fused=opt.get('fused', 'auto'))
    for b, f in enumerate(fits):
        f.params = x[b].copy()
        f.error = float(fbest[b])
        f.fit_info = dict(generations=int(it[b]), stop=int(stop[b]))
        if summary is True:
            f._print_summary()
    return fits
