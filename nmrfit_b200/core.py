"""Public entry points, mirroring ``nmrfit.core`` (core.py)."""
import numpy as np

from . import containers
from . import utils
from . import swarm as _swarm
from . import _cabi

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
    opt = options
    B = len(fits)
    lb = np.array(lowers, dtype=np.float64)
    n_peaks = (lb.shape[1] - 4) // 3
    W, U, V = (np.stack([_cabi.as_f64(getattr(f.data, k)) for f in fits]) for k in ('w', 'u', 'v'))
    with _cabi.pooled_context(B, W.shape[1], n_peaks, device=opt.get('device'),
                              precision=_swarm._precision(opt.get('precision', 'fp64'))) as ctx:
        # weights (utils.py:191-224) are computed on the device, next to the spectra; as in the reference they
        # are computed first and only then replaced by ones when dynamic weighting is off (utils.py:171-173)
        ctx.set_spectra(W, U, V)
        bounds, values = utils.peak_windows([f.data.peaks for f in fits], expon)
        weights = ctx.compute_weights(bounds, values)
        if dynamic_weighting is False:
            weights = np.ones_like(weights)
            ctx.set_spectra(W, U, V, weights)
        x, fbest, it, stop = _swarm.pso_batch(
            None, lowers, uppers, fit_im=fit_im,
            swarmsize=opt.get('swarmsize', 204), maxiter=opt.get('maxiter', 2000),
            omega=opt.get('omega', -0.2134), phip=opt.get('phip', -0.3344), phig=opt.get('phig', 2.3259),
            minstep=opt.get('minstep', 1e-8), minfunc=opt.get('minfunc', 1e-8),
            rng=opt.get('rng', 'device'), seeds=opt.get('seeds'), seed=opt.get('seed', 0),
            chunk=opt.get('chunk', 16), fused=opt.get('fused', 'auto'), ctx=ctx)
    for b, f in enumerate(fits):
        f.weights = weights[b]
        f.params = x[b].copy()
        f.error = float(fbest[b])
        f.fit_info = dict(generations=int(it[b]), stop=int(stop[b]))
        if summary is True:
            f._print_summary()
    return fits
