"""Phase rotation, mirroring the reference's ``nmrfit.proc_autophase.ps2``
(proc_autophase.py:9-36).  The rotation runs on the GPU (``nmrfit_ps2_host``).

Phase ESTIMATION before the fit is also provided on the GPU:
``ps`` (degrees), ``_ps_acme_score`` and ``approximate_phase`` (proc_autophase.py:39-68, 107-187)
keep the reference's names and meaning - the Nelder-Mead simplex stays scipy's, as in the
reference, and each score evaluation is one kernel launch - plus batched forms
(``brute_phase_batch``, ``acme_score_batch``) for many spectra at once.  The interactive
``manual_ps`` and the nmrglue-style ``autops`` wrapper are not provided.
"""
import numpy as np

from . import _cabi


def ps2(u, v, p0=0.0, p1=0.0, inv=False):
    """Linear phase correction of (u, v); phases in RADIANS (the reference's
    docstring says degrees, its code does not convert: proc_autophase.py:30-31).

    Returns ``(real, imag)`` of ``(u + 1j*v) * exp(+-1j*(p0 + p1*i/size))``.
    """
    u = _cabi.as_f64(u)
    v = _cabi.as_f64(v)
    if u.shape != v.shape or u.ndim != 1:
        raise ValueError('u and v must be 1-D arrays of equal length')
    re = np.empty_like(u)
    im = np.empty_like(u)
    _cabi.check(_cabi.lib().nmrfit_ps2_host(_cabi.default_device(), _cabi.ptr(u), _cabi.ptr(v), u.size,
                                            float(p0), float(p1), int(bool(inv)), _cabi.ptr(re), _cabi.ptr(im)))
    return re, im


def ps(data, p0=0.0, p1=0.0, inv=False):
    """Linear phase correction of a COMPLEX spectrum with (p0, p1) in DEGREES
    (proc_autophase.py:39-68); returns the complex result."""
    data = np.asarray(data)
    pi = np.pi
    re, im = ps2(np.real(data), np.imag(data), p0 * pi / 180., p1 * pi / 180., inv)
    return (re + 1j * im).astype(data.dtype if np.iscomplexobj(data) else complex)


def _as_scorer(data, device=None):
    if isinstance(data, _cabi.PhaseScorer):
        return data, False
    data = np.asarray(data)
    return _cabi.PhaseScorer(np.real(data), np.imag(data), device=device), True


def _ps_acme_score(ph, data):
    """ACME phase score (Chen Li et al., J. Magn. Reson. 158 (2002) 164) of complex ``data``
    phased by ``ph = (p0, p1)`` in degrees (proc_autophase.py:142-187).  ``data`` may be a
    ``_cabi.PhaseScorer`` that already holds the spectrum on the device."""
    scorer, own = _as_scorer(data)
    try:
        rad = np.array([[ph[0] * np.pi / 180., ph[1] * np.pi / 180.]])
        return float(scorer.acme(rad)[0, 0])
    finally:
        if own:
            scorer.close()


def approximate_phase(data, fn='acme', p0=0.0, p1=0.0):
    """Automatic linear phase correction (proc_autophase.py:107-139): Nelder-Mead
    (``scipy.optimize.fmin``, as the reference) on the ACME score evaluated on the GPU.
    ``p0, p1``: initial phases in degrees.  Returns (p0, p1) in RADIANS."""
    import scipy.optimize
    if fn != 'acme' and fn is not _ps_acme_score:
        raise NotImplementedError("only the 'acme' score is provided on the GPU")
    with _cabi.PhaseScorer(np.real(data), np.imag(data)) as scorer:
        opt = scipy.optimize.fmin(_ps_acme_score, x0=[p0, p1], args=(scorer, ), disp=False)
    return opt[0] * np.pi / 180, opt[1] * np.pi / 180


def brute_phase_batch(u, v, step=np.pi / 360, device=None):
    """``Data._brute_phase`` (containers.py:98-110) for a batch: ``u, v`` [B, N] -> p0 [B]
    (p1 is 0 by construction).  All B x len(arange(-pi, pi, step)) candidates in one launch."""
    with _cabi.PhaseScorer(u, v, device=device) as scorer:
        return scorer.brute(np.arange(-np.pi, np.pi, step))


def acme_score_batch(ph_degrees, u, v, device=None):
    """ACME scores [B, K] of B spectra (``u, v`` [B, N]) for K candidate phases [(p0, p1), ...] in degrees."""
    ph = np.atleast_2d(np.asarray(ph_degrees, dtype=float))
    with _cabi.PhaseScorer(u, v, device=device) as scorer:
        return scorer.acme(ph * np.pi / 180.)
