"""Phase rotation, mirroring the reference's ``nmrfit.proc_autophase.ps2``
(proc_autophase.py:9-36).  The rotation runs on the GPU (``nmrfit_ps2_host``).

The remaining autophase helpers of the reference (``ps``, ``autops``,
``approximate_phase``, ACME scoring, ``manual_ps``: proc_autophase.py:39-300) are
one-shot preprocessing outside the accelerated path and
are not provided here.
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
