"""Restatement of ``peakutils.baseline`` (iterative polynomial baseline) for the auto peak selector.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  peakutils is a third-party dependency of the reference (requirements.txt:4, no version pin;
imported at nmrfit/utils.py:5, called at utils.py:719 and :766 as ``peakutils.baseline(y, 0)[0]``) that is absent
from /root/reference and from this image, and the reference has no test that pins a peak list.  What follows restates
the published algorithm of ``peakutils/baseline.py`` (Lucas Hermann Negri, MIT licence; versions 1.1 - 1.3):

    coeffs = ones(deg + 1); cond = abs(y).max() ** (1 / (deg + 1)); x = linspace(0, cond, y.size)
    vander = np.vander(x, deg + 1); vander_pinv = pinv(vander); base = y.copy()
    repeat up to max_it (100) times:
        coeffs_new = vander_pinv @ y
        if norm(coeffs_new - coeffs) / norm(coeffs) < tol (1e-3): break
        coeffs = coeffs_new; base = vander @ coeffs; y = minimum(y, base)
    return base

(upstream computes the pseudo-inverse with scipy.linalg.pinv2, which scipy >= 1.9 no longer has; numpy's pinv is the
same Moore-Penrose inverse to rounding).  The reference only ever calls it with deg = 0, where the fit is the mean of
the clipped signal: the loop is "clip to the running mean until the mean moves by less than 0.1 %".  Note the two
quirks the restatement keeps: on convergence ``base`` is the PREVIOUS iteration's polynomial (the update follows the
test), and if the very first mean is within 0.1 % of 1.0 (the initial coefficient) ``base`` is the signal itself.
"""
import math

import numpy as np


def baseline(y, deg=None, max_it=None, tol=None):
    if deg is None:
        deg = 3
    if max_it is None:
        max_it = 100
    if tol is None:
        tol = 1e-3
    y = np.asarray(y, dtype=float)
    order = deg + 1
    coeffs = np.ones(order)
    cond = math.pow(abs(y).max(), 1. / order)
    x = np.linspace(0., cond, y.size)
    base = y.copy()
    vander = np.vander(x, order)
    vander_pinv = np.linalg.pinv(vander)
    for _ in range(max_it):
        coeffs_new = np.dot(vander_pinv, y)
        if np.linalg.norm(coeffs_new - coeffs) / np.linalg.norm(coeffs) < tol:
            break
        coeffs = coeffs_new
        base = np.dot(vander, coeffs)
        y = np.minimum(y, base)
    return base


def baseline0(y, max_it=100, tol=1e-3):
    """deg = 0 in closed form (what the device kernel computes): returns (base[0], iterations)."""
    y = np.asarray(y, dtype=float)
    c, cmin, first = 1.0, np.inf, True
    for it in range(max_it):
        c_new = float(np.mean(np.minimum(y, cmin)))
        if abs(c_new - c) / abs(c) < tol:
            return (float(y[0]) if first else c), it
        c = c_new
        cmin = min(cmin, c)
        first = False
    return c, max_it
