"""numpy float64 restatement of the reference's objective hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows.  The arithmetic keeps the reference's operation
order so that, on the same inputs, results agree with the unmodified reference
to the last bit (checked by ``tests/golden/make_golden.py`` when it is run in
the build container, and by ``tests/test_oracle_golden.py`` against the
committed fixtures everywhere else).

Parity: PINNED against outputs of the unmodified reference (tests/golden/).
"""
import numpy as np
import scipy.integrate
import scipy.special

LN2 = np.log(2)


# --------------------------------------------------------------------------
# phase rotation  (reference: nmrfit/proc_autophase.py:9-36, ``ps2``)
# --------------------------------------------------------------------------
def phase_ramp(n, p0, p1):
    """phi_i = p0 + (p1 * i) / n, radians  (proc_autophase.py:30-31)."""
    return p0 + (p1 * np.arange(n) / n)


def ps2(u, v, p0=0.0, p1=0.0, inv=False):
    """Rotate (u, v) by the linear phase ramp; ``inv`` divides instead.

    proc_autophase.py:29-36.  The reference builds a complex vector, multiplies
    by exp(1j*phi) (or its reciprocal) and splits real/imag; the same complex
    numpy ops are used here so rounding is identical.
    """
    z = u + 1j * v
    rot = np.exp(1.0j * phase_ramp(z.shape[-1], p0, p1)).astype(z.dtype)
    if inv:
        rot = 1 / rot
    z = rot * z
    return z.real, z.imag


def ps(data, p0=0.0, p1=0.0, inv=False):
    """Phase rotation of a complex spectrum with (p0, p1) in DEGREES (proc_autophase.py:39-68)."""
    pi = np.pi
    p0 = p0 * pi / 180.
    p1 = p1 * pi / 180.
    size = data.shape[-1]
    apod = np.exp(1.0j * (p0 + (p1 * np.arange(size) / size))).astype(data.dtype)
    if inv:
        apod = 1 / apod
    return apod * data


# --------------------------------------------------------------------------
# phase estimation before the fit  
# --------------------------------------------------------------------------
def brute_phase(u, v, step=np.pi / 360):
    """Zero-order phase by exhaustive scan (containers.py:98-110): the candidate whose phased real part has the
    most level baseline (first-n mean vs last-n mean) among those that point upwards.  Returns (p0, 0.0)."""
    p0_best = 0
    best = np.inf
    n = max(1, int(len(u) / 5000))
    for p0 in np.arange(-np.pi, np.pi, step):
        V, _ = ps2(u, v, p0, 0.0)
        error = np.sqrt((V[:n].mean() - V[-n:].mean())**2)
        if error < best and np.max(V) > abs(np.min(V)):
            best = error
            p0_best = p0
    return p0_best, 0.0


def brute_phase_errors(u, v, step=np.pi / 360):
    """The scan's per-candidate (p0, error, upward) triples - what the device kernel is compared with."""
    n = max(1, int(len(u) / 5000))
    cands = np.arange(-np.pi, np.pi, step)
    err = np.empty(cands.size)
    ok = np.empty(cands.size, dtype=bool)
    for k, p0 in enumerate(cands):
        V, _ = ps2(u, v, p0, 0.0)
        err[k] = np.sqrt((V[:n].mean() - V[-n:].mean())**2)
        ok[k] = np.max(V) > abs(np.min(V))
    return cands, err, ok


def acme_score(ph, data):
    """ACME phase score, (p0, p1) in degrees (proc_autophase.py:142-187): entropy of the normalised first
    derivative of the phased real part + 1000 x the squared negative excursions."""
    stepsize = 1
    phc0, phc1 = ph
    s0 = ps(data, p0=phc0, p1=phc1)
    data = np.real(s0)
    ds1 = np.abs((data[1:] - data[:-1]) / (stepsize * 2))
    p1 = ds1 / np.sum(ds1)
    p1[p1 == 0] = 1
    h1 = -p1 * np.log(p1)
    h1s = np.sum(h1)
    pfun = 0.0
    as_ = data - np.abs(data)
    sumas = np.sum(as_)
    if sumas < 0:
        pfun = pfun + np.sum((as_ / 2) ** 2)
    p = 1000 * pfun
    return h1s + p


def approximate_phase(data, score=acme_score, p0=0.0, p1=0.0):
    """Nelder-Mead on the phase score, result converted degrees -> radians (proc_autophase.py:107-139)."""
    import scipy.optimize
    opt = scipy.optimize.fmin(score, x0=[p0, p1], args=(data, ), disp=False)
    return opt[0] * np.pi / 180, opt[1] * np.pi / 180


# --------------------------------------------------------------------------
# lineshape  (reference: nmrfit/equations.py:115-149, ``voigt``)
# --------------------------------------------------------------------------
def voigt(w, r, yoff, width, loc, a):
    """Area-parameterised pseudo-Voigt: yoff + a*(r*Lorentz + (1-r)*Gauss).

    equations.py:141 (Lorentzian), :144 (Gaussian), :147 (mix; ``r`` weighs the
    Lorentzian, ``yoff`` is added per peak).
    """
    lor = (2 / (np.pi * width)) * 1 / (1 + ((w - loc) / (0.5 * width))**2)
    gau = (2 / width) * np.sqrt(LN2 / np.pi) * np.exp(-((w - loc) / (width / (2 * np.sqrt(LN2))))**2)
    return yoff + a * (r * lor + (1 - r) * gau)


# --------------------------------------------------------------------------
# Kramers-Kronig counterpart  (reference: equations.py:9-80, 242)
# --------------------------------------------------------------------------
def _kk_integrand(x, r, yoff, width, loc, a, w):
    """[V(w - x) - V(w + x)] / x   (equations.py:37-49)."""
    plus = voigt(x + w, r, yoff, width, loc, a)
    minus = voigt(-x + w, r, yoff, width, loc, a)
    return 1 / x * (minus - plus)


def kk_quad(w, r, yoff, width, loc, a):
    """(1/pi) * integral_0^inf of the integrand, scipy ``quad`` defaults
    (equations.py:79-80).  Scalar ``w``.  ~6 ms per call."""
    val, _ = scipy.integrate.quad(_kk_integrand, 0, np.inf, args=(r, yoff, width, loc, a, w))
    return val / np.pi


kk_quad_vectorized = np.vectorize(kk_quad, otypes=[float])  # equations.py:242


def kk_closed(w, r, yoff, width, loc, a):
    """Closed form of the integral above (Hilbert transform of the Voigt body).

    Lorentzian part -> dispersion Lorentzian t/(1+t^2); Gaussian part ->
    (2/sqrt(pi)) * Dawson(s).  ``yoff`` cancels in the integrand.  Agreement
    with ``kk_quad`` is limited by quad's own tolerance (~1e-9 absolute); see
    tests/test_oracle_golden.py::test_kk_closed_matches_reference_quad.
    """
    del yoff
    t = (w - loc) / (0.5 * width)
    s = (w - loc) * (2 * np.sqrt(LN2)) / width
    lor = (2 / (np.pi * width)) * t / (1 + t * t)
    gau = (2 / width) * np.sqrt(LN2 / np.pi) * (2 / np.sqrt(np.pi)) * scipy.special.dawsn(s)
    return a * (r * lor + (1 - r) * gau)


# --------------------------------------------------------------------------
# objective  (reference: equations.py:152-212)
# --------------------------------------------------------------------------
def objective(x, w, u, v, weights, fit_im=False, kk=kk_closed):
    """Weighted RMSE between the phase-rotated data and the sum of peaks.

    equations.py:177 (unpack), :180 (ps2), :188-199 (peak loop; with
    ``fit_im is True`` the imaginary fit is OVERWRITTEN by each peak, so only the
    last peak's curve survives - reproduced), :202 (real RMSE), :205-209
    (imaginary RMSE added, then halved).

    ``kk`` picks the imaginary-lineshape evaluator: ``kk_closed`` (default,
    fast) or ``kk_quad_vectorized`` (what the reference literally runs).
    """
    p0, p1, r, yoff = x[:4]
    v_data, i_data = ps2(u, v, p0=p0, p1=p1)
    v_fit = np.zeros_like(v_data)
    if fit_im is True:
        i_fit = np.zeros_like(i_data)
    for k in range(4, len(x), 3):
        width, loc, a = x[k], x[k + 1], x[k + 2]
        v_fit = v_fit + voigt(w, r, yoff, width, loc, a)
        if fit_im is True:
            i_fit = kk(w, r, yoff, width, loc, a)
    rmse = np.sqrt(np.square(np.multiply(weights, (v_data - v_fit))).mean(axis=None))
    if fit_im is True:
        rmse += np.sqrt(np.square(np.multiply(weights, (i_data - i_fit))).mean(axis=None))
        rmse /= 2.0
    return rmse


def objective_swarm(xs, w, u, v, weights, fit_im=False):
    """One call per particle, as pyswarm drives it (utils.py:176-182)."""
    return np.array([objective(x, w, u, v, weights, fit_im) for x in xs])


# --------------------------------------------------------------------------
# weights  (reference: equations.py:215-238 ``laplace1d``;
#           utils.py:191-224 ``FitUtility._compute_weights``)
# --------------------------------------------------------------------------
def laplace1d(x, n=10, omega=0.33333333):
    """n Jacobi sweeps, endpoints pinned, in place (equations.py:236-238)."""
    for _ in range(n):
        x[1:-1] = (1. - omega) * x[1:-1] + omega * 0.5 * (x[2:] + x[:-2])
    return x


def compute_weights(w, peaks, expon=0.5):
    """Per-peak windows weighted by (tallest/|height|)**expon, then smoothed.

    utils.py:201-211 (index windows from ``peak.bounds``, swapped if reversed),
    :213-215 (tallest), :217-221 (fill; later peaks overwrite earlier), :223.
    """
    lo = np.zeros(len(peaks), dtype=int)
    hi = np.zeros(len(peaks), dtype=int)
    mag = np.zeros(len(peaks))
    for i, pk in enumerate(peaks):
        lo[i] = np.argmin(np.abs(w - pk.bounds[0]))
        hi[i] = np.argmin(np.abs(w - pk.bounds[1]))
        if lo[i] > hi[i]:
            lo[i], hi[i] = hi[i], lo[i]
        mag[i] = np.abs(pk.height)
    tallest = np.amax(mag)
    weights = np.ones(len(w)) * 1.0
    for i in range(len(peaks)):
        weights[lo[i]:hi[i] + 1] = np.power(tallest / mag[i], expon)
    return laplace1d(weights)


# --------------------------------------------------------------------------
# final curves  (reference: utils.py:226-295 ``FitUtility.generate_result``)
# --------------------------------------------------------------------------
def generate_result(params, w_data, scale=1, kk=kk_closed):
    """Per-peak real/imag contributions, their sums, and the inverse-phased
    (u, v) on the (optionally upsampled) grid.

    utils.py:236-241 (grid: ``scale == 1.0`` keeps ``w``; otherwise an ascending
    linspace of int(scale*N) points), :248-249, :262-277 (each real contribution
    includes yoff; imag contributions ACCUMULATE here), :284 (inverse ps2 whose
    ramp uses the upsampled length).  Returns a dict with the attributes the
    reference sets at :289-295.  The side effect on ``data`` (:252) is the
    caller's business.
    """
    if scale == 1.0:
        w = w_data
    else:
        w = np.linspace(w_data.min(), w_data.max(), int(scale * w_data.shape[0]))
    v_fit = np.zeros_like(w)
    i_fit = np.zeros_like(w)
    p0, p1, r, yoff = params[:4]
    rest = params[4:]
    real_contribs, imag_contribs = [], []
    for k in range(0, len(rest), 3):
        width, loc, a = rest[k], rest[k + 1], rest[k + 2]
        real = voigt(w, r, yoff, width, loc, a)
        imag = kk(w, r, yoff, width, loc, a)
        real_contribs.append(real)
        imag_contribs.append(imag)
        v_fit = v_fit + real
        i_fit = i_fit + imag
    u_fit, v_fit2 = ps2(v_fit, i_fit, inv=True, p0=p0, p1=p1)
    return dict(w=w, V=v_fit, I=i_fit, u=u_fit, v=v_fit2,
                real_contribs=real_contribs, imag_contribs=imag_contribs)


def get_areas(params):
    """utils.py:322: params[6::3]."""
    return np.array([params[i] for i in range(6, len(params), 3)])


def area_fraction(areas):
    """utils.py:302-310 / containers.py:246-252: sum(areas < mean) / sum(all)."""
    areas = np.asarray(areas)
    m = np.mean(areas)
    big = areas[areas >= m].sum()
    small = areas[areas < m].sum()
    return small / (big + small)


def solution_bounds(peaks, p0=None, p1=None):
    """containers.py:193-215.  ``p0``/``p1`` not None == force_p0/force_p1."""
    lower, upper = [], []
    for ph in (p0, p1):
        if ph is not None:
            upper.append(ph + 0.001)
            lower.append(ph - 0.001)
        else:
            upper.append(np.pi)
            lower.append(-np.pi)
    upper.extend([1.0, 0.01])
    lower.extend([0.0, -0.01])
    for pk in peaks:
        lower.extend([pk.width * 0.5, pk.loc - 0.1 * (pk.loc - pk.bounds[0]), pk.area * 0.5])
        upper.extend([pk.width * 1.5, pk.loc - 0.1 * (pk.loc - pk.bounds[1]), pk.area * 1.5])
    return lower, upper


# ---- auto peak selection (utils.py:670-783) ------------------------------------------------------------------
class PeakRecord:
    """The attributes AutoPeakSelector gives a ``Peak`` (utils.py:58-93, filled at :735-770)."""


def argrelmax_clip(x, order):
    """``scipy.signal.argrelmax(x, order=order)[0]`` (utils.py:733): indices whose value is STRICTLY greater than
    every neighbour within ``order`` samples on either side, neighbours beyond the ends clipped to the end samples
    (mode='clip').  scipy's own loop costs O(len(x) * order) - minutes for the 88,554-sample window of a 16,384-point
    spectrum upsampled 100x - so candidates come from a running maximum and only those are checked exhaustively."""
    import scipy.ndimage
    n = x.size
    order = max(int(order), 1)
    run = scipy.ndimage.maximum_filter1d(x, size=2 * min(order, n) + 1, mode='nearest')
    out = []
    for i in np.nonzero(x == run)[0]:
        lo, hi = max(0, i - order), min(n - 1, i + order)
        if i == 0 or i == n - 1:                           # the clipped neighbour is the sample itself
            continue
        if np.all(x[lo:i] < x[i]) and np.all(x[i + 1:hi + 1] < x[i]):
            out.append(i)
    return np.array(out, dtype=np.int64)


def auto_peaks(w, u, thresh, window, baseline_fn=None):
    """AutoPeakSelector(w, u, thresh, window).find_peaks() restated step by step (utils.py:709-772).  Returns
    (peaks, aux): ``peaks`` a list of PeakRecord (loc, i, height, width, bounds, baseline, area, idx_lo, idx_hi), ``aux``
    the upsampled axis / signal / smoothed signal / global baseline / maxima before the width screen."""
    import scipy.interpolate
    import scipy.signal
    if baseline_fn is None:
        from . import peakutils_oracle
        baseline_fn = peakutils_oracle.baseline
    w, u = np.asarray(w, dtype=float), np.asarray(u, dtype=float)
    f = scipy.interpolate.interp1d(w, u)                                      # utils.py:711
    wu = np.linspace(w.min(), w.max(), int(len(w) * 100))                    # :713
    uu = f(wu)                                                                # :714
    us = scipy.signal.savgol_filter(uu, 11, 4)                                # :716
    base = baseline_fn(us, 0)[0]                                              # :718
    order = int(window / (wu[1] - wu[0]))                                     # :728-729
    pre = []
    for i in argrelmax_clip(us, order):                                       # :731
        p = PeakRecord()
        p.loc, p.i, p.height = wu[i], int(i), uu[i] - base                    # :735-737
        if p.height > thresh:                                                 # :738
            pre.append(p)
    peaks = []
    y = uu - base
    for p in pre:                                                             # :747
        half = p.height / 2.
        d = np.sign(half - y[0:-1]) - np.sign(half - y[1:])                   # :748
        right = np.where(d < 0)[0]
        left = np.where(d > 0)[0]
        if right.size == 0 or left.size == 0:
            raise ValueError('attempt to get argmin of an empty sequence')   # what the reference raises here
        x_right = wu[right[np.argmin(np.abs(wu[right] - p.loc))]]             # :752
        x_left = wu[left[np.argmin(np.abs(wu[left] - p.loc))]]                # :753
        if x_left < x_right:                                                  # :755
            p.width = x_right - x_left
            p.bounds = [p.loc - 2 * p.width, p.loc + 2 * p.width]             # :760
            idx = np.where((wu >= p.bounds[0]) & (wu <= p.bounds[1]))[0]      # :763
            p.idx_lo, p.idx_hi = int(idx[0]), int(idx[-1])
            p.baseline = baseline_fn(uu[idx], 0)[0]                           # :766
            p.pre_height = p.height
            p.height = uu[p.i] - p.baseline                                   # :767
            p.area = scipy.integrate.simpson(uu[idx] - p.baseline, x=wu[idx])  # :770 (simps renamed in scipy 1.14)
            peaks.append(p)
    return peaks, dict(wu=wu, uu=uu, us=us, baseline=base, pre=pre)
