"""Fit driver, mirroring ``nmrfit.utils`` for the hot path.

``FitUtility`` keeps the reference's constructor, methods and result attributes
(utils.py:96-339); what changes is underneath: the swarm lives on the GPU and a
generation is one batched launch instead of ``swarmsize`` Python callbacks.

``Peak`` / ``Peaks`` are the plain records the driver reads (utils.py:14-93).
``AutoPeakSelector`` (utils.py:670-783) and its batched form ``select_peaks_batch`` run on the GPU
(csrc/peaks.cu).  The interactive, matplotlib-driven selectors of the reference (BoundsSelector,
PeakSelector: utils.py:342-667) are outside the accelerated path and are not provided.
"""
import numpy as np

from . import _cabi
from . import equations
from . import swarm as _swarm


class Peaks(list):
    """List of ``Peak`` records (utils.py:14-55)."""

    def average_height(self):
        return sum(abs(p.height) for p in self) / len(self)

    def split(self):
        """(main peaks, satellites) by height relative to the average."""
        h = self.average_height()
        mains, sats = Peaks(), Peaks()
        for p in self:
            (mains if abs(p.height) >= h else sats).append(p)
        return mains, sats


class Peak:
    """Metadata of one peak: ``loc``, ``height``, ``bounds`` [lo, hi], ``width`` (FWHM), ``area``."""

    def __repr__(self):
        return 'Peak(loc=%s, height=%s, bounds=[%s, %s], width=%s, area=%s)' % (
            getattr(self, 'loc', None), getattr(self, 'height', None),
            *(getattr(self, 'bounds', [None, None])), getattr(self, 'width', None), getattr(self, 'area', None))


def compute_weights(w, peaks, expon=0.5):
    """Frequency-dependent residual weights (utils.py:191-224): 1 everywhere,
    (tallest/|height|)**expon inside each peak's index window (later peaks
    overwrite earlier ones), then 10 Jacobi smoothing sweeps.  Host-side: it runs
    once per fit and costs microseconds."""
    n_pk = len(peaks)
    lo = np.zeros(n_pk, dtype=int)
    hi = np.zeros(n_pk, dtype=int)
    mag = np.zeros(n_pk)
    for i, pk in enumerate(peaks):
        a = int(np.argmin(np.abs(w - pk.bounds[0])))
        b = int(np.argmin(np.abs(w - pk.bounds[1])))
        lo[i], hi[i] = min(a, b), max(a, b)
        mag[i] = np.abs(pk.height)
    tallest = np.amax(mag)
    weights = np.ones(len(w))
    for i in range(n_pk):
        weights[lo[i]:hi[i] + 1] = np.power(tallest / mag[i], expon)
    return equations.laplace1d(weights)


def peak_windows(peaks_list, expon=0.5):
    """Inputs of the device weights kernel for a batch: ``bounds`` [B, K, 2] (``Peak.bounds``) and ``values``
    [B, K] = (tallest/|height|)**expon per spectrum (utils.py:206-221) - the only host arithmetic left."""
    B, K = len(peaks_list), len(peaks_list[0])
    bounds = np.empty((B, K, 2))
    mag = np.empty((B, K))
    for b, peaks in enumerate(peaks_list):
        if len(peaks) != K:
            raise ValueError('every spectrum of a batch must have the same number of peaks')
        for k, pk in enumerate(peaks):
            bounds[b, k, 0], bounds[b, k, 1] = pk.bounds[0], pk.bounds[1]
            mag[b, k] = np.abs(pk.height)
    values = np.power(np.amax(mag, axis=1, keepdims=True) / mag, expon)
    return bounds, values


class FitUtility:
    """Interface used to perform a fit of the data (drop-in for utils.py:96-339).

    Attributes after ``fit()``: ``params`` (ndarray D), ``error`` (float),
    ``weights``; after ``generate_result()``: ``w, u, v, V, I`` and the per-peak
    lists ``real_contribs`` / ``imag_contribs``.

    ``options`` understands the reference's keys (``swarmsize`` 204, ``maxiter``
    2000, ``omega`` -0.2134, ``phip`` -0.3344, ``phig`` 2.3259; utils.py:177-181)
    plus optional extras that do not exist upstream:
      ``rng``       'host' (default): numpy's legacy global stream in pyswarm's order, so
                    ``np.random.seed(k)`` gives the reference's trajectory - the stream is continued ON THE
                    DEVICE from np.random's own state (bit-identical numbers, state handed back);
                    'host_arrays': the same numbers drawn on the host and copied; 'device': Philox on the GPU.
      ``seed``      Philox seed for rng='device'.
      ``minstep``, ``minfunc``   pyswarm's stop tolerances (1e-8).
      ``precision`` 'fp64' (default) or 'fp32'.
      ``chunk``     generations queued between host checks of the stop flag.
      ``fused``     'auto' (default) | 'off' | 'require': run the generations of a small swarm in one
                    cooperative launch (bit-identical to the per-step kernels).
      ``device``    CUDA device index.
    ``processes`` is accepted and ignored (the GPU evaluates every particle at once).
    """

    def __init__(self, data, lower, upper, expon=0.5, dynamic_weighting=True, fit_im=False, processes=1,
                 summary=True, options={}):
        self.data = data
        self.lower = lower
        self.upper = upper
        self.expon = expon
        self.dynamic_weighting = dynamic_weighting
        self.fit_im = fit_im
        self.summary = summary
        self.processes = processes
        self.options = options

    def fit(self):
        """Minimise the objective over the box [lower, upper] with the swarm."""
        self.weights = self._compute_weights()
        if self.dynamic_weighting is False:
            self.weights = np.ones_like(self.weights)

        opt = self.options
        xopt, fopt, info = _swarm.pso_single(
            self.data.w, self.data.u, self.data.v, self.weights, self.lower, self.upper,
            fit_im=self.fit_im,
            swarmsize=opt.get('swarmsize', 204), maxiter=opt.get('maxiter', 2000),
            omega=opt.get('omega', -0.2134), phip=opt.get('phip', -0.3344), phig=opt.get('phig', 2.3259),
            minstep=opt.get('minstep', 1e-8), minfunc=opt.get('minfunc', 1e-8),
            rng=opt.get('rng', 'host'), seed=opt.get('seed', 0), precision=opt.get('precision', 'fp64'),
            chunk=opt.get('chunk', 16), device=opt.get('device', None), fused=opt.get('fused', 'auto'))

        self.params = xopt
        self.error = fopt
        self.fit_info = info

        if self.summary is True:
            self._print_summary()

    def _compute_weights(self):
        return compute_weights(self.data.w, self.data.peaks, self.expon)

    def generate_result(self, scale=1):
        """Evaluate the fitted curves, optionally upsampled by ``scale`` (utils.py:226-295)."""
        if scale == 1.0:
            w = self.data.w
        else:
            w = np.linspace(self.data.w.min(), self.data.w.max(), int(scale * self.data.w.shape[0]))

        p0, p1 = self.params[0], self.params[1]
        # the reference re-phases the data object here (utils.py:252)
        self.data.shift_phase(method='manual', p0=p0, p1=p1)

        params = _cabi.as_f64(self.params)
        n_peaks = (params.size - 4) // 3
        w = _cabi.as_f64(w)
        n = w.size
        # one result block (pinned and pooled when large: _cabi.result_empty), carved into the six outputs
        block = _cabi.result_empty((2 * n_peaks + 4) * n)
        real = block[:n_peaks * n].reshape(n_peaks, n)
        imag = block[n_peaks * n:2 * n_peaks * n].reshape(n_peaks, n)
        V, I, u, v = (block[(2 * n_peaks + k) * n:(2 * n_peaks + k + 1) * n] for k in range(4))
        dev = self.options.get('device', None)
        _cabi.check(_cabi.lib().nmrfit_generate_result_host(
            _cabi.default_device() if dev is None else int(dev), _cabi.ptr(params), n_peaks, _cabi.ptr(w), n,
            _cabi.ptr(real), _cabi.ptr(imag), _cabi.ptr(V), _cabi.ptr(I), _cabi.ptr(u), _cabi.ptr(v)))

        self.u = u
        self.v = v
        self.V = V
        self.I = I
        self.w = w
        self.real_contribs = [real[k] for k in range(n_peaks)]
        self.imag_contribs = [imag[k] for k in range(n_peaks)]

    def calculate_area_fraction(self):
        """Satellite area / total area, satellites = areas below the mean (utils.py:297-310)."""
        areas = self.get_areas()
        m = np.mean(areas)
        mains = areas[areas >= m].sum()
        sats = areas[areas < m].sum()
        return sats / (mains + sats)

    def get_areas(self):
        """Fitted areas: every third parameter from index 6 (utils.py:322)."""
        return np.array([self.params[i] for i in range(6, len(self.params), 3)])

    def _print_summary(self):
        import pandas as pd
        res = np.array(self.params)
        res_globals = pd.DataFrame(res[:4].reshape((1, -1)), columns=['p0', 'p1', 'r', 'y-off'])
        res_peaks = pd.DataFrame(res[4:].reshape((-1, 3)), columns=['width', 'location', 'area'])
        print('\nFit Summary:')
        print('------------')
        print('Global parameters')
        print(res_globals.to_string(index=False))
        print('\nPeak parameters')
        print(res_peaks.to_string(index=False))
        print("Error:\t", self.error)


# ---- automatic peak selection (utils.py:670-783) on the GPU ----------------------------------------------------
_SG_TABLES = None


def _sg_tables():
    """Savitzky-Golay(11, 4) coefficients and edge maps handed to the device: scipy's own when scipy is importable
    (what the reference would compute: utils.py:716), else the library's built-in copy (None)."""
    global _SG_TABLES
    if _SG_TABLES is None:
        try:
            import scipy.signal
            c = scipy.signal.savgol_coeffs(11, 4)
            pinv = np.linalg.pinv(np.vander(np.arange(11.0), 5))
            left = np.vander(np.arange(0, 5.0), 5) @ pinv
            right = np.vander(np.arange(6, 11.0), 5) @ pinv
            _SG_TABLES = (np.concatenate([c, left.ravel(), right.ravel()]),)
        except ImportError:
            _SG_TABLES = (None,)
    return _SG_TABLES[0]


def _ascending(w, u):
    """interp1d sorts its abscissae (stable argsort) before interpolating (utils.py:711): mirror it on the host."""
    w, u = _cabi.as_f64(w), _cabi.as_f64(u)
    if w.size > 1 and np.all(w[1:] > w[:-1]):
        return w, u
    ind = np.argsort(w, kind='mergesort')
    return np.ascontiguousarray(w[ind]), np.ascontiguousarray(u[ind])


def select_peaks_batch(ws, us, thresh=0.0, window=0.02, upsample=100, max_peaks=64, device=None, details=False):
    """``AutoPeakSelector(w, u, thresh, window).find_peaks()`` for every spectrum of a batch in one pass on the GPU.

    ``ws``, ``us``: [B, N] (or [N]) axis and phased real part of each spectrum.  Returns a list of ``Peaks`` (one per
    spectrum; ``Peak`` records with ``loc, i, height, width, bounds, baseline, area, idx``), or with ``details`` also the
    global baselines and the maxima before the width screen.  Raises ValueError where the reference does (a maximum
    without a half-height crossing of either kind: utils.py:752-753 take the argmin of an empty sequence)."""
    ws, us = np.atleast_2d(np.asarray(ws, dtype=np.float64)), np.atleast_2d(np.asarray(us, dtype=np.float64))
    if ws.shape != us.shape:
        raise ValueError('ws and us must have the same shape')
    B, N = ws.shape
    rows = [_ascending(ws[b], us[b]) for b in range(B)]
    W = np.ascontiguousarray(np.stack([r[0] for r in rows]))
    U = np.ascontiguousarray(np.stack([r[1] for r in rows]))
    with _cabi.PeakPicker(B, N, upsample=upsample, max_peaks=max_peaks, device=device) as pk:
        n_max, idx, val, base = pk.maxima(W, U, window, sg_coeffs=_sg_tables())
        n_keep = np.zeros(B, dtype=np.int32)
        keep_i = np.zeros((B, max_peaks), dtype=np.int64)
        keep_h = np.zeros((B, max_peaks))
        pre = []
        for b in range(B):
            order = np.argsort(idx[b, :n_max[b]])            # argrelmax returns ascending indices (utils.py:731)
            i_b, h_b = idx[b, :n_max[b]][order], val[b, :n_max[b]][order] - base[b]      # :737
            sel = h_b > thresh                              # :738
            n_keep[b] = int(sel.sum())
            keep_i[b, :n_keep[b]], keep_h[b, :n_keep[b]] = i_b[sel], h_b[sel]
            pre.append((i_b[sel], h_b[sel]))
        out = pk.measure(n_keep, keep_i, keep_h)
    result = []
    for b in range(B):
        peaks = Peaks()
        for k in range(n_keep[b]):
            if out['ok'][b, k] < 0:
                raise ValueError('attempt to get argmin of an empty sequence')       # as the reference does here
            if out['ok'][b, k] == 0:
                continue                                    # utils.py:755: x_left >= x_right, the peak is dropped
            p = Peak()
            p.loc, p.i = float(out['loc'][b, k]), int(keep_i[b, k])
            p.width = float(out['width'][b, k])
            p.bounds = [float(out['bounds'][b, k, 0]), float(out['bounds'][b, k, 1])]
            lo, hi = (int(t) for t in out['idx_range'][b, k])
            p.idx = (np.arange(lo, hi + 1),)                # what np.where returns (utils.py:763)
            p.baseline = float(out['baseline'][b, k])
            p.height = float(out['height'][b, k])
            p.area = float(out['area'][b, k])
            peaks.append(p)
        result.append(peaks)
    if details:
        return result, dict(baseline=base, pre=pre)
    return result


class AutoPeakSelector:
    """Drop-in for the reference's AutoPeakSelector (utils.py:670-783) with the arithmetic on the GPU: same constructor,
    ``find_maxima`` / ``find_width`` / ``find_peaks``, ``peaks`` and ``baseline`` attributes.  The upsampled arrays
    (``w``, ``u``, ``u_smoothed``: 100x the input, utils.py:713-716) stay on the device; ``plot`` is not provided."""

    def __init__(self, w, u, thresh, window):
        self.thresh = thresh
        self.window = window
        self._w, self._u = np.asarray(w, dtype=np.float64), np.asarray(u, dtype=np.float64)
        self.peaks = Peaks()
        self.baseline = None
        self._found = None

    def _run(self):
        if self._found is None:
            peaks, aux = select_peaks_batch(self._w[None], self._u[None], self.thresh, self.window, details=True)
            self._found = peaks[0]
            self.baseline = float(aux['baseline'][0])
            self._pre = aux['pre'][0]

    def find_maxima(self):
        self._run()
        w_lo, w_hi, M = self._w.min(), self._w.max(), self._w.size * 100
        step = (w_hi - w_lo) / (M - 1)
        self.peaks = Peaks()
        for i, h in zip(*self._pre):
            p = Peak()
            p.i, p.height = int(i), float(h)
            p.loc = float(w_hi if i == M - 1 else i * step + w_lo)     # np.linspace's own expression (utils.py:713)
            self.peaks.append(p)

    def find_width(self):
        self._run()
        self.peaks = self._found

    def find_peaks(self):
        self.find_maxima()
        self.find_width()
