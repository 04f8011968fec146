"""``Data`` container, mirroring the attribute contract of ``nmrfit.containers.Data``
that the fit path consumes (containers.py:8-252): ``w, u, v, V, I, p0, p1, peaks,
roibounds``; ``shift_phase(method='manual', p0=, p1=)``; ``select_bounds(low, high)``;
``generate_solution_bounds``; ``approximate_areas``; ``approximate_area_fraction``.

Phase estimation (``shift_phase(method='auto'|'brute')``, containers.py:71-74, 98-110) runs on
the GPU (csrc/phase.cu), and so does automatic peak picking (``select_peaks('auto')``,
containers.py:159-161 -> csrc/peaks.cu).  The interactive matplotlib selectors (containers.py:112-158)
are outside the accelerated path: they raise NotImplementedError and peaks can be attached with ``set_peaks``.
"""
import numpy as np

from . import proc_autophase


class Data:
    def __init__(self, w, u, v):
        self.w = w
        self.u = u
        self.v = v
        self.V = self.u[:]
        self.I = self.v[:]

    def shift_phase(self, method='auto', p0=0.0, p1=0.0, step=np.pi / 360, plot=False):
        """Phase shift u and v by (p0, p1) radians to generate V and I (GPU ``ps2``)."""
        if method.lower() == 'manual':
            self.p0 = p0
            self.p1 = p1
        elif method.lower() == 'auto':
            self.p0, self.p1 = proc_autophase.approximate_phase(self.u + 1j * self.v, 'acme')
        elif method.lower() == 'brute':
            self.p0, self.p1 = self._brute_phase(step=step)
        else:
            raise ValueError("Method must be 'auto', 'brute', or 'manual'.")
        self.V, self.I = proc_autophase.ps2(self.u, self.v, self.p0, self.p1)

    def _brute_phase(self, step=np.pi / 360):
        """Exhaustive zero-order scan (containers.py:98-110), every candidate in one launch.  Like the
        reference it leaves ``V, I`` phased by the LAST candidate until ``shift_phase`` recomputes them."""
        p0 = float(proc_autophase.brute_phase_batch(self.u[None], self.v[None], step)[0])
        return p0, 0.0

    def select_bounds(self, low=None, high=None):
        """Keep the points with low < w < high (strict, as utils.py:433)."""
        if low is None or high is None:
            raise NotImplementedError('interactive bound selection is not provided; pass low and high')
        idx = np.where((self.w > low) & (self.w < high))
        self.w, self.u, self.v = self.w[idx], self.u[idx], self.v[idx]

    def select_peaks(self, method='auto', n=None, one_click=False, thresh=0.0, window=0.02, plot=False):
        """Automatic peak selection on the GPU (containers.py:159-161 -> utils.AutoPeakSelector on ``self.V``); the
        interactive ``method='manual'`` selector of the reference is not provided."""
        if method.lower() == 'manual':
            raise NotImplementedError('interactive peak picking is not provided; attach Peak records with Data.set_peaks')
        if method.lower() != 'auto':
            raise ValueError("Method must be 'auto' or 'manual'.")
        from . import utils
        ps = utils.AutoPeakSelector(self.w, self.V, thresh=thresh, window=window)
        ps.find_peaks()
        self.peaks = ps.peaks
        self.roibounds = [p.bounds for p in self.peaks]

    def set_peaks(self, peaks):
        from .utils import Peaks
        self.peaks = peaks if isinstance(peaks, Peaks) else Peaks(peaks)
        self.roibounds = [p.bounds for p in self.peaks]

    def generate_solution_bounds(self, force_p0=False, force_p1=False):
        """Parameter box around the initial estimates (containers.py:175-217):
        phases +-pi (or +-1e-3 around the estimate when forced), r in [0,1], yoff in
        +-0.01, width and area x0.5..x1.5, centre within 10 % of the peak's bounds."""
        lower, upper = [], []
        for forced, ph in ((force_p0, getattr(self, 'p0', 0.0)), (force_p1, getattr(self, 'p1', 0.0))):
            if forced is True:
                upper.append(ph + 0.001)
                lower.append(ph - 0.001)
            else:
                upper.append(np.pi)
                lower.append(-np.pi)
        upper.extend([1.0, 0.01])
        lower.extend([0.0, -0.01])
        for p in self.peaks:
            lower.extend([p.width * 0.5, p.loc - 0.1 * (p.loc - p.bounds[0]), p.area * 0.5])
            upper.extend([p.width * 1.5, p.loc - 0.1 * (p.loc - p.bounds[1]), p.area * 1.5])
        return lower, upper

    def approximate_areas(self):
        return [p.area for p in self.peaks]

    def approximate_area_fraction(self):
        areas = np.array(self.approximate_areas())
        m = np.mean(areas)
        mains = areas[areas >= m].sum()
        sats = areas[areas < m].sum()
        return sats / (mains + sats)
