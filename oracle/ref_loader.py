"""Import the UNMODIFIED reference package (pnnl/nmrfit) for the CPU arms of the benchmarks and for tests.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): nothing under ``nmrfit_b200/`` imports this.

Where the reference comes from, in order: ``baseline/_ref`` (the pip-installed copy made by
``tools/install_reference.sh``; git-ignored, travels to the GPU box with gpurun), then ``/root/reference`` (build
container only).  The reference's files are never edited; what SURVEY.md Appendix A found necessary to import
them under numpy >= 1.24 / scipy >= 1.14 without their optional third-party packages is applied around them:

  * ``np.float`` / ``np.int`` aliases (used at equations.py:242 and utils.py:201-202);
  * empty stub modules for matplotlib, peakutils, nmrglue (imported at module level, never called on this path);
  * ``pyswarm`` bound to a caller-supplied module-like object (default: the restated ``oracle.pso_oracle``, PARITY
    UNPINNED for that piece) - pyswarm itself is not installable here;
  * optionally ``scipy.integrate.simps`` -> ``scipy.integrate.simpson`` and a restated ``peakutils.baseline`` for the
    auto peak selector (``with_selector=True``; see ``oracle/peakutils_oracle.py``).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference')


def reference_root():
    """Directory that holds the reference's ``nmrfit`` package, or None."""
    for base in CANDIDATES:
        if os.path.isfile(os.path.join(base, 'nmrfit', 'equations.py')):
            return base
    return None


def _pyswarm_module(pso_fn=None):
    """A module object named ``pyswarm`` whose ``pso`` has pyswarm's signature and returns (xopt, fopt)."""
    from . import pso_oracle
    mod = types.ModuleType('pyswarm')

    def pso(func, lb, ub, ieqcons=[], f_ieqcons=None, args=(), kwargs={}, swarmsize=100, omega=0.5, phip=0.5,
            phig=0.5, maxiter=100, minstep=1e-8, minfunc=1e-8, debug=False, processes=1, particle_output=False):
        x, f, info = pso_oracle.pso(func, lb, ub, args=args, swarmsize=swarmsize, omega=omega, phip=phip, phig=phig,
                                    maxiter=maxiter, minstep=minstep, minfunc=minfunc, quiet=True)
        mod.last_info = info
        return x, f
    mod.pso = pso_fn or pso
    mod.last_info = None
    return mod


def load_reference(pyswarm=None, with_selector=False):
    """Import and return the reference's ``nmrfit`` package (whole: core, utils, containers, equations,
    proc_autophase).  Raises ImportError when neither location holds it."""
    base = reference_root()
    if base is None:
        raise ImportError('the reference package is neither under baseline/_ref (tools/install_reference.sh) nor at '
                          '/root/reference')
    mod = sys.modules.get('nmrfit')
    if mod is not None and os.path.realpath(getattr(mod, '__file__', '')).startswith(os.path.realpath(base)):
        return mod
    np.float = float
    np.int = int
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec', 'matplotlib.widgets', 'peakutils',
                 'nmrglue'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].gridspec = sys.modules['matplotlib.gridspec']
    sys.modules['matplotlib'].widgets = sys.modules['matplotlib.widgets']
    # names the reference binds at import time (`from matplotlib.widgets import SpanSelector`)
    for attr in ('SpanSelector', 'Slider', 'Button', 'RadioButtons'):
        setattr(sys.modules['matplotlib.widgets'], attr, getattr(sys.modules['matplotlib.widgets'], attr, object))
    sys.modules['pyswarm'] = pyswarm if pyswarm is not None else _pyswarm_module()
    if with_selector:
        import scipy.integrate
        from . import peakutils_oracle
        if not hasattr(scipy.integrate, 'simps'):
            scipy.integrate.simps = scipy.integrate.simpson      # renamed in scipy 1.14 (utils.py:591,632,770)
        sys.modules['peakutils'].baseline = peakutils_oracle.baseline
    sys.modules.pop('nmrfit', None)
    sys.path.insert(0, base)
    try:
        import nmrfit
    finally:
        sys.path.remove(base)
    nmrfit.__reference_root__ = base
    return nmrfit


def reference_data(ref, data):
    """A reference ``Data`` (with reference ``Peak`` records) carrying the numbers of an ``nmrfit_b200`` Data."""
    d = ref.containers.Data(np.array(data.w, dtype=float), np.array(data.u, dtype=float), np.array(data.v, dtype=float))
    d.p0, d.p1 = getattr(data, 'p0', 0.0), getattr(data, 'p1', 0.0)
    peaks = ref.utils.Peaks()
    for p in data.peaks:
        q = ref.utils.Peak()
        q.loc, q.width, q.area, q.height, q.bounds = p.loc, p.width, p.area, p.height, list(p.bounds)
        peaks.append(q)
    d.peaks = peaks
    return d
