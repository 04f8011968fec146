"""The UNMODIFIED reference package (pip-installed into the git-ignored baseline/_ref by tools/install_reference.sh),
with nmrfit_b200.pyswarm_compat bound under the name `pyswarm`: the reference's own nmrfit.fit / FitUtility.fit runs
its swarm on the GPU and reproduces the golden fits (made by the same reference code driving the restated CPU pso)."""
import contextlib
import io
import os
import sys
import types

import numpy as np
import pytest

from conftest import load_golden, peaks_from_golden, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


@pytest.fixture(scope='module')
def reference():
    if not os.path.isdir(os.path.join(REF, 'nmrfit')):
        pytest.skip('baseline/_ref/nmrfit not installed (tools/install_reference.sh; needs /root/reference)')
    import nmrfit_b200.pyswarm_compat as compat
    np.float = float                                      # removed from numpy >= 1.24; used by the reference
    np.int = int
    saved = {k: sys.modules.get(k) for k in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec',
                                             'matplotlib.widgets', 'peakutils', 'nmrglue', 'pyswarm', 'nmrfit')}
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec', 'matplotlib.widgets', 'peakutils', 'nmrglue'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['pyswarm'] = compat                       # the whole integration: one module binding
    sys.modules.pop('nmrfit', None)
    sys.path.insert(0, REF)
    try:
        import nmrfit
        assert os.path.realpath(nmrfit.__file__).startswith(os.path.realpath(REF))
        yield nmrfit
    finally:
        sys.path.remove(REF)
        for k in [m for m in sys.modules if m == 'nmrfit' or m.startswith('nmrfit.')]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _ref_data(ref, g):
    d = ref.containers.Data(g['w'].copy(), g['u'].copy(), g['v'].copy())
    peaks = ref.utils.Peaks()
    for p in peaks_from_golden(g):
        q = ref.utils.Peak()
        q.loc, q.width, q.area, q.height, q.bounds = p.loc, p.width, p.area, p.height, list(p.bounds)
        peaks.append(q)
    d.peaks = peaks
    return d


@pytest.mark.parametrize('case', ['fit_lite_1024x6', 'fit_c1_4096x6', 'fit_default_2048x6'])
def test_unmodified_reference_fit_runs_on_the_gpu(reference, case):
    g = load_golden(case)
    rd = _ref_data(reference, g)
    np.random.seed(int(g['seed']))
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        fobj = reference.fit(rd, list(g['lower']), list(g['upper']),
                             options={'swarmsize': int(g['swarmsize']), 'maxiter': int(g['maxiter'])})
    assert type(fobj).__module__ == 'nmrfit.utils'       # the reference's FitUtility, not ours
    assert np.array_equal(fobj.weights, g['weights'])
    assert relerr(fobj.params, g['params']) < 1e-6 and abs(fobj.error / g['error'] - 1) < 1e-9
    assert np.random.rand() == g['next_rand']            # the legacy stream is where pyswarm would have left it
    assert 'Stopping search:' in sink.getvalue() and 'Fit Summary:' in sink.getvalue()
    assert abs(fobj.calculate_area_fraction() / g['area_fraction'] - 1) < 1e-6
