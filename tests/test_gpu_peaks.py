"""Automatic peak selection on the GPU (csrc/peaks.cu, utils.select_peaks_batch / AutoPeakSelector / Data.select_peaks)
against the golden fixtures of the reference's own AutoPeakSelector (utils.py:670-783; tests/golden/make_golden.py) and
the oracle restatement.  Integer results (maxima, index ranges) exactly; loc and bounds exactly (np.linspace's own
expression); heights, widths, baselines and Simpson areas to 1e-12 relative (summation order)."""
import numpy as np
import pytest

import nmrfit_b200
from conftest import load_golden, relerr
from nmrfit_b200 import _cabi, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
CASES = ['peaks_1024x6', 'peaks_2500x12', 'peaks_4096x6', 'peaks_desc_1500x6',
         'peaks_c3_16384x6']      # the last: a BASELINE configs[2] spectrum (1,638,400 upsampled samples, window 88,554)
TOL = 1e-12


def _check(peaks, g):
    assert [p.i for p in peaks] == list(g['i'])
    assert np.array_equal([p.loc for p in peaks], g['loc'])
    assert [int(p.idx[0][0]) for p in peaks] == list(g['idx_lo']) and [int(p.idx[0][-1]) for p in peaks] == list(g['idx_hi'])
    assert np.array_equal([p.width for p in peaks], g['width'])
    assert np.array_equal(np.array([p.bounds for p in peaks]), g['bounds'])
    for key, gk in (('height', 'height'), ('area', 'area'), ('baseline', 'local_baseline')):
        got = np.array([getattr(p, key) for p in peaks])
        assert np.max(np.abs(got - g[gk])) <= TOL * np.max(np.abs(g[gk])), key


@pytest.mark.parametrize('case', CASES)
def test_selector_matches_the_reference_class(case):
    g = load_golden(case)
    sel = utils.AutoPeakSelector(g['w'], g['V'], float(g['thresh']), float(g['window']))
    sel.find_maxima()
    assert [p.i for p in sel.peaks] == list(g['pre_i'])
    assert np.array_equal([p.loc for p in sel.peaks], g['pre_loc'])
    assert relerr([p.height for p in sel.peaks], g['pre_height']) < 1e-10
    assert abs(sel.baseline - g['baseline']) <= 1e-12 * max(abs(float(g['baseline'])), 1e-3)
    sel.find_width()
    _check(sel.peaks, g)


def test_upsampled_signal_and_smoothing_at_probe_points():
    g = load_golden('peaks_1024x6')
    with _cabi.PeakPicker(1, g['w'].size, max_peaks=64) as pk:
        pk.maxima(g['w'][None], g['V'][None], float(g['window']), sg_coeffs=utils._sg_tables())
        wu, uu, us = pk.probe(0, g['probe'])
    assert np.array_equal(wu, g['wu_probe'])                 # np.linspace's own expression
    assert np.array_equal(uu, g['uu_probe'])                 # interp1d's own expression
    interior = (g['probe'] >= 5) & (g['probe'] < g['w'].size * 100 - 5)
    assert np.array_equal(us[interior], g['us_probe'][interior])         # ndimage's symmetric correlation, term by term
    assert np.max(np.abs(us - g['us_probe'])) < 1e-12 * np.abs(g['us_probe']).max()   # edge fits: least squares to rounding


def test_batch_of_spectra_and_data_select_peaks():
    B, N = 5, 2048
    datas = [synth.multiplet(N, 6, seed=40 + b)[0] for b in range(B)]
    Vs = []
    for d in datas:
        d.shift_phase(method='manual', p0=d.p0, p1=d.p1)
        Vs.append(d.V)
    got = utils.select_peaks_batch(np.array([d.w for d in datas]), np.array(Vs), thresh=0.003, window=0.02)
    for b, d in enumerate(datas):
        want, _ = orc.auto_peaks(d.w, Vs[b], 0.003, 0.02)
        assert [p.i for p in got[b]] == [p.i for p in want] and len(want) == 6
        for key in ('loc', 'width', 'height', 'area', 'baseline'):
            assert relerr([getattr(p, key) for p in got[b]], [getattr(p, key) for p in want]) < 1e-11, key
    # the drop-in entry point: Data.select_peaks('auto') fills peaks and roibounds, and the bounds feed the fit
    d = datas[0]
    d.select_peaks(method='auto', thresh=0.003, window=0.02)
    assert len(d.peaks) == 6 and d.roibounds == [p.bounds for p in d.peaks]
    lo, up = d.generate_solution_bounds()
    assert len(lo) == 22 and all(u > l for l, u in zip(lo, up))
    with pytest.raises(NotImplementedError):
        d.select_peaks(method='manual', n=6)


def test_fit_batch_from_raw_spectra():
    """a batch fitted from raw (w, u, v) - phase known, peaks picked on the device - ends where a
    batch with hand-built Peak records of the same numbers ends."""
    import contextlib
    import io
    B, N = 3, 4096
    datas = []
    for b in range(B):
        d, true = synth.multiplet(N, 6, seed=60 + b)
        d.shift_phase(method='manual', p0=d.p0, p1=d.p1)
        datas.append(d)
    picked = utils.select_peaks_batch(np.array([d.w for d in datas]), np.array([d.V for d in datas]), thresh=0.003, window=0.02)
    for d, peaks in zip(datas, picked):
        assert len(peaks) == 6
        d.set_peaks(peaks)
    bounds = [d.generate_solution_bounds() for d in datas]
    with contextlib.redirect_stdout(io.StringIO()):
        fits = nmrfit_b200.fit_batch(datas, [b[0] for b in bounds], [b[1] for b in bounds],
                                     options={'swarmsize': 64, 'maxiter': 40, 'rng': 'device', 'seed': 1})
    for f in fits:
        assert np.isfinite(f.error) and f.error < 0.05 and 0.0 < f.calculate_area_fraction() < 0.1
