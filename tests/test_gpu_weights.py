"""Batched residual weights on the device (csrc/weights.cu) against the reference's own `_compute_weights`
outputs (golden fixtures) and the oracle - bit for bit (integer windows + separately rounded FP64 sweeps)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden, peaks_from_golden, PeakRec
import nmrfit_b200
from nmrfit_b200 import _cabi, synth, utils
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
OBJ_CASES = ['c1_4096x6', 'ragged_1000x6', 'tiny_257x6', 'p12_2048', 'p24_1536']


def _device_weights(ws, peaks_list, expon=0.5, sweeps=10, omega=0.33333333):
    ws = np.asarray(ws)
    B, N = ws.shape
    with _cabi.Context(B, N, len(peaks_list[0])) as ctx:
        z = np.zeros_like(ws)
        ctx.set_spectra(ws, z, z)
        bounds, values = utils.peak_windows(peaks_list, expon)
        return ctx.compute_weights(bounds, values, sweeps=sweeps, omega=omega)


@pytest.mark.parametrize('case', OBJ_CASES)
def test_weights_match_reference_golden(case):
    g = load_golden('objective_' + case)
    got = _device_weights(g['w'][None], [peaks_from_golden(g)])
    assert np.array_equal(got[0], g['weights'])          # the unmodified reference's output


def _peak(loc, width, height, reverse=False):
    p = PeakRec()
    p.loc, p.width, p.height = loc, width, height
    p.bounds = [loc - 2 * width, loc + 2 * width][::-1] if reverse else [loc - 2 * width, loc + 2 * width]
    return p


def test_weights_edge_cases_match_oracle():
    """Overlapping windows (later peak wins), reversed bounds, a descending axis, windows hanging off both ends,
    negative heights, another exponent - several spectra in one launch."""
    N = 777
    asc = np.linspace(3.0, 4.0, N)
    axes = [asc, asc[::-1].copy(), asc + 1e-3 * np.sin(np.arange(N))]      # ascending, descending, non-uniform
    rng = np.random.default_rng(3)
    peaks_list = []
    for b in range(3):
        pk = [_peak(3.5, 0.05, 1.0), _peak(3.55, 0.05, -0.2, reverse=True), _peak(2.99, 0.02, 0.05),
              _peak(4.02, 0.03, 0.5), _peak(3.2 + 0.1 * rng.random(), 0.004, 0.013)]
        peaks_list.append(pk)
    for expon in (0.5, 0.25, 1.0):
        got = _device_weights(axes, peaks_list, expon=expon)
        for b in range(3):
            want = orc.compute_weights(axes[b], peaks_list[b], expon)
            assert np.array_equal(got[b], want), (expon, b)
    # odd / zero sweep counts end in the right buffer
    for sweeps in (0, 1, 3):
        got = _device_weights(axes[:1], peaks_list[:1], sweeps=sweeps)
        lo_hi = orc.compute_weights(axes[0], peaks_list[0])      # 10 sweeps; rebuild with the requested count below
        wts = np.ones(N)
        mags = np.array([abs(p.height) for p in peaks_list[0]])
        for p, m in zip(peaks_list[0], mags):
            a = int(np.argmin(np.abs(axes[0] - p.bounds[0]))); c = int(np.argmin(np.abs(axes[0] - p.bounds[1])))
            wts[min(a, c):max(a, c) + 1] = np.power(mags.max() / m, 0.5)
        want = orc.laplace1d(wts, n=sweeps)
        assert np.array_equal(got[0], want), sweeps
        assert lo_hi.shape == want.shape


def test_fit_batch_uses_device_weights_and_matches_single_fits():
    B, N, P = 6, 2048, 6
    datas, los, ups = [], [], []
    for b in range(B):
        d, _ = synth.multiplet(N, P, seed=300 + b)
        lo, up = d.generate_solution_bounds()
        datas.append(d); los.append(lo); ups.append(up)
    opts = {'swarmsize': 40, 'maxiter': 30, 'rng': 'host', 'seeds': list(range(50, 50 + B))}
    with contextlib.redirect_stdout(io.StringIO()):
        fits = nmrfit_b200.fit_batch(datas, los, ups, options=opts)
    for b, f in enumerate(fits):
        assert np.array_equal(f.weights, utils.compute_weights(datas[b].w, datas[b].peaks))
        np.random.seed(50 + b)
        with contextlib.redirect_stdout(io.StringIO()):
            one = nmrfit_b200.fit(datas[b], los[b], ups[b], summary=False, options={'swarmsize': 40, 'maxiter': 30})
        assert np.array_equal(f.params, one.params) and f.error == one.error
    # dynamic_weighting=False: ones, as in the reference (utils.py:171-173)
    with contextlib.redirect_stdout(io.StringIO()):
        flat = nmrfit_b200.fit_batch(datas[:2], los[:2], ups[:2], dynamic_weighting=False, options=opts | {'seeds': [1, 2]})
    assert all(np.array_equal(f.weights, np.ones(N)) for f in flat)
