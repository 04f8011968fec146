"""Accuracy of the device math (nmrfit_b200/csrc/nmrfit_math.cuh) compiled for the host.
The header is the single source of the lineshape arithmetic; compiling it with g++
lets the polynomial tables, the range reduction and the reciprocal refinement be
checked against libm / mpmath without a GPU.  (The MUFU reciprocal seed is emulated
by truncating to the 20 mantissa bits of a high word.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.special import dawsn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope='module')
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('hm') / 'host_math.so')
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', so,
                    os.path.join(ROOT, 'tests', 'host_math_harness.cpp')], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize('tb', [0, 6, 8, 10])
def test_exp_neg(hm, tb):
    rng = np.random.default_rng(tb)
    x = -np.concatenate([rng.random(100000) * 40, 10.0 ** rng.uniform(-14, 1, 20000), [0.0, 1e-300]])
    out = np.empty_like(x)
    hm.h_exp_neg(tb, P(x), len(x), P(out))
    rel = np.abs(out / np.exp(x) - 1)
    assert rel.max() < 2e-15, rel.max()
    # deep tail: single-FMA reduction costs accuracy proportional to |x|, irrelevant in absolute terms
    x = -rng.uniform(40, 700, 50000)
    hm.h_exp_neg(tb, P(x), len(x), P(out[:len(x)]))
    assert np.abs(out[:len(x)] / np.exp(x) - 1).max() < 5e-14
    # below the clamp everything collapses to ~exp(-700): finite, tiny, never NaN/inf
    x = -np.array([700.0, 701.0, 1e4, 1e9, 1e300])
    hm.h_exp_neg(tb, P(x), len(x), P(out[:len(x)]))
    assert np.all(np.isfinite(out[:len(x)])) and np.all(out[:len(x)] < 1e-300) and np.all(out[:len(x)] > 0)


def test_rcp_pos(hm):
    rng = np.random.default_rng(1)
    q = 1 + 10.0 ** rng.uniform(-9, 12, 200000)
    out = np.empty_like(q)
    hm.h_rcp_pos(P(q), len(q), P(out))
    assert np.abs(out * q - 1).max() < 3e-16


def test_dawson(hm):
    rng = np.random.default_rng(2)
    s = np.concatenate([rng.uniform(-12, 12, 200000), 10.0 ** rng.uniform(-12, 5, 20000),
                        [0.0, 8.0, np.nextafter(8.0, 0), -8.0, 0.25, 0.5]])
    out = np.empty_like(s)
    hm.h_dawson(P(s), len(s), P(out))
    assert np.abs(out - dawsn(s)).max() < 4e-16
    assert np.array_equal(out[:1000], -_daw(hm, -s[:1000]))      # odd


def _daw(hm, s):
    s = np.ascontiguousarray(s)
    out = np.empty_like(s)
    hm.h_dawson(P(s), len(s), P(out))
    return out


@pytest.mark.parametrize('tb', [0, 6])
def test_voigt_body_matches_reference_formula(hm, tb):
    from oracle import nmrfit_oracle as orc
    w = np.linspace(3.23, 3.60, 5000)
    for r, width, loc, a in ((0.55, 0.004, 3.41, 0.012), (0.0, 0.002, 3.3, 1.5), (1.0, 0.006, 3.59, 0.2)):
        out = np.empty_like(w)
        hm.h_voigt_body.argtypes = [dp, ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.c_int, dp]
        hm.h_voigt_body(P(w), len(w), r, width, loc, a, tb, P(out))
        want = orc.voigt(w, r, 0.0, width, loc, a)
        assert np.abs(out - want).max() < 4e-16 * np.abs(want).max() + 1e-300
        assert np.abs(out / want - 1)[want > 1e-12 * want.max()].max() < 1e-13


def test_philox_known_answer_and_range(hm):
    # Random123 known-answer test for philox4x32-10: counter = key = 0
    out = np.empty(2)
    hm.h_philox.argtypes = [ctypes.c_ulonglong] * 3 + [dp]
    hm.h_philox(0, 0, 0, P(out))
    c = [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    want = [((c[0] >> 5) * 67108864.0 + (c[1] >> 6)) / 9007199254740992.0,
            ((c[2] >> 5) * 67108864.0 + (c[3] >> 6)) / 9007199254740992.0]
    assert list(out) == want
    vals = []
    for i in range(2000):
        hm.h_philox(12345, i, 7, P(out))
        vals.extend(out)
    vals = np.array(vals)
    assert vals.min() >= 0 and vals.max() < 1 and abs(vals.mean() - 0.5) < 0.02


def _span_fit(hm, w, h, x, n_peaks, R, tb):
    hm.h_span_fit.argtypes = [dp, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(len(w))
    hm.h_span_fit(P(w), len(w), h, P(x), n_peaks, R, tb, P(out))
    return out


@pytest.mark.parametrize('R', [4, 8, 16])
@pytest.mark.parametrize('shape', [(4096, 6, 1000), (1024, 6, 3), (301, 6, 4), (8192, 12, 2000)])
def test_uniform_axis_span_recurrence(hm, R, shape):
    """peak_span (two exponentials per peak and R points, then a multiplicative recurrence) against the
    reference formula on the uniform axes the synthetic spectra use.  The axis is treated as exactly
    uniform, which moves an abscissa by <= 1 ulp(w): <= 5e-13 of the curve's scale."""
    from oracle import nmrfit_oracle as orc
    from nmrfit_b200 import synth
    N, n_peaks, seed = shape
    data, true = synth.multiplet(N, n_peaks, seed=seed)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 5, seed=11)
    xs[0] = true
    w = data.w
    h = (w[-1] - w[0]) / (N - 1)
    for x in xs:
        want = sum(orc.voigt(w, x[2], x[3], *x[4 + 3 * k:7 + 3 * k]) for k in range(n_peaks))
        for tb in (6, 10):
            got = _span_fit(hm, w, h, x, n_peaks, R, tb)
            assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()


def test_uniform_axis_span_edge_cases(hm):
    """Descending axes, coarse grids (every point its own exponential), far tails and absurd widths:
    finite everywhere and equal to the reference formula."""
    from oracle import nmrfit_oracle as orc
    N = 512
    for w in (np.linspace(3.6, 3.23, N), np.linspace(-1.0, 1.0, N), np.linspace(3.23, 3.60, N)):
        h = (w[-1] - w[0]) / (N - 1)
        mid = 0.5 * (w[0] + w[-1])
        for width in (1e-6, 2e-4, 1e-3, 4e-3, 0.05, 10.0):
            for loc in (mid + 0.01, w[0] - 50 * abs(h), mid + 100.0):
                x = np.array([0.1, 0.2, 0.4, 0.003, width, loc, 0.7])
                want = orc.voigt(w, x[2], x[3], width, loc, x[6])
                for R in (4, 8, 16):
                    got = _span_fit(hm, w, h, x, 1, R, 6)
                    assert np.all(np.isfinite(got))
                    assert np.abs(got - want).max() < 2e-12 * max(np.abs(want).max(), 1e-3), (width, loc, R)


@pytest.mark.parametrize('R', [4, 8, 16])
@pytest.mark.parametrize('shape', [(4096, 6, 1000), (32768, 12, 2000), (300, 6, 4), (16384, 24, 4000)])
def test_far_field_split(hm, R, shape):
    """Near peaks by peak_span, far peaks through one degree-11 polynomial per 32*R-point region
    (far_accumulate / far_eval): same curve as the reference formula, and on fine grids most peaks are far."""
    from oracle import nmrfit_oracle as orc
    from nmrfit_b200 import synth
    N, n_peaks, seed = shape
    data, true = synth.multiplet(N, n_peaks, seed=seed)
    lo, up = data.generate_solution_bounds()
    xs = synth.particles(lo, up, 4, seed=5)
    xs[0] = true
    w = data.w
    h = (w[-1] - w[0]) / (N - 1)
    hm.h_region_fit.argtypes = [dp, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int, dp]
    n_regions = (N + 32 * R - 1) // (32 * R)
    for x in xs:
        x = np.ascontiguousarray(x)
        got = np.empty(N)
        near = hm.h_region_fit(P(w), N, h, P(x), n_peaks, R, P(got))
        want = sum(orc.voigt(w, x[2], x[3], *x[4 + 3 * k:7 + 3 * k]) for k in range(n_peaks))
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
        if N >= 16384 and R <= 8:
            assert near / n_regions < 0.25 * n_peaks


def test_far_field_extremes(hm):
    """Broad peaks (far because smooth), peaks outside the window, mixed widths, descending axis."""
    from oracle import nmrfit_oracle as orc
    hm.h_region_fit.argtypes = [dp, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int, dp]
    N = 8192
    for w in (np.linspace(3.23, 3.60, N), np.linspace(3.60, 3.23, N), np.linspace(-200.0, 200.0, N)):
        h = (w[-1] - w[0]) / (N - 1)
        span = abs(w[-1] - w[0])
        mid = 0.5 * (w[0] + w[-1])
        x = np.array([0.3, -0.2, 0.45, 1e-3,
                      0.011 * span, mid + 0.1 * span, 0.5,        # ordinary line
                      2.0 * span, mid - 0.2 * span, 3.0,          # broader than the window
                      0.004 * span, mid + 30 * span, 1.0,         # far outside
                      1e-7 * span, mid, 1e-4,                     # narrower than the grid: exact path
                      0.02 * span, w[0], 0.2, 0.02 * span, w[-1], 0.2])   # on the edges
        n_peaks = (len(x) - 4) // 3
        want = sum(orc.voigt(w, x[2], x[3], *x[4 + 3 * k:7 + 3 * k]) for k in range(n_peaks))
        for R in (4, 8, 16):
            got = np.empty(N)
            hm.h_region_fit(P(w), N, h, P(x), n_peaks, R, P(got))
            assert np.all(np.isfinite(got))
            assert np.abs(got - want).max() < 2e-12 * np.abs(want).max(), R


def test_far_polynomial_economisation(hm):
    """far_economise: the 12-term series of a far-field cell folded into 10 coefficients (Chebyshev, the T11 and T10 parts
    dropped) and evaluated as even + odd parts.  With coefficients decaying like the series' own (|C_n| <= |u|^n,
    |u| <= 1/16) the result stays within 2e-15 of C_0's scale of the 12-term Horner value everywhere on the cell."""
    rng = np.random.default_rng(3)
    x = np.ascontiguousarray(np.concatenate([np.linspace(-1, 1, 1001), [-1.0, 1.0, 0.0]])[:1002])
    worst = 0.0
    for rep in range(200):
        rho = rng.uniform(0.01, 1 / 16)
        C = rng.uniform(-1, 1, 12) * rho ** np.arange(12)
        C[0] = rng.choice([-1.0, 1.0])
        full, econ = np.empty_like(x), np.empty_like(x)
        hm.h_far_poly(P(np.ascontiguousarray(C)), P(x), x.size, P(full), P(econ))
        assert np.abs(full - np.polyval(C[::-1], x)).max() < 1e-15
        worst = max(worst, np.abs(econ - full).max())
    assert worst < 2e-15, worst
    # and the bound is about the decay: a polynomial that does NOT decay loses digits, as it must
    C = np.ones(12)
    full, econ = np.empty_like(x), np.empty_like(x)
    hm.h_far_poly(P(C), P(x), x.size, P(full), P(econ))
    assert 1e-4 < np.abs(econ - full).max() < 4e-3          # |C11|/1024 + |C10|/512 = 2.9e-3


def test_stepsize_sum_reproduces_numpys_pairwise_order(hm):
    """pyswarm: stepsize = np.sqrt(np.sum((g - p_min)**2)).  The device commit sums the rounded squares in numpy's
    pairwise order (eight interleaved accumulators up to 128 elements, halves rounded to multiples of 8 beyond), so the
    `stepsize <= minstep` comparison sees the same double as the CPU run - for every parameter count 4 + 3P, P <= 256."""
    hm.h_numpy_sum.restype = ctypes.c_double
    rng = np.random.default_rng(11)
    sizes = sorted({4 + 3 * p for p in (1, 2, 6, 12, 24, 36, 41, 42, 43, 66, 85, 86, 170, 171, 256)} | set(range(1, 40)) |
                   {127, 128, 129, 255, 256, 257, 264, 511, 512, 513, 772})
    seq_differs = 0
    for n in sizes:
        for rep in range(20):
            d = rng.normal(size=n) * 10.0 ** rng.uniform(-9, 0, n)
            sq = d ** 2
            got = hm.h_numpy_sum(P(sq), n)
            assert got == np.sum(sq), (n, rep)
            seq = 0.0
            for v in sq:
                seq += v
            seq_differs += seq != got
    assert seq_differs > 100          # the order matters at the last ulp: a sequential sum is NOT what numpy computes
