"""Phase estimation on the GPU (csrc/phase.cu) against the reference's own outputs (tests/golden/phase.npz:
Data._brute_phase, _ps_acme_score, approximate_phase) and the oracle.  Tolerances: the reference computes
exp(1j*phi) with the host libm, CUDA's sincos differs by <= 1 ulp -> errors/scores to 1e-12 relative, the selected
brute candidate exactly, the Nelder-Mead result to the simplex's own xtol (1e-4 degrees)."""
import numpy as np
import pytest

from conftest import load_golden
import nmrfit_b200
from nmrfit_b200 import _cabi, proc_autophase, synth
from oracle import nmrfit_oracle as orc

pytestmark = pytest.mark.gpu
CASES = 'abc'


@pytest.mark.parametrize('tag', CASES)
def test_brute_phase_matches_reference(tag):
    g = load_golden('phase')
    u, v = g['u_' + tag], g['v_' + tag]
    d = nmrfit_b200.containers.Data(np.arange(u.size, dtype=float), u, v)
    d.shift_phase(method='brute')
    assert d.p0 == g['brute_' + tag][0] and d.p1 == 0.0          # the reference's pick, exactly
    V, I = orc.ps2(u, v, d.p0, 0.0)
    assert np.max(np.abs(d.V - V)) < 1e-12
    # per-candidate errors and flags against the oracle's restatement of the scan
    cands, err, ok = orc.brute_phase_errors(u, v)
    with _cabi.PhaseScorer(u, v) as sc:
        best, berr, gerr, gok = sc.brute(cands, details=True)
    assert np.array_equal(gok[0], ok)
    assert np.max(np.abs(gerr[0] - err)) < 1e-13 * max(1.0, np.abs(u).max())
    assert berr[0] == gerr[0][gok[0]].min()


def test_brute_phase_batch_and_ragged_sizes():
    rows_u, rows_v, want = [], [], []
    for seed in range(5):
        data, true = synth.multiplet(3000, 6, seed=70 + seed)
        u, v = orc.ps2(data.u, data.v, 0.6 * seed - 1.0, 0.0, inv=True)
        rows_u.append(u); rows_v.append(v)
        want.append(orc.brute_phase(u, v)[0])
    got = proc_autophase.brute_phase_batch(np.array(rows_u), np.array(rows_v))
    assert np.array_equal(got, np.array(want))
    # a spectrum that never points upwards keeps the reference's default p0 = 0 (containers.py:99)
    z = np.zeros(700)
    assert proc_autophase.brute_phase_batch(z[None], z[None])[0] == 0.0
    # N >= 10,000: the baseline means cover n = N // 5000 > 1 points (numpy's summation order)
    data, _ = synth.multiplet(66000, 6, seed=9)
    assert proc_autophase.brute_phase_batch(data.u[None], data.v[None], step=np.pi / 90)[0] == \
        orc.brute_phase(data.u, data.v, step=np.pi / 90)[0]


def test_brute_phase_long_axis_means_over_more_than_128_points():
    """N >= 645,000: the baseline means V[:n].mean(), V[-n:].mean() (containers.py:103-104, n = N // 5000) cover more
    than 128 points, where numpy's pairwise summation starts to split recursively - the device follows it, so the
    per-candidate errors are the oracle's to rounding and the pick is the same."""
    N = 1_300_000                                          # n = 260: two levels of the recursion
    rng = np.random.default_rng(3)
    data, _ = synth.multiplet(4096, 6, seed=5)
    u = np.interp(np.linspace(0, 1, N), np.linspace(0, 1, 4096), data.u) + rng.normal(0, 1e-3, N)
    v = np.interp(np.linspace(0, 1, N), np.linspace(0, 1, 4096), data.v) + rng.normal(0, 1e-3, N)
    cands, err, ok = orc.brute_phase_errors(u, v, step=np.pi / 30)
    with _cabi.PhaseScorer(u, v) as sc:
        best, berr, gerr, gok = sc.brute(cands, details=True)
    assert np.array_equal(gok[0], ok)
    assert np.max(np.abs(gerr[0] - err)) < 1e-13 * max(1.0, np.abs(u).max())      # (a 128-point window is off by ~1e-6 here)
    assert best[0] == orc.brute_phase(u, v, step=np.pi / 30)[0]


@pytest.mark.parametrize('tag', CASES)
def test_acme_score_matches_reference(tag):
    g = load_golden('phase')
    z = g['u_' + tag] + 1j * g['v_' + tag]
    got = np.array([proc_autophase._ps_acme_score(ph, z) for ph in g['acme_ph_' + tag]])
    assert np.max(np.abs(got / g['acme_score_' + tag] - 1)) < 1e-12
    batch = proc_autophase.acme_score_batch(g['acme_ph_' + tag], g['u_' + tag][None], g['v_' + tag][None])
    assert np.array_equal(batch[0], got)


@pytest.mark.parametrize('tag', CASES)
def test_approximate_phase_matches_reference(tag):
    g = load_golden('phase')
    d = nmrfit_b200.containers.Data(np.arange(g['u_' + tag].size, dtype=float), g['u_' + tag], g['v_' + tag])
    d.shift_phase(method='auto')
    ref = g['auto_' + tag]
    # same simplex, same start, scores equal to ~1e-15: the optimum agrees far inside fmin's xtol (1e-4 deg = 1.7e-6 rad)
    assert abs(d.p0 - ref[0]) < 2e-6 and abs(d.p1 - ref[1]) < 2e-6
    z = g['u_' + tag] + 1j * g['v_' + tag]
    s_got = orc.acme_score((d.p0 * 180 / np.pi, d.p1 * 180 / np.pi), z)
    s_ref = orc.acme_score((ref[0] * 180 / np.pi, ref[1] * 180 / np.pi), z)
    assert abs(s_got / s_ref - 1) < 1e-8


def test_ps_degrees_matches_oracle():
    rng = np.random.default_rng(4)
    z = rng.normal(size=333) + 1j * rng.normal(size=333)
    for p0, p1, inv in ((30.0, -12.0, False), (-170.0, 400.0, True)):
        assert np.max(np.abs(proc_autophase.ps(z, p0, p1, inv) - orc.ps(z, p0, p1, inv))) < 1e-13
