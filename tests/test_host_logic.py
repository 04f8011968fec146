"""Host-side logic of the package (no GPU): weights, Data, bounds, synthetic inputs,
option plumbing, sharding helpers."""
import numpy as np
import pytest

from conftest import load_golden, peaks_from_golden
import nmrfit_b200
from nmrfit_b200 import synth, utils, containers, swarm, equations, _cabi


def test_api_surface_matches_reference():
    import inspect
    sig = inspect.signature(nmrfit_b200.fit)
    assert list(sig.parameters) == ['data', 'lower', 'upper', 'expon', 'dynamic_weighting', 'fit_im', 'processes',
                                    'summary', 'options']
    assert sig.parameters['expon'].default == 0.5 and sig.parameters['processes'].default == 1
    sig = inspect.signature(utils.FitUtility.__init__)
    assert list(sig.parameters)[1:] == ['data', 'lower', 'upper', 'expon', 'dynamic_weighting', 'fit_im', 'processes',
                                        'summary', 'options']
    for name in ('fit', 'generate_result', 'calculate_area_fraction', 'get_areas', '_compute_weights', '_print_summary'):
        assert hasattr(utils.FitUtility, name)
    assert list(inspect.signature(equations.objective).parameters) == ['x', 'w', 'u', 'v', 'weights', 'fit_im']
    assert list(inspect.signature(equations.voigt).parameters) == ['w', 'r', 'yoff', 'width', 'loc', 'a']
    assert list(inspect.signature(nmrfit_b200.proc_autophase.ps2).parameters) == ['u', 'v', 'p0', 'p1', 'inv']


@pytest.mark.parametrize('case', ['c1_4096x6', 'ragged_1000x6', 'p24_1536'])
def test_weights_and_bounds_match_reference(case):
    g = load_golden('objective_' + case)
    peaks = peaks_from_golden(g)
    assert np.array_equal(utils.compute_weights(g['w'], peaks), g['weights'])
    d = containers.Data(g['w'], g['u'], g['v'])
    d.set_peaks(peaks)
    lo, up = d.generate_solution_bounds()
    assert np.array_equal(lo, g['lower']) and np.array_equal(up, g['upper'])
    d.p0, d.p1 = 0.2, -0.1
    lo, up = d.generate_solution_bounds(force_p0=True, force_p1=True)
    assert lo[:2] == [0.2 - 0.001, -0.1 - 0.001] and up[:2] == [0.2 + 0.001, -0.1 + 0.001]


def test_weights_reversed_bounds_and_overwrite_order():
    w = np.linspace(0, 1, 101)

    class P:
        pass
    a, b = P(), P()
    a.bounds, a.height = [0.6, 0.2], 1.0        # reversed: indices are swapped
    b.bounds, b.height = [0.4, 0.5], -0.25      # later peak overwrites; |height| is used
    from oracle import nmrfit_oracle as orc
    got = utils.compute_weights(w, [a, b], expon=0.5)
    assert np.array_equal(got, orc.compute_weights(w, [a, b], expon=0.5))
    assert got[45] > 1.5 and got[0] == 1.0 and got[-1] == 1.0


def test_synth_is_deterministic_and_matches_golden_inputs():
    g = load_golden('objective_c1_4096x6')
    d, true = synth.multiplet(4096, 6, seed=0)
    assert np.array_equal(d.w, g['w']) and np.array_equal(d.u, g['u']) and np.array_equal(d.v, g['v'])
    assert np.array_equal(true, g['true'])
    d12, t12 = synth.multiplet(512, 12, seed=5)
    assert len(d12.peaks) == 12 and t12.size == 40
    with pytest.raises(ValueError):
        synth.multiplet(100, 5)
    mains, sats = d.peaks.split()
    assert len(mains) == 2 and len(sats) == 4
    assert abs(d.approximate_area_fraction() - 0.04 / 2.04) < 1e-12


def test_data_contract():
    d, _ = synth.multiplet(300, 6)
    n0 = d.w.size
    d.select_bounds(3.3, 3.5)
    assert d.w.min() > 3.3 and d.w.max() < 3.5 and d.w.size < n0 and d.u.size == d.w.size
    with pytest.raises(ValueError):
        d.shift_phase('nonsense')
    with pytest.raises(NotImplementedError):
        d.select_peaks(method='manual', n=6)                 # the interactive selector is not provided
    with pytest.raises(ValueError):
        d.select_peaks(method='nonsense')


def test_fit_im_identity_semantics():
    assert equations._fit_im_mode(True) == _cabi.IM_REFERENCE
    assert equations._fit_im_mode(1) == _cabi.REAL_ONLY          # truthy but not True: real only (equations.py:184)
    assert equations._fit_im_mode(False) == _cabi.REAL_ONLY
    assert equations._fit_im_mode('sum') == _cabi.IM_SUM


def test_bounds_check_messages():
    with pytest.raises(AssertionError, match='greater than lower-bound'):
        swarm._check_bounds([0, 1], [1, 1])
    with pytest.raises(AssertionError, match='same length'):
        swarm._check_bounds([0, 1], [1])


def test_shard_range_partitions():
    for total in (1, 7, 100, 204, 65536):
        for world in (1, 2, 3, 4, 8):
            parts = [swarm.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (o1, c1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_select_record_first_index_tie_break():
    recs = np.array([[0.5, 40.0, 1, 1], [0.25, 90.0, 2, 2], [0.25, 70.0, 3, 3], [np.inf, 0.0, 4, 4]])
    assert list(swarm.select_record(recs)) == [0.25, 70.0, 3, 3]


def test_area_helpers():
    f = utils.FitUtility(None, None, None)
    f.params = np.array([0, 0, .5, 0, 1, 3.3, 0.01, 1, 3.4, 1.0, 1, 3.5, 0.03])
    assert list(f.get_areas()) == [0.01, 1.0, 0.03]
    assert abs(f.calculate_area_fraction() - 0.04 / 1.04) < 1e-15


def test_draw_order_matches_pyswarm():
    np.random.seed(5)
    a = np.random.uniform(size=(3, 2)); b = np.random.uniform(size=(3, 2)); c = np.random.uniform(size=(3, 2))
    rs = np.random.RandomState(5)
    rp, rg = swarm._draw_generations(rs, 2, 3, 2)
    assert np.array_equal(rp[0], a) and np.array_equal(rg[0], b) and np.array_equal(rp[1], c)


def test_pyswarm_compat_refuses_what_it_cannot_accelerate():
    from nmrfit_b200 import pyswarm_compat
    with pytest.raises(NotImplementedError, match='nmrfit objective'):
        pyswarm_compat.pso(lambda x: float(np.sum(x ** 2)), [0, 0], [1, 1])
    from nmrfit_b200 import equations
    assert pyswarm_compat._is_nmrfit_objective(equations.objective)
    # a look-alike - same name, a module that merely ends in "equations" - is not silently replaced by the nmrfit objective
    import types
    fake = types.ModuleType('other.equations')
    exec('def objective(x, *a):\n    return 0.0', fake.__dict__)
    fake.objective.__module__ = 'other.equations'
    assert not pyswarm_compat._is_nmrfit_objective(fake.objective)
    impostor = types.FunctionType(fake.objective.__code__, {}, 'objective')
    impostor.__module__ = 'nmrfit_b200.equations'        # claims the module, but is not the module's attribute
    assert not pyswarm_compat._is_nmrfit_objective(impostor)
    with pytest.raises(NotImplementedError, match='constraints'):
        pyswarm_compat.pso(equations.objective, [0] * 7, [1] * 7, ieqcons=[lambda x: 1.0], args=(None,) * 5)
    with pytest.raises(ValueError, match='args'):
        pyswarm_compat.pso(equations.objective, [0] * 7, [1] * 7, args=(1, 2))
