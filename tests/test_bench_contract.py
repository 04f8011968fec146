"""bench.py's reference arm runs on the CPU: it times the UNMODIFIED reference's nmrfit.equations.objective from
baseline/_ref (or /root/reference), falling back to the oracle port only where neither exists.  One tiny invocation, and
the JSON line carries the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c1',
                          '--steps', '2', '--warmup', '3'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['value'] > 0 and d['higher_is_better'] is True
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    want_kind = 'reference' if ref_loader.reference_root() else 'port'
    assert d['cpu_baseline']['kind'] == want_kind and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['vs_baseline'] is None and 'workload' in d['config']


def test_default_workload_is_the_shape_the_metric_is_quoted_on():
    """BASELINE.json: "Voigt objective evals/s (6 peaks, 4k pts)" - both arms default to that shape and describe it with
    the same config, so the driver's ratio compares like with like."""
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS['metric']
    assert (wl['P'], wl['N']) == (6, 4096)
    cfg = bench.workload_config('metric', 1)
    assert cfg['n_peaks'] == 6 and cfg['n_points'] == 4096 and cfg['workload'] == wl['title']
    import argparse
    saved = sys.argv
    try:
        sys.argv = ['bench.py']
        captured = {}
        bench.run_b200 = lambda a: captured.setdefault('args', a) and 0
        bench.main()
        assert captured['args'].workload == 'metric' and captured['args'].gpus == 1 and captured['args'].warmup >= 3
    finally:
        sys.argv = saved


def test_reference_arm_under_torchrun_ranks_other_than_zero_do_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
