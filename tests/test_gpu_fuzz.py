"""Randomised differential run (tests/fuzz_parity.py): random shapes - points, peaks, particles, spectra, axis
direction - objective vs the CPU oracle (real-only and fit_im, 1e-10), device weights vs the oracle (bitwise), fused
swarm kernel vs the per-step kernels (bitwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('seed', [11, 12])
def test_random_shapes(seed):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'fuzz_parity.py'), '--cases', '10', '--seed', str(seed)],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    rep = json.loads(out.stdout)
    assert rep['failures'] == [] and rep['worst']['objective_rel'] < 1e-10


def test_adversarial_parameters():
    """tests/fuzz_wild.py: widths over seven decades (incl. the recurrence / exact-path switch), centres at the edges
    and outside the window, pure Lorentzian / Gaussian, large phases - both kernels, real-only and fit_im, 1e-9."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'fuzz_wild.py'), '5'], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    rep = json.loads(out.stdout)
    assert rep['n_bad'] == 0 and max(rep['worst'].values()) < 1e-9
