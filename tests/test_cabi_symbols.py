"""The shared library builds for sm_100a without a GPU, loads, and exports every symbol
include/nmrfit_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'nmrfit_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nmrfit_[a-z0-9_]+)\s*\(', text)))


def test_exports_match_header(built):
    from nmrfit_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 25
    handle = ctypes.CDLL(built)
    for n in names:
        assert hasattr(handle, n), n
    # the ctypes table binds exactly the declared set
    assert sorted(_cabi.SIGNATURES) == names


def test_abi_version_and_error_string(built):
    from nmrfit_b200 import _cabi
    lib = _cabi.lib()
    assert lib.nmrfit_abi_version() == 2
    assert isinstance(lib.nmrfit_last_error(), bytes)
    assert _cabi.launch_count() >= 0


def test_pso_opts_layout_matches_header():
    from nmrfit_b200 import _cabi
    # struct nmrfit_pso_opts: 2 int, 5 double, 2 int, u64, 2 x i64  -> 80 bytes with natural alignment
    assert ctypes.sizeof(_cabi.PsoOpts) == 80
    assert _cabi.PsoOpts.omega.offset == 8 and _cabi.PsoOpts.fit_im.offset == 48
    assert _cabi.PsoOpts.seed.offset == 56 and _cabi.PsoOpts.particle_offset.offset == 64
    assert _cabi.PsoOpts.spectrum_offset.offset == 72


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nmrfit_b200 import _cabi
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _cabi.lib()


def test_c_consumer_compiles_against_the_header(tmp_path):
    """examples/c_abi_demo.c is plain C99 using only include/nmrfit_b200.h (no Python, no torch)."""
    import subprocess
    exe = str(tmp_path / 'c_abi_demo')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-O2', '-I', os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'examples', 'c_abi_demo.c'), '-o', exe, '-ldl', '-lm'], check=True)


@pytest.mark.gpu
def test_c_consumer_runs_on_the_gpu(tmp_path):
    import subprocess
    exe = str(tmp_path / 'c_abi_demo')
    subprocess.run(['gcc', '-std=c99', '-O2', '-I', os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'examples', 'c_abi_demo.c'), '-o', exe, '-ldl', '-lm'], check=True)
    out = subprocess.run([exe, os.path.join(ROOT, 'nmrfit_b200', 'csrc', 'libnmrfit_b200.so')], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'swarm:' in out.stdout
