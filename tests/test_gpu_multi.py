"""NCCL path of the particle-sharded swarm on real GPUs (skipped with fewer than two):
tools/dist_check.py under torch.distributed.run, sharded result bit-identical to the one-GPU run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from nmrfit_b200 import _cabi
    return _cabi.device_count()


@pytest.mark.parametrize('exchange', ['nccl', 'p2p'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_swarm_over_nccl_is_bit_identical(world, exchange):
    """exchange='nccl': one all-gather of best records per generation; 'p2p': the exchange + commit kernel over peer
    memory (CUDA IPC windows, NVLink stores), no collective per generation."""
    if _gpus() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29520 + world + (10 if exchange == 'p2p' else 0)), os.path.join(ROOT, 'tools', 'dist_check.py'),
           '--shape', 'c1', '--swarm', '250', '--maxiter', '30', '--exchange', exchange]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert line['world'] == world and line['identical_on_all_ranks'] and line['bit_identical_to_one_gpu']


@pytest.mark.parametrize('world', [2, 4])
def test_spectra_sharded_batch_over_nccl_is_bit_identical(world):
    """fit_batch_sharded (contiguous blocks of spectra per rank, no data-path collective, one final all-gather) against
    fit_batch of the whole batch on one GPU."""
    if _gpus() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29540 + world), os.path.join(ROOT, 'tools', 'dist_check.py'),
           '--mode', 'spectra', '--maxiter', '25']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert line['world'] == world and line['identical_on_all_ranks'] and line['bit_identical_to_one_gpu']
