"""Sharded path on real GPUs: needs at least 2 devices (skipped on the single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_circuit_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs 2 GPUs')
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29653', os.path.join(ROOT, 'tools', 'dist_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'dist_check ok' in out.stdout


def test_exchange_kernel_one_process_per_gpu():
    """qcs_dist_swap_rank + CUDA IPC (the exchange ShardedState uses) is bit-identical to all_to_all_single"""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs 2 GPUs')
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29655', os.path.join(ROOT, 'tools', 'dist_exchange_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, QX_BITS='26'))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'dist_exchange_check ok' in out.stdout


def test_exchange_kernel_single_process_c_abi():
    """qcs_dist_init / qcs_dist_swap / qcs_dist_destroy (one process, all GPUs) against a NumPy index map: random
    subsets of rank bits against random local bits, both amplitude types, pushing and pulling"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    for push in ('1', '0'):
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dist_exchange_check.py'), 'single'],
                             capture_output=True, text=True, timeout=900, env=dict(os.environ, QX_BITS='26', QCS_DIST_PUSH=push))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert 'dist_exchange_single ok' in out.stdout
