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
