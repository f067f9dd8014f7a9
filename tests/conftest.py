import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v1.npz'))


def pytest_terminal_summary(terminalreporter):
    """achieved error / tolerance of the parity assertions (tests/util.check_close), worst first; also written to
    gpurun_out/parity_achieved.json when that directory exists"""
    from tests import util
    if not util.ACHIEVED:
        return
    rows = sorted(util.ACHIEVED.items(), key=lambda kv: -kv[1])
    terminalreporter.write_line('parity: achieved error / tolerance (1e-5 complex64, 1e-12 complex128)')
    for key, ratio in rows[:40]:
        terminalreporter.write_line('  %-86s %.4f' % (key, ratio))
    out = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(out):
        import json
        with open(os.path.join(out, 'parity_achieved.json'), 'w') as f:
            json.dump(dict(rows), f, indent=1)
