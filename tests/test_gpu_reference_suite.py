"""The reference's OWN test-suite and examples, unmodified, on the CUDA engine ("(1) run tests.py
unmodified with `import qcircuits` resolving to the new engine").

oracle/_ref/tests/tests.py (95 unittest cases incl. the examples: Deutsch, Deutsch-Jozsa, Grover d = 8..10, teleportation,
superdense coding, Bell states, phase estimation -- reference tests/tests.py:1607-1697) is a byte-identical copy made
by oracle/make_ref.py.  It runs in a subprocess from its own directory (its imports are cwd-relative: tests.py:9-11,
1593) in which ``qcircuits`` has been bound to qcircuits_b200 by install_as_qcircuits() BEFORE the file is imported.
The reduced-count statistical sampling test (slow_measurement_tests.py:16-85) runs the same way."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, 'oracle', '_ref', 'tests')

BOOT = r'''
import sys, warnings
warnings.simplefilter('ignore')
sys.path.append(%(root)r)
import qcircuits_b200
qcircuits_b200.install_as_qcircuits()
import qcircuits
assert qcircuits is qcircuits_b200 and qcircuits.State is qcircuits_b200.State
# the suite draws its cases from the global NumPy / random streams and is unseeded; a few of its tests are statistical
# (phase-estimation failure bound, sampling): the harness fixes the seeds so that a run is reproducible
import random, numpy
random.seed(%(seed)d); numpy.random.seed(%(seed)d)
import pytest
sys.exit(pytest.main(%(args)r))
'''


def _run(args, timeout, seed=20260):
    code = BOOT % {'root': ROOT, 'args': args, 'seed': seed}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    return subprocess.run([sys.executable, '-c', code], cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_TESTS, 'tests.py')), reason='oracle/_ref has not been made')
def test_the_reference_suite_passes_unmodified_on_the_engine():
    res = _run(['tests.py', '-q', '-p', 'no:cacheprovider', '--no-header', '--tb=short'], 3000)
    tail = (res.stdout + res.stderr)[-6000:]
    print(tail)
    if res.returncode != 0:
        # a statistical test of the reference can fail by chance (it does on the reference itself): one more run with
        # another seed must pass completely
        res = _run(['tests.py', '-q', '-p', 'no:cacheprovider', '--no-header', '--tb=short'], 3000, seed=4711)
        tail = (res.stdout + res.stderr)[-6000:]
        print(tail)
    assert res.returncode == 0, tail
    assert ' passed' in res.stdout and 'failed' not in res.stdout.split('\n')[-2]
    import re
    m = re.search(r'(\d+) passed', res.stdout)
    assert m and int(m.group(1)) >= 95, res.stdout[-500:]


SAMPLING = r'''
import sys, warnings
warnings.simplefilter('ignore')
sys.path.append(%(root)r)       # after the working directory: `from tests import random_state` must find ./tests.py
import qcircuits_b200
qcircuits_b200.install_as_qcircuits()
import types
# the reference's statistical test draws 300 000 measurements x 3 with a tqdm progress bar (slow_measurement_tests.py:
# 16-85): same code, reduced count, no progress bar
sys.modules['tqdm'] = types.SimpleNamespace(tqdm=lambda it, *a, **k: it)
import random, numpy
random.seed(7); numpy.random.seed(7)
src = open('slow_measurement_tests.py').read()
assert src.count('samples = 300000') == 3 and src.count('prob_epsilon = 0.05') == 3
# 100 x fewer samples: the sampling error of a frequency grows 10 x, and so does the allowed proportional deviation
src = src.replace('samples = 300000', 'samples = 3000').replace('prob_epsilon = 0.05', 'prob_epsilon = 0.5')
ns = {'__name__': 'slow_measurement_tests_reduced'}
exec(compile(src, 'slow_measurement_tests.py', 'exec'), ns)
import unittest
suite = unittest.TestSuite()
for v in list(ns.values()):
    if isinstance(v, type) and issubclass(v, unittest.TestCase):
        suite.addTests(unittest.defaultTestLoader.loadTestsFromTestCase(v))
r = unittest.TextTestRunner(verbosity=1).run(suite)
print('ran', r.testsRun, 'failures', len(r.failures), 'errors', len(r.errors))
sys.exit(0 if r.wasSuccessful() and r.testsRun >= 1 else 1)
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_TESTS, 'slow_measurement_tests.py')), reason='oracle/_ref has not been made')
def test_the_reference_sampling_test_at_reduced_count():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    res = subprocess.run([sys.executable, '-c', SAMPLING % {'root': ROOT}], cwd=REF_TESTS, env=env, capture_output=True,
                         text=True, timeout=3000)
    tail = (res.stdout + res.stderr)[-2000:]
    print(tail)
    assert res.returncode == 0, tail
