"""BASELINE config 1: examples/grover_algorithm.py (the vendored, unmodified file) at d = 10, full circuit + measurement,
timed end to end (a) on the CUDA engine with `qcircuits` bound to qcircuits_b200 and (b) on the unmodified reference on
this host's CPU.  Each leg runs in its own process; prints one JSON object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, 'oracle', '_ref', 'examples')

LEG = r'''
import sys, time, json, warnings
warnings.simplefilter('ignore')
sys.path.append(%(root)r)
if %(engine)r == 'b200':
    import qcircuits_b200
    qcircuits_b200.install_as_qcircuits()
else:
    sys.path.insert(0, %(ref)r)
import numpy as np
import grover_algorithm as ga
times, ok = [], []
for rep in range(%(reps)d):
    np.random.seed(rep)
    t0 = time.perf_counter()
    f = ga.construct_problem(10)
    bits = ga.grover_algorithm(10, f)
    times.append(time.perf_counter() - t0)
    ok.append(int(f(*bits)) == 1)
print(json.dumps({'seconds': times, 'found': ok}))
'''


def leg(engine, reps):
    code = LEG % {'root': ROOT, 'engine': engine, 'ref': os.path.join(ROOT, 'oracle', '_ref'), 'reps': reps}
    out = subprocess.run([sys.executable, '-c', code], cwd=EX, capture_output=True, text=True, timeout=1200,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    if out.returncode != 0:
        return {'error': (out.stdout + out.stderr)[-800:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


if __name__ == '__main__':
    res = {'workload': 'examples/grover_algorithm.py, d = 10 (11 qubits), 25 Grover iterations + measurement, complex128',
           'b200': leg('b200', 4), 'reference_cpu': leg('reference', 3)}
    for k in ('b200', 'reference_cpu'):
        if 'seconds' in res[k]:
            res[k]['best_s'] = min(res[k]['seconds'])
    print(json.dumps(res))
