"""GPU: cProfile of BASELINE config 1 (examples/grover_algorithm.py, d = 10) on the engine: where the host time goes."""
import cProfile
import os
import pstats
import sys
import warnings

warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append(os.path.join(ROOT, 'oracle', '_ref', 'examples'))
import numpy as np
import qcircuits_b200
qcircuits_b200.install_as_qcircuits()
import grover_algorithm as ga

for rep in range(2):
    np.random.seed(rep)
    f = ga.construct_problem(10)
    ga.grover_algorithm(10, f)
np.random.seed(5)
f = ga.construct_problem(10)
pr = cProfile.Profile()
pr.enable()
ga.grover_algorithm(10, f)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
