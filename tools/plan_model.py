"""CPU-only: plan every fused pass of the headline circuit (qcs_plan_stats) and predict its time with the measured
model of profiles/r2_pass_times.json: t = 2.7 ms + 1.25 ms x blocks + 1.3 ms x register sweeps + 0.1 ms x register ops
(30 qubits complex64).  QCS_LIB points at an alternative build of libqcs (planner experiments)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from qcircuits_b200 import _lib
    from qcircuits_b200.workloads import layered_circuit
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    max_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    verbose = len(sys.argv) > 3
    lib = _lib.load()
    circ = layered_circuit(n, 40, seed=1234)
    comp = circ._compiled(np.complex64, True, max_ops)
    stats = (ctypes.c_int32 * 128)()
    tot = dict(passes=0, blocks=0, regsweeps=0, regops=0, gates_in_blocks=0, ms=0.0)
    for item in comp:
        _lib.check(lib.qcs_plan_stats(n, _lib.dtype_code(np.complex64), item[1], len(item[1]), stats))
        nb = stats[7] & 255
        sweeps, regops = stats[1], stats[2] - stats[1]
        ms = 2.7 + 1.25 * nb + 1.3 * (sweeps - nb) + 0.1 * regops
        tot['passes'] += 1; tot['blocks'] += nb; tot['regsweeps'] += sweeps - nb; tot['regops'] += regops
        tot['gates_in_blocks'] += stats[7] >> 8; tot['ms'] += ms
        if verbose:
            print('pass ops %3d blocks %d regsweeps %d regops %3d  -> %.1f ms' % (len(item[1]), nb, sweeps - nb, regops, ms))
    print(tot)


if __name__ == '__main__':
    main()
