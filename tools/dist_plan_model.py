"""CPU-only: passes per step of the sharded schedule (dist_schedule + schedule_passes per communication-free step) next to
the single-GPU schedule of the same circuit.  usage: dist_plan_model.py <qubits> <log2 ranks>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from qcircuits_b200.workloads import layered_circuit
    from qcircuits_b200.dist import dist_schedule, specialize, initial_layout, _LogicalGate
    from qcircuits_b200.circuit import _Gate, schedule_passes, merge_single_qubit_gates, expand_swaps
    n, g = int(sys.argv[1]), int(sys.argv[2])
    circ = layered_circuit(n, 40, seed=1234)
    n_local = n - g
    single = schedule_passes(merge_single_qubit_gates(expand_swaps(circ.gates)), n, np.complex64)
    print('single-GPU schedule of the %d-qubit circuit: %d gates, %d passes' % (n, len(circ.gates), len(single)))
    gates = [_LogicalGate(gt.plan, [circ.n - 1 - b for b in gt.bits]) for gt in circ.gates]
    steps, phys = dist_schedule(gates, n, g, initial_layout(n, g))
    for rank in (0, (1 << g) - 1):
        total = 0
        for step in steps:
            if step[0] == 'local':
                rank_gates = []
                for idx, bits in step[1]:
                    sp = specialize(gates[idx].plan, bits, n_local, rank)
                    if sp is not None:
                        rank_gates.append(_Gate(None, sp[1], plan=sp[0]))
                merged = merge_single_qubit_gates(expand_swaps(rank_gates))
                groups = schedule_passes(rank_gates, n_local, np.complex64) if rank_gates else []
                groups_m = schedule_passes(merged, n_local, np.complex64) if merged else []
                print('  rank %d local step: %d gates (%d after specialisation) -> %d passes; merged: %d gates -> %d passes' % (rank, len(step[1]), len(rank_gates), len(groups), len(merged), len(groups_m)))
                total += len(groups)
            else:
                print('  rank %d step: %s' % (rank, step[0]))
        print('rank %d: %d passes' % (rank, total))


if __name__ == '__main__':
    main()
