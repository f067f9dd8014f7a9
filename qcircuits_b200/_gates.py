"""Host-side structure analysis of gate matrices.

A gate handed to the kernels is described by a ``GatePlan``: which operator qubits are pure
controls (the gate is |0><0| (x) I + |1><1| (x) U' on them), which are targets, and whether the
remaining target matrix is diagonal.  Controls cost nothing in the tile kernel (a predicate on the
amplitude index) and diagonal gates may address any bit of the state, so recognising CNOT / CZ /
Toffoli / ControlledU(R_k) / RotationZ here is what lets them fuse freely.
The tests are exact (== 0, == 1): the factories of operators.py produce exact zeros and ones.
"""
import numpy as np


ANALYZE_MAX_QUBITS = 6        # wider operators are dense-kernel operands: no structure analysis (Operator._plan)


class GatePlan:
    __slots__ = ('k', 'controls', 'targets', 'diagonal', 'matrix', 'full')

    def __init__(self, k, controls, targets, diagonal, matrix, full):
        self.k = k                  # operator qubits
        self.controls = controls    # operator-qubit indices that are controls
        self.targets = targets      # operator-qubit indices the reduced matrix acts on (MSB-first)
        self.diagonal = diagonal    # reduced matrix is diagonal -> ``matrix`` holds the 2^t diagonal
        self.matrix = matrix        # contiguous complex128: (2^t, 2^t) or (2^t,)
        self.full = full            # the original (2^k, 2^k) matrix


def _blocks(M, nq, p):
    """View M (2^nq x 2^nq) as 2x2 blocks over qubit position p (0 = most significant)."""
    t = M.reshape([2] * (2 * nq))
    t = np.moveaxis(t, (p, nq + p), (0, 1))            # (out_p, in_p, other outs..., other ins...)
    side = 2 ** (nq - 1)
    return t.reshape(2, 2, side, side)


def analyze(M):
    M = np.ascontiguousarray(np.asarray(M, dtype=np.complex128))
    k = int(round(np.log2(M.shape[0])))
    qubits = list(range(k))
    controls = []
    cur = M
    progress = True
    while progress and len(qubits) > 1:
        progress = False
        nq = len(qubits)
        for p in range(nq):
            B = _blocks(cur, nq, p)
            if (not B[0, 1].any()) and (not B[1, 0].any()) and np.array_equal(B[0, 0], np.eye(2 ** (nq - 1))):
                controls.append(qubits.pop(p))
                cur = np.ascontiguousarray(B[1, 1])
                progress = True
                break
    diagonal = not (cur - np.diag(np.diag(cur))).any()
    red = np.ascontiguousarray(np.diag(cur)) if diagonal else np.ascontiguousarray(cur)
    return GatePlan(k, controls, qubits, diagonal, red, M)
