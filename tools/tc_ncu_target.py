"""A short, fixed sequence of tensor-core tile-kernel launches for ncu (run once plain, then under ncu):
synthetic passes of four dense blocks at NCU_N qubits (default 28), complex64."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _gates, _lib
    from qcircuits_b200._device import make_ops
    n = int(os.environ.get('NCU_N', '28'))
    rng = np.random.default_rng(0)

    def runit(k):
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        Q, R = np.linalg.qr(A)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    s = qc.zeros(n, dtype=np.complex64)
    s = qc.Hadamard()(s, qubit_indices=[0])
    buf = s._dv.buf.data_ptr()
    lib = _lib.load()
    ws, stream = _lib.workspace().data_ptr(), _lib.stream_ptr()
    variants = {'low': [[0, 1, 2, 3, 4], [0, 1, 5, 6, 7], [0, 1, 8, 9, 10], [0, 2, 4, 6, 11]],
                'mid': [[2, 3, 4, 5, 6], [7, 8, 9, 10, 11], [1, 3, 5, 7, 9], [4, 5, 6, 7, 8]]}
    for name in os.environ.get('NCU_VARIANTS', 'low,mid').split(','):
        gates = []
        for bset in variants[name]:
            gates += [(runit(2), [bset[0], bset[1]]), (runit(2), [bset[2], bset[3]]), (runit(2), [bset[3], bset[4]]),
                      (runit(2), [bset[1], bset[2]])]
        ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.check(lib.qcs_apply_fused(buf, buf, n, 0, ops, len(ops), None, None, ws, stream))
        e0.record()
        for _ in range(2):
            _lib.check(lib.qcs_apply_fused(buf, buf, n, 0, ops, len(ops), None, None, ws, stream))
        e1.record()
        torch.cuda.synchronize()
        print(name, 'ms per pass', e0.elapsed_time(e1) / 2)
    print('ncu target done')


if __name__ == '__main__':
    main()
