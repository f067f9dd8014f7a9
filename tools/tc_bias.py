"""GPU: systematic norm drift of the tensor-core sweep (the TF32 split truncates: hi = top 19 bits, and the tensor core
truncates lo again, so every block shrinks the amplitudes by a tiny, statistically constant factor).  Applies passes of
four random unitary blocks repeatedly without renormalising and prints the squared norm; QCS_TC=0 gives the
register-sweep kernel for comparison."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _lib, _gates
    from qcircuits_b200._device import make_ops
    n = int(os.environ.get('PT_N', '26'))
    lib = _lib.load()
    rng = np.random.default_rng(0)

    def runit(k):
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        Q, R = np.linalg.qr(A)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    s = qc.zeros(n, dtype=np.complex64)
    for q in range(n):
        s = qc.Hadamard()(s, qubit_indices=[q])
    s = s._dv.scaled(1.0)
    buf = s.buf.data_ptr()
    ws, stream, code = _lib.workspace().data_ptr(), _lib.stream_ptr(), 0
    sets = [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [2, 3, 4, 5, 6], [7, 8, 9, 10, 11]]
    norm = _lib.alloc_scalar()
    out = []
    total_blocks = 0
    for rep in range(int(os.environ.get('PT_REPS', '40'))):
        gates = []
        for bset in sets:
            gates += [(runit(2), [bset[0], bset[1]]), (runit(2), [bset[2], bset[3]]), (runit(2), [bset[3], bset[4]]),
                      (runit(2), [bset[1], bset[2]]), (runit(1), [bset[0]]), (runit(1), [bset[4]])]
        ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
        stats = (ctypes.c_int32 * 128)()
        _lib.check(lib.qcs_plan_stats(n, code, ops, len(ops), stats))
        _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), None, norm.data_ptr(), ws, stream))
        total_blocks += stats[7] & 255
        if rep % 8 == 7:
            out.append({'block_sweeps': total_blocks, 'norm2_minus_1': float(norm[0].item()) - 1.0})
            print(out[-1], flush=True)
    per = out[-1]['norm2_minus_1'] / out[-1]['block_sweeps']
    print(json.dumps({'norm2_drift_per_block_sweep': per, 'amplitude_factor_per_block_sweep': 1 + per / 2}))


if __name__ == '__main__':
    main()
