"""A short, fixed sequence of tile-kernel launches for ncu (run once plain, then under ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200._device import make_ops
    from qcircuits_b200.workloads import layered_circuit

    n = int(os.environ.get('NCU_N', '30'))
    s = qc.zeros(n, dtype=np.complex64)
    s = qc.Hadamard()(s, qubit_indices=[0])
    dst = s._dv._new_like()
    RX, H = qc.RotationX(0.3), qc.Hadamard()
    seqs = [[(RX._plan(), [n - 1], False)], [(RX._plan(), [0], False)], [(RX._plan(), [5], False)]]
    for plans in seqs:
        ops, keep = make_ops(plans)
        for _ in range(3):
            s._dv.apply_ops(ops, keep, out=dst)
    torch.cuda.synchronize()
    # one fused pass of the headline circuit (the schedule's first passes), three launches each
    circ = layered_circuit(n, 4, seed=1234)
    comp = circ._compiled(np.complex64, True, int(os.environ.get('NCU_MAX_OPS', '12')))
    from qcircuits_b200 import _lib
    lib = _lib.load()
    for item in comp[2:5]:
        for _ in range(2):
            _lib.check(lib.qcs_apply_fused(dst.buf.data_ptr(), s._dv.buf.data_ptr(), n, s._dv.code, item[1], len(item[1]),
                                           None, None, _lib.workspace().data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    del s, dst
    s = qc.zeros(n - 2, dtype=np.complex128)
    s = qc.Hadamard()(s, qubit_indices=[0])
    dst = s._dv._new_like()
    ops, keep = make_ops([(RX._plan(), [n - 3], False)])
    for _ in range(3):
        s._dv.apply_ops(ops, keep, out=dst)
    torch.cuda.synchronize()
    print('ncu target done')


if __name__ == '__main__':
    main()
