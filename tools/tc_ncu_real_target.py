"""ncu target: the heaviest passes of the HEADLINE circuit (30 qubits complex64, depth 40) on the tensor-core kernel,
two launches each (run once plain, then under ncu)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _lib
    from qcircuits_b200.workloads import layered_circuit
    n = int(os.environ.get('NCU_N', '30'))
    lib = _lib.load()
    circ = layered_circuit(n, 40, seed=1234)
    comp = circ._compiled(np.complex64, True, 128)
    s = qc.zeros(n, dtype=np.complex64)
    s = qc.Hadamard()(s, qubit_indices=[0])
    buf = s._dv.buf.data_ptr()
    unit = torch.ones(2, dtype=torch.float64, device='cuda')     # the state is normalised: its lazy squared norm (ranges the fp16 operands)
    norm_in = unit.data_ptr()
    ws, stream = _lib.workspace().data_ptr(), _lib.stream_ptr()
    stats = (ctypes.c_int32 * 128)()
    for idx in (2, 8):                      # passes with four blocks and few register ops (L = 4 / 5 tiles)
        item = comp[idx]
        _lib.check(lib.qcs_plan_stats(n, 0, item[1], len(item[1]), stats))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.check(lib.qcs_apply_fused(buf, buf, n, 0, item[1], len(item[1]), norm_in, None, ws, stream))
        e0.record()
        _lib.check(lib.qcs_apply_fused(buf, buf, n, 0, item[1], len(item[1]), norm_in, None, ws, stream))
        e1.record()
        torch.cuda.synchronize()
        print('pass %d: %d gates, %d sweeps (%d with a block, %d gates in blocks), %d register ops: %.3f ms' % (
            idx, len(item[1]), stats[1], stats[7] & 255, stats[7] >> 8, stats[2] - stats[1], e0.elapsed_time(e1)))
    print('ncu target done')


if __name__ == '__main__':
    main()
