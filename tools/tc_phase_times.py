"""GPU: where a tensor-core sweep spends its time.  Needs a scratch build with -DQCS_TC_DBG (tools/build_variant.sh dbg
-DQCS_TC_DBG; run with QCS_LIB=build/dbg/libqcs.so).  Times synthetic passes of 1..4 dense blocks with parts of the
sweep switched off (QCS_TC_DEBUG bits: 1 no MMA, 2 no split / tcgen05.st, 4 no tcgen05.ld, 8 no tile LDS / STS) --
the results of those runs are wrong by construction, only the times mean something."""
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
    n = int(os.environ.get('PT_N', '30'))
    lib = _lib.load()
    s = qc.zeros(n, dtype=np.complex64)
    s = qc.Hadamard()(s, qubit_indices=[0])
    buf = s._dv.buf.data_ptr()
    ws, stream, code = _lib.workspace().data_ptr(), _lib.stream_ptr(), s._dv.code
    rng = np.random.default_rng(0)

    def runit(k):
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        Q, R = np.linalg.qr(A)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    set_lists = {'std': [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [2, 3, 4, 5, 6], [7, 8, 9, 10, 11]],
                 'low': [[0, 1, 2, 3, 4], [0, 1, 2, 3, 5], [0, 1, 2, 3, 6], [0, 1, 2, 3, 7]],
                 'high': [[7, 8, 9, 10, 11], [6, 8, 9, 10, 11], [5, 8, 9, 10, 11], [4, 8, 9, 10, 11]]}
    out = []
    for name, sets in set_lists.items():
        for nblocks in (1, 2, 4):
            gates = []
            for bset in sets[:nblocks]:
                gates += [(runit(2), [bset[0], bset[1]]), (runit(2), [bset[2], bset[3]]), (runit(2), [bset[3], bset[4]]),
                          (runit(2), [bset[1], bset[2]])]
            ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
            for dbg in [int(x) for x in os.environ.get('PT_DBG', '0,1,2,4,8,3,7,15,6,14').split(',')]:
                os.environ['QCS_TC_DEBUG'] = str(dbg)
                _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), None, None, ws, stream))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), None, None, ws, stream))
                e1.record()
                torch.cuda.synchronize()
                rec = {'sets': name, 'blocks': nblocks, 'dbg': dbg, 'ms': e0.elapsed_time(e1) / 4}
                out.append(rec)
                print(rec, flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/tc_phase_times.json', 'w'))


if __name__ == '__main__':
    main()
