"""GPU: clock-stamp timeline of a tensor-core pass (scratch build with -DQCS_TC_DBG: tools/build_variant.sh dbg
-DQCS_TC_DBG; run with QCS_LIB=build/dbg/libqcs.so).  Thread 0 of CTA 0 (group 0) stamps clock64 at the phase
boundaries of four steady-state tiles; this prints the cycles between consecutive stamps."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TAGS = {1: 'sweep start', 2: 'LDS issued', 3: 'regops done', 4: 'split+st issued', 5: 'st complete', 6: 'barrier 1',
        7: 'MMAs issued', 8: 'MMAs complete', 9: 'barrier 2', 10: 'D in registers', 11: 'STS issued', 12: 'barrier 3',
        20: 'tile start', 21: 'cp.async landed', 22: 'tile barrier', 23: 'sweeps done', 24: 'store-out issued'}


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

    sets = [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [2, 3, 4, 5, 6], [7, 8, 9, 10, 11]]
    gates = []
    for bset in sets[:int(os.environ.get('PT_BLOCKS', '4'))]:
        gates += [(runit(2), [bset[0], bset[1]]), (runit(2), [bset[2], bset[3]]), (runit(2), [bset[3], bset[4]]),
                  (runit(2), [bset[1], bset[2]])]
    ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
    for dbg in [int(x) for x in os.environ.get('PT_DBG', '0').split(',')]:
        os.environ['QCS_TC_DEBUG'] = str(dbg)
        for _ in range(2):
            _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), None, None, ws, stream))
        torch.cuda.synchronize()
        out = np.zeros(2 * 2048, dtype=np.uint64)
        lib.qcs_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
        k = lib.qcs_debug_tc_trace(out.ctypes.data, 2048)
        print(f'--- dbg {dbg}: {k} stamps')
        prev = None
        for i in range(k):
            tag, t = int(out[2 * i]), int(out[2 * i + 1])
            print(f'{TAGS.get(tag, tag):>20s}  +{(t - prev) if prev is not None else 0:6d}')
            prev = t


if __name__ == '__main__':
    main()
