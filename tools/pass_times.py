"""GPU: time every fused pass of the headline circuit on its own (CUDA events, 3 launches each) next to
its plan statistics.  Writes gpurun_out/pass_times.json."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _lib
    from qcircuits_b200.workloads import layered_circuit
    n = int(os.environ.get('PT_N', '30'))
    max_ops = int(os.environ.get('PT_MAX_OPS', '128'))
    dtype = np.complex64 if os.environ.get('PT_DTYPE', 'c64') == 'c64' else np.complex128
    lib = _lib.load()
    circ = layered_circuit(n, 40, seed=1234)
    comp = circ._compiled(dtype, True, max_ops)
    s = qc.zeros(n, dtype=dtype)
    s = qc.Hadamard()(s, qubit_indices=[0])
    buf = s._dv.buf.data_ptr()
    unit = torch.ones(2, dtype=torch.float64, device='cuda')     # the state is normalised: its lazy squared norm (ranges the fp16 operands)
    norm_in = unit.data_ptr()
    ws, stream, code = _lib.workspace().data_ptr(), _lib.stream_ptr(), s._dv.code
    stats = (ctypes.c_int32 * 128)()
    out = []
    for idx, item in enumerate(comp):
        if item[0] != 'tile':
            continue
        _lib.check(lib.qcs_plan_stats(n, code, item[1], len(item[1]), stats))
        st = list(stats)
        _lib.check(lib.qcs_apply_fused(buf, buf, n, code, item[1], len(item[1]), norm_in, None, ws, stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _lib.check(lib.qcs_apply_fused(buf, buf, n, code, item[1], len(item[1]), norm_in, None, ws, stream))
        e1.record()
        torch.cuda.synchronize()
        out.append({'pass': idx, 'gates': item[3], 'ms': e0.elapsed_time(e1) / 3, 'tile_bits': st[0], 'sweeps': st[1],
                    'regops': st[2], 'tc_sweeps': st[7] & 255, 'block_ops': st[7] >> 8, 'generic': st[3], 'outer': st[4], 'scale_mode': st[5],
                    'arms': {str(a): st[8 + a] for a in range(100) if st[8 + a]}})
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/pass_times.json', 'w'))
    # synthetic passes: k dense blocks on disjoint 5-bit sets (tensor-core sweep cost), with and without register ops
    from qcircuits_b200 import _gates
    from qcircuits_b200._device import make_ops
    rng = np.random.default_rng(0)

    def runit(k):
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        Q, R = np.linalg.qr(A)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    synth = []
    if dtype == np.complex64:
        sets = [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [2, 3, 4, 5, 6], [7, 8, 9, 10, 11]]
        for nblocks in (1, 2, 3, 4):
            for extra in (0, 12):
                gates = []
                for bset in sets[:nblocks]:
                    gates += [(runit(2), [bset[0], bset[1]]), (runit(2), [bset[2], bset[3]]), (runit(2), [bset[3], bset[4]]),
                              (runit(2), [bset[1], bset[2]])]
                    # register ops of the same sweep: controlled phases reaching outside the block / tile
                    gates += [(np.diag([1, 1, 1, np.exp(0.3j)]), [bset[i % 4], 20 + i % 5]) for i in range(extra)]
                ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
                _lib.check(lib.qcs_plan_stats(n, code, ops, len(ops), stats))
                st = list(stats)
                _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), norm_in, None, ws, stream))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    _lib.check(lib.qcs_apply_fused(buf, buf, n, code, ops, len(ops), norm_in, None, ws, stream))
                e1.record()
                torch.cuda.synchronize()
                rec = {'blocks': nblocks, 'extra_regops': extra, 'ms': e0.elapsed_time(e1) / 5, 'sweeps': st[1], 'regops': st[2],
                       'tc_sweeps': st[7] & 255, 'block_ops': st[7] >> 8}
                synth.append(rec)
                print(rec)
    json.dump({'passes': out, 'synthetic': synth}, open('gpurun_out/pass_times.json', 'w'))
    for o in out:
        print({k: o[k] for k in ('gates', 'ms', 'sweeps', 'tc_sweeps', 'block_ops', 'regops')})
    ms = [o['ms'] for o in out]
    print('passes', len(out), 'total ms', sum(ms), 'min', min(ms), 'max', max(ms))


if __name__ == '__main__':
    main()
