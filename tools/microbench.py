"""Per-bit-position bandwidth of the gate pass, both staging variants, plus fused-run scaling.
Run on the GPU box:  python tools/microbench.py [--n 30] > gpurun_out/microbench.json"""
import argparse
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _lib
    from qcircuits_b200._device import make_ops

    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--modes', default='0,1')
    args = ap.parse_args()
    lib = _lib.load()
    res = {'n': args.n, 'rows': []}

    def timeit(fn, reps=args.reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for dtype, n in ((np.complex64, args.n), (np.complex128, args.n - 1)):
        amp = 8 if dtype == np.complex64 else 16
        bytes_pass = 2.0 * 2 ** n * amp
        s = qc.zeros(n, dtype=dtype)
        s = qc.Hadamard()(s, qubit_indices=[0])
        dst = s._dv._new_like()
        ms = timeit(lambda: dst.buf.copy_(s._dv.buf))
        res['rows'].append({'dtype': np.dtype(dtype).name, 'what': 'torch copy_', 'ms': ms, 'GBs': bytes_pass / ms / 1e6})
        print(res['rows'][-1], flush=True)
        H, RX, CN, CZ, RZ = qc.Hadamard(), qc.RotationX(0.3), qc.CNOT(), qc.ControlledU(qc.PauliZ()), qc.RotationZ(0.4)
        U2 = qc.Operator.from_matrix(np.linalg.qr(np.random.default_rng(0).normal(size=(4, 4)))[0])
        for mode in [int(m) for m in args.modes.split(',')]:
            lib.qcs_set_tile_mode(mode)
            for bit in (0, 1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 16, 22, n - 1):
                ops, keep = make_ops([(RX._plan(), [bit], False)])
                ms = timeit(lambda: s._dv.apply_ops(ops, keep, out=dst))
                res['rows'].append({'dtype': np.dtype(dtype).name, 'mode': mode, 'what': '1q bit %d' % bit, 'ms': ms,
                                    'GBs': bytes_pass / ms / 1e6})
                print(res['rows'][-1], flush=True)
            for name, g, bits in (('cnot hi-hi', CN, [n - 1, n - 2]), ('cnot lo-lo', CN, [1, 0]), ('cnot hi-lo', CN, [n - 1, 3]),
                                  ('cz hi-hi', CZ, [n - 1, n - 5]), ('rz hi', RZ, [n - 3]), ('dense2 hi-hi', U2, [n - 1, n - 2]),
                                  ('dense2 lo-lo', U2, [5, 2]), ('dense2 3hi', U2, [n - 1, 14])):
                ops, keep = make_ops([(g._plan(), bits, False)])
                ms = timeit(lambda: s._dv.apply_ops(ops, keep, out=dst))
                res['rows'].append({'dtype': np.dtype(dtype).name, 'mode': mode, 'what': name, 'ms': ms,
                                    'GBs': bytes_pass / ms / 1e6})
                print(res['rows'][-1], flush=True)
            # fused runs: g one-qubit dense gates on low bits + two high bits
            for g in (1, 2, 4, 8, 12, 16, 24, 32):
                bits_cycle = [0, 3, 5, 7, 9, n - 1, n - 2, 1, 2, 4, 6, 8]
                plans = [(RX._plan() if i % 2 else H._plan(), [bits_cycle[i % len(bits_cycle)]], False) for i in range(g)]
                ops, keep = make_ops(plans)
                ms = timeit(lambda: s._dv.apply_ops(ops, keep, out=dst))
                res['rows'].append({'dtype': np.dtype(dtype).name, 'mode': mode, 'what': 'fused %d x 1q' % g, 'ms': ms,
                                    'GBs': bytes_pass / ms / 1e6, 'gates_per_s': g / ms * 1e3})
                print(res['rows'][-1], flush=True)
            for g in (4, 16, 32):
                plans = [(RZ._plan(), [(7 * i) % n], False) for i in range(g)]
                ops, keep = make_ops(plans)
                ms = timeit(lambda: s._dv.apply_ops(ops, keep, out=dst))
                res['rows'].append({'dtype': np.dtype(dtype).name, 'mode': mode, 'what': 'fused %d x diag' % g, 'ms': ms,
                                    'GBs': bytes_pass / ms / 1e6, 'gates_per_s': g / ms * 1e3})
                print(res['rows'][-1], flush=True)
        del s, dst
        torch.cuda.empty_cache()
    lib.qcs_set_tile_mode(0)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/microbench.json', 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
