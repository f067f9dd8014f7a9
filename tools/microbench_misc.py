"""Achieved HBM bandwidth of the streaming / reduction kernels of qcs_misc.cu (measurement, collapse, kron,
abs2, dot, lincomb, scale, permute, density-operator helpers) against their algorithmic bytes.
Run on the GPU box:  python tools/microbench_misc.py [--n 30]  -> gpurun_out/microbench_misc.json"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200 import _lib
    from qcircuits_b200.benchutil import measured_peak

    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    peak, _ = measured_peak()
    rows = []

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

    def row(dtype, what, ms, nbytes):
        r = {'dtype': np.dtype(dtype).name, 'what': what, 'ms': ms, 'algorithmic_bytes': nbytes,
             'GBs': nbytes / ms / 1e6, 'frac_of_measured_peak': nbytes / ms / 1e6 / peak}
        rows.append(r)
        print(r, flush=True)

    for dtype, n in ((np.complex64, args.n), (np.complex128, args.n - 2)):
        B = 8 if dtype == np.complex64 else 16
        N = 2 ** n
        s = qc.zeros(n, dtype=dtype)
        for q in (0, n // 2, n - 1):
            s = qc.Hadamard()(s, qubit_indices=[q])
        dv = s._dv
        # references: what a write-only and a copy stream reach on this box (torch fill / copy of the same buffer)
        scratch = torch.empty_like(dv.buf)
        row(dtype, 'reference: write-only stream (torch zero_)', timeit(lambda: scratch.zero_()), N * B)
        row(dtype, 'reference: copy stream (torch copy_)', timeit(lambda: scratch.copy_(dv.buf)), 2 * N * B)
        del scratch
        row(dtype, 'abs2 sum only (norm)', timeit(lambda: dv.norm2()), N * B)
        other = dv.clone()                      # a DIFFERENT buffer: <a|a> on one buffer reads half the bytes from L2
        row(dtype, 'dot (two distinct buffers)', timeit(lambda: dv.dot(other)), 2 * N * B)
        del other
        row(dtype, 'scale (renormalize_)', timeit(lambda: dv.scaled(0.5)), 2 * N * B)
        for m, bits in ((1, [n - 1]), (3, [n - 1, 5, 0]), (8, list(range(n - 1, n - 9, -1))), (10, list(range(9, -1, -1))),
                        (16, list(range(2, 18)))):
            row(dtype, 'marginal_probs m=%d' % m, timeit(lambda: dv.marginals(bits)), N * B)
        # keeping reads the kept half and writes everything (zeros elsewhere)
        row(dtype, 'collapse keep (1 qubit)', timeit(lambda: dv.collapsed([n - 1], 0, False)), N * B // 2 + N * B)
        # removal reads only the kept 1 / 2^m of the amplitudes and writes as many
        row(dtype, 'collapse remove (1 high qubit)', timeit(lambda: dv.collapsed([n - 1], 0, True)), N * B // 2 + N * B // 2)
        row(dtype, 'collapse remove (1 low qubit: every other amplitude)', timeit(lambda: dv.collapsed([0], 0, True)), N * B // 2 + N * B // 2)
        row(dtype, 'collapse remove (3 qubits)', timeit(lambda: dv.collapsed([n - 1, 7, 0], 5, True)), N * B // 8 + N * B // 8)
        perm = list(range(n))
        perm[0], perm[n - 1] = perm[n - 1], perm[0]
        row(dtype, 'permute_bits (swap bit 0 <-> top)', timeit(lambda: dv.permuted(perm)), 2 * N * B)
        perm3 = list(range(n))
        perm3[0], perm3[7] = perm3[7], perm3[0]
        row(dtype, 'permute_bits (swap bit 0 <-> bit 7)', timeit(lambda: dv.permuted(perm3)), 2 * N * B)
        perm2 = list(range(n))
        perm2[12], perm2[n - 1] = perm2[n - 1], perm2[12]
        row(dtype, 'permute_bits (swap bit 12 <-> top)', timeit(lambda: dv.permuted(perm2)), 2 * N * B)
        del s, dv
        torch.cuda.empty_cache()
        a = qc.zeros(n - 4, dtype=dtype)
        a = qc.Hadamard()(a, qubit_indices=[0])
        b = qc.zeros(4, dtype=dtype)
        b = qc.Hadamard()(b, qubit_indices=[1])
        row(dtype, 'kron (n-4) x 4', timeit(lambda: a._dv.kron(b._dv)), N * B + N * B // 16)
        row(dtype, 'kron 4 x (n-4)', timeit(lambda: b._dv.kron(a._dv)), N * B + N * B // 16)
        lc, lc2 = a._dv, a._dv.clone()          # two distinct operands
        row(dtype, 'lincomb (State + State, distinct operands), n-4 qubits', timeit(lambda: lc.lincomb(lc2, 1.0, 1.0)), 3 * (N // 16) * B)
        del a, b, lc, lc2
        torch.cuda.empty_cache()
        # density operator helpers on d = n // 2 qubits (rank 2d tensor)
        d = (n - 2) // 2
        psi = qc.zeros(d, dtype=dtype)
        psi = qc.Hadamard()(psi, qubit_indices=[0])
        rho = psi.density_operator()
        M = 4 ** d
        row(dtype, 'outer_accumulate (rho = psi psi^dagger), d=%d' % d, timeit(lambda: psi.density_operator()), M * B)
        row(dtype, 'diag_probs, d=%d' % d, timeit(lambda: rho.probabilities), 2 ** d * (B + 8))
        Hg = qc.Hadamard()
        row(dtype, 'U rho U^dagger (1-qubit gate, one fused pass), d=%d' % d, timeit(lambda: Hg(rho, qubit_indices=[1])), 2 * M * B)
        del psi, rho
        torch.cuda.empty_cache()
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'n': args.n, 'peak_GBs': peak, 'rows': rows}, open('gpurun_out/microbench_misc.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
