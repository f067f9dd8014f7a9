"""A small run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck), each checked
against the oracle so that the run also proves the sanitized code path is the real one:
tensor-core passes (tcgen05 + TMEM + cp.async ring), register-sweep passes of both amplitude types (cp.async ring and the
TMA bulk ring), the dense path, marginals / collapse / measure, kron, permute, abs2 / dot / lincomb, density operators
(outer product, U rho U^dagger, diagonal, partial trace)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import qcircuits_b200 as qc
    from oracle import qc_oracle as orc
    from qcircuits_b200 import _gates, _lib
    from qcircuits_b200._device import make_ops
    rng = np.random.default_rng(5)

    def runit(k):
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        Q, R = np.linalg.qr(A)
        return Q * (np.diag(R) / np.abs(np.diag(R)))

    def rvec(n):
        v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        return v / np.linalg.norm(v)

    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    CNOT = np.eye(4, dtype=np.complex128)[[0, 1, 3, 2]]
    CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
    worst = 0.0
    for dtype, tol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        for mode in (0, 1):
            _lib.load().qcs_set_tile_mode(mode)
            for n in (12, 14):
                vec = rvec(n)
                gates = []
                tb = 12 if dtype == np.complex64 else 11         # dense targets must fit one tile
                for _ in range(40):
                    kind = int(rng.integers(6))
                    if kind == 0:
                        gates.append((H, [int(rng.integers(tb))]))
                    elif kind == 1:
                        gates.append((runit(1), [int(rng.integers(tb))]))
                    elif kind == 2:
                        a = int(rng.integers(tb - 1))
                        gates.append((CNOT, [a, a + 1]))
                    elif kind == 3:
                        gates.append((CZ, [int(b) for b in rng.permutation(n)[:2]]))
                    elif kind == 4:
                        gates.append((runit(2), [int(b) for b in rng.permutation(tb)[:2]]))
                    else:
                        gates.append((runit(3), [int(b) for b in rng.permutation(tb)[:3]]))
                ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
                s = qc.State.from_column_vector(vec, dtype=dtype)
                got = s._dv.apply_ops(ops, keep).to_host()
                want = vec
                for M, bits in gates:
                    want = orc.apply_matrix_flat(M, want, bits, renorm=False)
                want /= np.linalg.norm(want)
                worst = max(worst, float(np.max(np.abs(got - want))) / tol)
        _lib.load().qcs_set_tile_mode(0)
        n = 10
        vec = rvec(n)
        s = qc.State.from_column_vector(vec, dtype=dtype)
        U = runit(6)
        got = qc.Operator.from_matrix(U)(s, qubit_indices=[5, 0, 9, 2, 7, 3]).to_column_vector()
        want = orc.apply_matrix_flat(U, vec, [n - 1 - q for q in (5, 0, 9, 2, 7, 3)])
        worst = max(worst, float(np.max(np.abs(got - want))) / tol)
        ps = s.marginal_probabilities([3, 0, 7])
        wantp, _, _ = orc.measurement_probabilities(vec.reshape([2] * n), [3, 0, 7])
        worst = max(worst, float(np.max(np.abs(ps - wantp))) / tol)
        for remove in (False, True):
            s2 = qc.State.from_column_vector(vec, dtype=dtype)
            bits = s2.measure([3, 0, 7], remove=remove, u=0.4321)
            bw, post = orc.measure(vec.reshape([2] * n), [3, 0, 7], 0.4321, remove=remove)
            assert tuple(bits) == tuple(bw)
            worst = max(worst, float(np.max(np.abs(s2.to_column_vector() - post.reshape(-1)))) / (4 * tol))
        a, b = qc.State.from_column_vector(rvec(4), dtype=dtype), qc.State.from_column_vector(rvec(7), dtype=dtype)
        worst = max(worst, float(np.max(np.abs((a * b).to_column_vector() - np.kron(a.to_column_vector(), b.to_column_vector())))) / tol)
        s3 = qc.State.from_column_vector(vec, dtype=dtype)
        s3.permute_qubits([3, 1, 0, 2, 9, 8, 7, 6, 5, 4])
        worst = max(worst, float(np.max(np.abs(s3._t - np.transpose(vec.reshape([2] * n), [3, 1, 0, 2, 9, 8, 7, 6, 5, 4])))) / tol)
        worst = max(worst, abs(s.dot(s) - 1.0) / tol, abs(float(s.probabilities.sum()) - 1.0) / (10 * tol))
        _ = (s + s3) - s
        rho = qc.DensityOperator.from_ensemble([qc.State.from_column_vector(rvec(5), dtype=dtype) for _ in range(2)], [0.3, 0.7])
        rho2 = qc.Operator.from_matrix(runit(2))(rho, qubit_indices=[4, 1])
        worst = max(worst, abs(float(rho2.probabilities.sum()) - 1.0) / (10 * tol))
        _ = rho2.reduced_density_operator([0, 3])
        rho2.measure([2], u=0.3)
    print('sanitize target done: worst error / tolerance %.3f' % worst)
    assert worst < 1.0


if __name__ == '__main__':
    main()
