"""CPU experiment for the round-2 tensor-core question: does a 3xTF32 split hold the complex64 tolerance
(1e-5 relative to the largest amplitude) over the depth-40 layered circuit?

Every gate of the circuit is applied as a real matrix-vector product whose multiplications are done the way
a TF32 MMA does them -- operands rounded to 10 mantissa bits, products exact, FP32 accumulation -- once with
single TF32 operands and once with the 3-term split a = hi + lo (hi*hi + hi*lo + lo*hi).  Gate by gate is the
pessimistic count of roundings: a dense 4-qubit block applies ~12 gates in one product."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qc_oracle as orc  # noqa: E402  (a measurement tool, not the product path)


def tf32(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def matvec(M, X, mode):
    """rows of X are vectors (n_vec, k); returns X @ M.T with the chosen arithmetic, float32 accumulation"""
    M = M.astype(np.float32)
    X = X.astype(np.float32)
    out = np.zeros((X.shape[0], M.shape[0]), dtype=np.float32)
    if mode == 'fp32':
        for k in range(M.shape[1]):
            out += X[:, k:k + 1] * M[:, k][None, :]
        return out
    Mh, Xh = tf32(M), tf32(X)
    Ml, Xl = tf32(M - Mh), tf32(X - Xh)
    for k in range(M.shape[1]):
        out += Xh[:, k:k + 1] * Mh[:, k][None, :]
        if mode == '3xtf32':
            out += Xh[:, k:k + 1] * Ml[:, k][None, :]
            out += Xl[:, k:k + 1] * Mh[:, k][None, :]
    return out


def real_form(U):
    k = U.shape[0]
    R = np.zeros((2 * k, 2 * k))
    R[0::2, 0::2] = U.real; R[0::2, 1::2] = -U.imag
    R[1::2, 0::2] = U.imag; R[1::2, 1::2] = U.real
    return R


def run(n, depth, mode):
    gates = orc.layered_circuit(n, depth, seed=1234)
    psi = np.zeros(2 ** n, dtype=np.complex64)
    psi[0] = 1
    for name, theta, qs in gates:
        U = orc.op_to_matrix(orc.gate_tensor(name, theta))
        k = len(qs)
        t = psi.reshape([2] * n)
        t = np.moveaxis(t, list(qs), list(range(n - k, n)))
        shape = t.shape
        v = np.ascontiguousarray(t).reshape(-1, 2 ** k)
        X = np.empty((v.shape[0], 2 ** (k + 1)), dtype=np.float32)
        X[:, 0::2] = v.real; X[:, 1::2] = v.imag
        Y = matvec(real_form(U), X, mode)
        v = (Y[:, 0::2] + 1j * Y[:, 1::2]).astype(np.complex64)
        psi = np.moveaxis(v.reshape(shape), list(range(n - k, n)), list(qs)).reshape(-1)
    return psi


def main():
    n, depth = 12, 40
    ref = orc.run_circuit(orc.zeros_state(n), orc.layered_circuit(n, depth, seed=1234)).reshape(-1)
    for mode in ('fp32', '3xtf32', 'tf32'):
        got = run(n, depth, mode)
        got = got / np.linalg.norm(got)
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        print('%-7s max |error| / max |amplitude| = %.2e after %d gates' % (mode, err, len(orc.layered_circuit(n, depth, seed=1234))))


if __name__ == '__main__':
    main()
