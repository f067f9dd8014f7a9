"""BASELINE.json configs at (or near) their full sizes, checked through properties that do not need the
oracle to hold the state: config 1 (Grover, 10+1 qubits), config 3 (phase estimation / inverse QFT with
fused diagonal gates, complex64), config 5 (DensityOperator U rho U^dagger, 14 qubits = rank-28 tensor).
Configs 2 and the 30-qubit headline are covered by test_gpu_parity.py::test_full_size_properties."""
import numpy as np
import pytest

import qcircuits_b200 as qc
from qcircuits_b200.circuit import Circuit

pytestmark = pytest.mark.gpu


def test_config1_grover_d10_finds_the_marked_item():
    """examples/grover_algorithm.py at d = 10: composed 11-qubit operators on the host, 25 applications of
    the Grover iterate through the dense path (k = n = 11), measurement of the first 10 qubits."""
    d = 10
    rng = np.random.default_rng(0)
    answers = np.zeros(2 ** d, dtype=np.int32)
    marked = int(rng.integers(2 ** d))
    answers[marked] = 1

    def f(*bits):
        return answers[sum(v * 2 ** i for i, v in enumerate(bits))]

    Oracle = qc.U_f(f, d=d + 1)
    H_d = qc.Hadamard(d)
    N = 2 ** d
    proj = np.zeros((N, N))
    proj[0, 0] = 1
    Inversion = H_d((2 * qc.Operator.from_matrix(proj) - qc.Identity(d))(H_d))
    Grover = Inversion(Oracle, qubit_indices=range(d))
    state = qc.zeros(d) * qc.ones(1)
    state = (H_d * qc.Hadamard())(state)
    iterations = int(round(np.arccos(np.sqrt(1 / N)) / (2 * np.arcsin(np.sqrt(1 / N)))))
    assert iterations == 25
    for _ in range(iterations):
        state = Grover(state)
    ps = state.marginal_probabilities(list(range(d)))
    want = sum(((marked >> i) & 1) << (d - 1 - i) for i in range(d))      # f's bit order: bit i of the index = qubit i
    assert ps[want] > 0.99
    np.random.seed(3)
    bits = state.measure(qubit_indices=range(d))
    assert f(*bits) == 1


@pytest.mark.parametrize('t,dtype', [(12, np.complex128), (26, np.complex64)])
def test_config3_phase_estimation_gate_by_gate(t, dtype):
    """Phase estimation applied gate by gate: H on the t-qubit register, controlled powers
    of a one-qubit phase gate, inverse QFT as H / controlled-phase / swap gates, all through the fusion
    scheduler.  For an eigenphase with an exact t-bit expansion the register ends in one basis state."""
    n = t + 1
    phi_bits = [1, 0, 1, 1, 0, 1] + [0] * (t - 6)
    phi = sum(b * 2.0 ** -(i + 1) for i, b in enumerate(phi_bits))
    c = Circuit(n)
    H = qc.Hadamard()
    for q in range(t):
        c.append(H, [q])
    # U = diag(1, e^{2 pi i phi}); the eigenvector |1> sits on qubit t; register qubit q controls U^(2^(t-1-q))
    for q in range(t):
        ang = 2 * np.pi * phi * (2 ** (t - 1 - q))
        U = qc.Operator.from_matrix(np.array([[1, 0], [0, np.exp(1j * (ang % (2 * np.pi)))]]))
        c.append(qc.ControlledU(U), [q, t])
    # inverse QFT on qubits 0..t-1: swaps, then (reversed QFT) controlled R_k^dagger and H
    for i in range(t // 2):
        c.append(qc.Swap(), [i, t - 1 - i])
    for q in range(t - 1, -1, -1):
        for k in range(t - q, 1, -1):
            Rk = qc.Operator.from_matrix(np.array([[1, 0], [0, np.exp(-2j * np.pi / 2 ** k)]]))
            c.append(qc.ControlledU(Rk), [q + k - 1, q])
        c.append(H, [q])
    state = qc.zeros(t, dtype=dtype) * qc.ones(1, dtype=dtype)
    out = c.run(state)
    passes, gates = c.stats(dtype)
    assert passes < gates / 3                       # the controlled phases ride along in fused passes
    qi = list(range(min(t, 10)))
    ps = out.marginal_probabilities(qi)
    want = int(''.join(str(b) for b in phi_bits[:len(qi)]), 2)
    tol = 1e-3 if dtype == np.complex64 else 1e-9
    assert ps[want] > 1 - tol
    rest = list(range(len(qi), t))
    if rest:
        ps2 = out.marginal_probabilities(rest[:10])
        assert ps2[int(''.join(str(b) for b in phi_bits[len(qi):len(qi) + len(rest[:10])]), 2)] > 1 - tol


def test_config5_density_operator_14_qubits():
    """rho = |0..0><0..0| on 14 qubits (rank-28 tensor, 4 GiB complex128); a layered circuit applied as
    U rho U^dagger must keep trace 1 and reproduce the probabilities of the pure-state evolution."""
    from qcircuits_b200.workloads import layered_gate_list, operator_for
    n = 14
    gates = layered_gate_list(n, 2, seed=21)
    rho = qc.zeros(n).density_operator()
    psi = qc.zeros(n)
    for name, theta, qs in gates:
        op = operator_for(name, theta)
        rho = op(rho, qubit_indices=list(qs))
        psi = op(psi, qubit_indices=list(qs))
    p_rho = rho.probabilities
    p_psi = psi.probabilities
    assert abs(p_rho.sum() - 1) < 1e-10
    assert np.max(np.abs(p_rho - p_psi)) < 1e-12
    ps, _ = rho._measurement_probabilities([13, 0, 6])
    assert np.max(np.abs(ps - psi.marginal_probabilities([13, 0, 6]))) < 1e-12
    u = 0.61
    assert rho.measure([13, 0, 6], u=u) == psi.measure([13, 0, 6], u=u)
    assert np.max(np.abs(rho.probabilities - psi.probabilities)) < 1e-12
