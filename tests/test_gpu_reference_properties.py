"""The property tests of the reference's own suite (reference tests/tests.py), restated with seeded
generators and run on the CUDA engine: every identity below is pinned there with unseeded random
inputs (class names in the docstrings).  Tolerance 1e-10 as in the reference (tests.py:14)."""
import copy

import numpy as np
import pytest

import qcircuits_b200 as qc
from tests.util import rand_state_vec, rand_unitary

pytestmark = pytest.mark.gpu
EPS = 1e-10


def diff(a, b):
    return float(np.max(np.abs(np.asarray(a[:]) - np.asarray(b[:]))))


def rstate(rng, d):
    return qc.State.from_column_vector(rand_state_vec(rng, d))


def rop(rng, d):
    return qc.Operator.from_matrix(rand_unitary(rng, d))


def test_state_unit_length():
    """StateUnitLengthTests (tests.py:70-129)."""
    rng = np.random.default_rng(1)
    for d in range(1, 8):
        for s in (qc.zeros(d), qc.ones(d), qc.positive_superposition(d), qc.bitstring(*rng.integers(0, 2, d))):
            assert abs(s.probabilities.sum() - 1) < EPS
    for a in (0, 1):
        for b in (0, 1):
            assert abs(qc.bell_state(a, b).probabilities.sum() - 1) < EPS
    x = rstate(rng, 3) * rstate(rng, 2) * qc.qubit(theta=0.3, phi=1.1)
    assert x.rank == 6 and abs(x.probabilities.sum() - 1) < EPS


def test_operator_on_subsystem_equals_permute_route():
    """StateSwapPermuteTests (tests.py:189-228): U(x, qubit_indices=p) == permute -> U on leading qubits -> inverse."""
    rng = np.random.default_rng(2)
    for _ in range(10):
        d = int(rng.integers(3, 8))
        k = int(rng.integers(1, d))
        perm = [int(p) for p in rng.permutation(d)]
        U, x = rop(rng, k), rstate(rng, d)
        direct = U(x, qubit_indices=perm[:k])
        y = copy.deepcopy(x).permute_qubits(perm)
        y = U(y, qubit_indices=range(k))
        y.permute_qubits(perm, inverse=True)
        assert diff(direct, y) < EPS
    x = rstate(rng, 5)
    y = copy.deepcopy(x)
    y.swap_qubits(1, 3)
    U = rop(rng, 1)
    y = U(y, qubit_indices=[1])
    y.swap_qubits(1, 3)
    assert diff(U(x, qubit_indices=[3]), y) < EPS


def test_tensor_product_of_operators_and_states():
    """TensorProductTests (tests.py:289-305): (A*B)(x*y) == A(x)*B(y)."""
    rng = np.random.default_rng(3)
    for _ in range(10):
        d1, d2 = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        A, B, x, y = rop(rng, d1), rop(rng, d2), rstate(rng, d1), rstate(rng, d2)
        assert diff((A * B)(x * y), A(x) * B(y)) < EPS


def test_apply_to_subsystem_equals_identity_padding_and_arguments_unchanged():
    """OperatorApplyToSubSystemTest (tests.py:308-385)."""
    rng = np.random.default_rng(4)
    for _ in range(10):
        d = int(rng.integers(2, 7))
        k = int(rng.integers(1, d + 1))
        start = int(rng.integers(0, d - k + 1))
        U, x = rop(rng, k), rstate(rng, d)
        x_before, U_before = x.to_column_vector(), U.to_matrix().copy()
        padded = U
        if start:
            padded = qc.Identity(start) * padded
        if d - k - start:
            padded = padded * qc.Identity(d - k - start)
        qi = list(range(start, start + k))
        assert diff(U(x, qubit_indices=qi), padded(x)) < EPS
        assert diff(U(x, qubit_indices=qi), U(qc.Identity(d), qubit_indices=qi)(x)) < EPS
        assert np.array_equal(x.to_column_vector(), x_before) and np.array_equal(U.to_matrix(), U_before)
        qi = [int(q) for q in rng.choice(d, size=k, replace=False)]       # ndarray-style indices too
        assert diff(U(x, qubit_indices=np.array(qi)), U(qc.Identity(d), qubit_indices=qi)(x)) < EPS


def test_operator_composition():
    """OperatorCompositionTests (tests.py:708-751)."""
    rng = np.random.default_rng(5)
    for _ in range(6):
        d = int(rng.integers(1, 6))
        U1, U2, U3, x = rop(rng, d), rop(rng, d), rop(rng, d), rstate(rng, d)
        a = U1(U2(U3(x)))
        assert diff(a, (U1(U2))(U3(x))) < EPS
        assert diff(a, (U1(U2(U3)))(x)) < EPS
        assert diff(U1(U1.adj(x)), x) < EPS


def test_product_of_one_qubit_gates_on_a_subset():
    """ApplyingToSubsetsOfQubitsTests (tests.py:754-800)."""
    rng = np.random.default_rng(6)
    d = 6
    x = rstate(rng, d)
    gates = [rop(rng, 1) for _ in range(3)]
    qi = [4, 0, 2]
    y = x
    for g, q in zip(gates, qi):
        y = g(y, qubit_indices=[q])
    full = gates[0] * gates[1] * gates[2]
    assert diff(y, full(x, qubit_indices=qi)) < EPS


def test_kronecker_convention_against_dense_matrix():
    """KroneckerProductTests (tests.py:803-872): Op(s) == Op.to_matrix() @ s.to_column_vector()."""
    rng = np.random.default_rng(7)
    for d in range(1, 8):
        U, x = rop(rng, d), rstate(rng, d)
        assert diff(U(x).to_column_vector(), U.to_matrix() @ x.to_column_vector()) < EPS
    s = qc.bitstring(1, 0, 1)
    assert np.array_equal(s.to_column_vector(), np.eye(8)[5])
    assert np.allclose((qc.PauliX() * qc.Identity()).to_matrix(), np.kron([[0, 1], [1, 0]], np.eye(2)))


def test_schmidt_number_on_the_device_matches_a_host_svd():
    """state.py:225-260 -- random states built as sums of r product terms across a cut have Schmidt number r"""
    rng = np.random.default_rng(3)
    for d, cut, r in ((6, [0, 1, 2], 3), (7, [1, 4], 2), (8, [0, 2, 5, 7], 5), (5, [4], 1), (9, [3, 8, 0], 8)):
        a, b = len(cut), d - len(cut)
        rest = [q for q in range(d) if q not in cut]
        M = np.zeros((2 ** a, 2 ** b), dtype=np.complex128)
        for _ in range(r):
            M += np.outer(rng.normal(size=2 ** a) + 1j * rng.normal(size=2 ** a),
                          rng.normal(size=2 ** b) + 1j * rng.normal(size=2 ** b))
        M /= np.linalg.norm(M)
        # tensor with the cut qubits first, then moved to their places
        tens = np.moveaxis(M.reshape([2] * d), list(range(d)), cut + rest)
        s = qc.State(tens)
        want = int(np.sum(np.linalg.svd(M, compute_uv=False) > 1e-10))
        assert want == min(r, 2 ** a, 2 ** b)
        assert s.schmidt_number(cut) == want


def test_schmidt_number_and_arithmetic():
    """SchmidtTests / addition and scalar multiplication tests (tests.py:652-705, 875-924)."""
    assert qc.bell_state(0, 0).schmidt_number([0]) == 2
    assert (qc.zeros(1) * qc.ones(1)).schmidt_number([1]) == 1
    with pytest.raises(ValueError):
        qc.bell_state().schmidt_number([])
    s = (qc.zeros(2) + qc.ones(2)) / np.sqrt(2)
    assert diff(s, qc.bell_state(0, 0)) < EPS
    assert diff(-(-s), s) < EPS and diff(s - s, np.zeros((2, 2))) < EPS


def test_density_operator_is_the_ensemble_of_states():
    """DensityOperatorTests (tests.py:958-1054): Op(rho) == ensemble of Op(state); marginals are mixtures."""
    rng = np.random.default_rng(8)
    for _ in range(4):
        d = int(rng.integers(2, 5))
        states = [rstate(rng, d) for _ in range(3)]
        ps = rng.dirichlet(np.ones(3))
        rho = qc.DensityOperator.from_ensemble(states, ps)
        k = int(rng.integers(1, d + 1))
        qi = [int(q) for q in rng.permutation(d)[:k]]
        U = rop(rng, k)
        want = qc.DensityOperator.from_ensemble([U(s, qubit_indices=qi) for s in states], ps)
        np.testing.assert_allclose(U(rho, qubit_indices=qi)._t, want._t, atol=1e-12)
        mq = [int(q) for q in rng.permutation(d)[:2]]
        mix = sum(p * s.marginal_probabilities(mq) for p, s in zip(ps, states))
        np.testing.assert_allclose(rho._measurement_probabilities(mq)[0], mix, atol=1e-12)
        r2 = copy.deepcopy(rho)
        r2.permute_qubits(list(range(d))[::-1]).permute_qubits(list(range(d))[::-1], inverse=True)
        np.testing.assert_allclose(r2._t, rho._t, atol=1e-13)
    bell = qc.bell_state(0, 0)
    np.testing.assert_allclose(bell.reduced_density_operator([0])._t, np.diag([0.5, 0.5]), atol=1e-12)


def test_measurement_behaviour():
    """FastMeasurementTests (tests.py:1524-1590): basis states are deterministic, repeated measurement
    is idempotent, U_f computes f."""
    rng = np.random.default_rng(9)
    for _ in range(10):
        d = int(rng.integers(1, 8))
        bits = tuple(int(b) for b in rng.integers(0, 2, d))
        s = qc.bitstring(*bits)
        assert s.measure() == bits
        q = int(rng.integers(d))
        assert s.measure(q) == bits[q]
    s = rstate(rng, 6)
    first = s.measure([4, 1])
    for _ in range(3):
        assert s.measure([4, 1]) == first
    rho = rstate(rng, 3).density_operator()
    first = rho.measure([2, 0])
    assert rho.measure([2, 0]) == first

    def f(a, b):
        return a ^ b
    Uf = qc.U_f(f, 3)
    for a in (0, 1):
        for b in (0, 1):
            assert Uf(qc.bitstring(a, b, 0)).measure() == (a, b, f(a, b))


def test_example_protocols():
    """TestExamples (tests.py:1607-1697): Bell pair, superdense coding, teleportation, Deutsch."""
    rng = np.random.default_rng(10)
    # superdense coding: two classical bits through one qubit of a Bell pair
    for b1 in (0, 1):
        for b2 in (0, 1):
            phi = qc.bell_state(0, 0)
            if b2:
                phi = qc.PauliX()(phi, qubit_indices=[0])
            if b1:
                phi = qc.PauliZ()(phi, qubit_indices=[0])
            phi = qc.CNOT()(phi)
            phi = qc.Hadamard()(phi, qubit_indices=[0])
            assert phi.measure() == (b1, b2)
    # teleportation of a random qubit
    for _ in range(5):
        alice_v = rand_state_vec(rng, 1)
        alice = qc.State.from_column_vector(alice_v)
        phi = alice * qc.bell_state(0, 0)
        phi = qc.CNOT()(phi, qubit_indices=[0, 1])
        phi = qc.Hadamard()(phi, qubit_indices=[0])
        m1, m2 = phi.measure(qubit_indices=[0, 1], remove=True)
        if m2:
            phi = qc.PauliX()(phi)
        if m1:
            phi = qc.PauliZ()(phi)
        assert diff(phi.to_column_vector(), alice_v) < EPS
    # Deutsch: constant vs balanced in one query
    for f, balanced in ((lambda x: 0, 0), (lambda x: 1, 0), (lambda x: x, 1), (lambda x: 1 - x, 1)):
        phi = qc.Hadamard(2)(qc.bitstring(0, 1))
        phi = qc.U_f(f, 2)(phi)
        phi = qc.Hadamard()(phi, qubit_indices=[0])
        assert phi.measure(0) == balanced


def test_measurement_statistics_small_sample():
    """slow_measurement_tests.py, at a reduced sample count: empirical frequencies follow probabilities."""
    rng = np.random.default_rng(11)
    np.random.seed(12)
    vec = rand_state_vec(rng, 2)
    counts = np.zeros(4)
    trials = 4000
    for _ in range(trials):
        s = qc.State.from_column_vector(vec)
        b = s.measure()
        counts[2 * b[0] + b[1]] += 1
    p = np.abs(vec) ** 2
    assert np.max(np.abs(counts / trials - p)) < 5 * np.sqrt(0.25 / trials)
