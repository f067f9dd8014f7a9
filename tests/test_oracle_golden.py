"""Pin oracle/qc_oracle.py against outputs of the unmodified reference (tests/golden/golden_v1.npz,
made by tests/golden/make_golden.py) and against the known answers printed in the reference's
docs/source/tutorial.rst.  CPU only."""
import numpy as np
import pytest

from oracle import qc_oracle as orc

TOL = 1e-12


def _t(vec):
    n = int(round(np.log2(vec.size)))
    return np.asarray(vec).reshape([2] * n)


@pytest.mark.parametrize('n,depth', [(3, 4), (5, 6), (8, 6), (10, 8), (12, 4)])
def test_layered_circuit_matches_reference(golden, n, depth):
    gates = orc.layered_circuit(n, depth, seed=1234)
    want = golden['layered_n%d_d%d_amps' % (n, depth)]
    got = orc.run_circuit(orc.zeros_state(n), gates)
    assert np.max(np.abs(got.reshape(-1) - want)) < TOL
    got_flat = orc.run_circuit(orc.zeros_state(n), gates, flat=True)
    assert np.max(np.abs(got_flat.reshape(-1) - want)) < TOL
    assert np.max(np.abs(orc.probabilities(got).reshape(-1) - golden['layered_n%d_d%d_probs' % (n, depth)])) < TOL


def test_gate_count_formula():
    # n=28,d=40 -> 1660 gates; n=30,d=40 -> 1780; n=16 -> 940; n=20 -> 1180
    assert len(orc.layered_circuit(28, 40)) == 1660
    assert len(orc.layered_circuit(30, 40)) == 1780
    assert len(orc.layered_circuit(16, 40)) == 940
    assert len(orc.layered_circuit(20, 40)) == 1180


def test_apply_random_unitaries(golden):
    for c in range(int(golden['apply_count'])):
        vin, M = golden['apply%d_in' % c], golden['apply%d_op' % c]
        qi, want = [int(q) for q in golden['apply%d_qi' % c]], golden['apply%d_out' % c]
        n = int(round(np.log2(vin.size)))
        got = orc.apply_operator(orc.op_from_matrix(M), _t(vin), qi)
        assert np.max(np.abs(got.reshape(-1) - want)) < TOL, c
        got2 = orc.apply_matrix_flat(M, vin, [n - 1 - q for q in qi])
        assert np.max(np.abs(got2 - want)) < TOL, c


def test_non_unitary_is_renormalised(golden):
    vin = golden['proj_in']
    P0 = orc.op_from_matrix(np.array([[1, 0], [0, 0]]))
    got = orc.apply_operator(P0, _t(vin), [2])
    assert np.max(np.abs(got.reshape(-1) - golden['proj_out'])) < TOL
    got = orc.apply_operator(orc.op_from_matrix(golden['nonunitary_op']), _t(vin), [4, 1])
    assert np.max(np.abs(got.reshape(-1) - golden['nonunitary_out'])) < TOL
    got = orc.apply_matrix_flat(golden['nonunitary_op'], vin, [0, 3])
    assert np.max(np.abs(got - golden['nonunitary_out'])) < TOL


def test_measurement_matches_reference_given_same_uniform(golden):
    for c in range(int(golden['meas_count'])):
        vin = golden['meas%d_in' % c]
        qi = golden['meas%d_qi' % c]
        n = int(round(np.log2(vin.size)))
        int_arg = qi.ndim == 0 and int(qi) >= 0
        if qi.ndim == 0:
            qi = list(range(n)) if int(qi) < 0 else [int(qi)]
        else:
            qi = [int(q) for q in qi]
        remove = bool(golden['meas%d_remove' % c])
        bits, new_t = orc.measure(_t(vin), qi, float(golden['meas%d_u' % c]), remove=remove)
        want_bits = golden['meas%d_bits' % c]
        if int_arg:
            assert bits[0] == int(want_bits)
        else:
            assert tuple(bits) == tuple(int(b) for b in want_bits), c
        assert np.max(np.abs(np.asarray(new_t).reshape(-1) - golden['meas%d_out' % c].reshape(-1))) < TOL, c


def test_sample_outcome_is_numpy_choice():
    rng = np.random.RandomState(3)
    for _ in range(200):
        p = rng.dirichlet(np.ones(8))
        seed = int(rng.randint(1 << 30))
        np.random.seed(seed)
        u = np.random.random_sample()
        np.random.seed(seed)
        assert orc.sample_outcome(p, u) == np.random.choice(8, p=p)


def test_tensor_product_and_arithmetic(golden):
    a, b = _t(golden['kron_a']), _t(golden['kron_b'])
    assert np.max(np.abs(orc.tensor_product(a, b).reshape(-1) - golden['kron_ab'])) < TOL
    p3 = orc.tensor_product(orc.tensor_product(b, b), b)
    assert np.max(np.abs(p3.reshape(-1) - golden['kron_pow3'])) < TOL
    assert abs(orc.dot(a, a[::-1]) - golden['dot_ab']) < TOL


def test_density_operator(golden):
    states = [_t(v) for v in golden['rho_states']]
    rho = orc.from_ensemble(states, golden['rho_ps'])
    assert np.max(np.abs(rho.reshape(-1) - golden['rho_t'])) < TOL
    assert np.max(np.abs(orc.density_probabilities(rho).reshape(-1) - golden['rho_probs'])) < TOL
    U2 = orc.op_from_matrix(golden['rho_op'])
    rho2 = orc.apply_operator_density(U2, rho, [2, 0])
    assert np.max(np.abs(rho2.reshape(-1) - golden['rho_applied'])) < TOL
    rho_h = orc.apply_operator_density(orc.hadamard(), rho, [1])
    assert np.max(np.abs(rho_h.reshape(-1) - golden['rho_h1'])) < TOL
    red = orc.reduced_tensor(rho, [0, 2])
    assert np.max(np.abs(red.reshape(-1) - golden['rho_reduced_02'])) < TOL
    for c in range(int(golden['rhomeas_count'])):
        qi = golden['rhomeas%d_qi' % c]
        qi = list(range(3)) if qi.ndim == 0 else [int(q) for q in qi]
        bits, new_t = orc.density_measure(rho2, qi, float(golden['rhomeas%d_u' % c]),
                                          remove=bool(golden['rhomeas%d_remove' % c]))
        assert tuple(bits) == tuple(int(b) for b in golden['rhomeas%d_bits' % c])
        assert np.max(np.abs(np.asarray(new_t).reshape(-1) - golden['rhomeas%d_out' % c])) < TOL


def test_density_is_two_gate_passes_on_doubled_bits(golden):
    """U rho U^dagger == U on bits 2(n-1-q)+1 then conj(U) on bits 2(n-1-q) of the
    flat rank-2n tensor.  This identity is what the CUDA DensityOperator path relies on."""
    n = 3
    rho_flat = golden['rho_t']
    M = golden['rho_op']
    qi = [2, 0]
    v = orc.apply_matrix_flat(M, rho_flat, [2 * (n - 1 - q) + 1 for q in qi], renorm=False)
    v = orc.apply_matrix_flat(np.conj(M), v, [2 * (n - 1 - q) for q in qi], renorm=False)
    assert np.max(np.abs(v - golden['rho_applied'])) < TOL


def test_compose_and_grover(golden):
    A, B = orc.op_from_matrix(golden['compose_a']), orc.op_from_matrix(golden['compose_b'])
    AB = orc.compose_operators(A, B, [2, 0])
    assert np.max(np.abs(orc.op_to_matrix(AB) - golden['compose_ab'])) < TOL
    d = 5
    G = orc.op_from_matrix(golden['grover_d5_op'])
    state = orc.tensor_product(orc.zeros_state(d), np.array([0, 1], dtype=complex))
    Hall = orc.hadamard(d + 1)
    state = orc.apply_operator(Hall, state)
    for _ in range(4):
        state = orc.apply_operator(G, state)
    assert np.max(np.abs(state.reshape(-1) - golden['grover_d5_amps'])) < 1e-11
    bits, _ = orc.measure(state, list(range(d)), float(golden['grover_d5_u']))
    assert tuple(bits) == tuple(int(b) for b in golden['grover_d5_bits'])
    answers = golden['grover_d5_answers']
    assert answers[sum(v * 2 ** i for i, v in enumerate(bits))] == 1


def test_gate_tensors(golden):
    pairs = {'I': orc.identity(), 'X': orc.pauli_x(), 'Y': orc.pauli_y(), 'Z': orc.pauli_z(),
             'H': orc.hadamard(), 'H3': orc.hadamard(3), 'S': orc.phase(), 'T': orc.pi_by_8(),
             'RX': orc.rotation_x(0.7), 'RY': orc.rotation_y(1.3), 'RZ': orc.rotation_z(2.1),
             'R': orc.rotation([0.6, 0.0, 0.8], 0.9), 'CNOT': orc.cnot(), 'CZ': orc.cz(), 'Swap': orc.swap(),
             'CCH': orc.controlled(orc.controlled(orc.hadamard()))}
    for name, t in pairs.items():
        assert np.array_equal(t.reshape(-1), golden['gate_' + name]), name


def test_tutorial_known_answers():
    """docs/source/tutorial.rst:184-211 (X/Y/Z on basis states), :223-228 (H|0>), :260-279 (CNOT
    truth table), :505-518 ((H(x)I)|00>, (I(x)H)|00>)."""
    z, o = np.array([1, 0], dtype=complex), np.array([0, 1], dtype=complex)
    assert np.allclose(orc.apply_operator(orc.pauli_x(), z), o)
    assert np.allclose(orc.apply_operator(orc.pauli_y(), z), 1j * o)
    assert np.allclose(orc.apply_operator(orc.pauli_y(), o), -1j * z)
    assert np.allclose(orc.apply_operator(orc.pauli_z(), o), -o)
    assert np.allclose(orc.apply_operator(orc.hadamard(), z), np.array([1, 1]) / np.sqrt(2))
    for a in (0, 1):
        for b in (0, 1):
            s = np.zeros((2, 2), dtype=complex)
            s[a, b] = 1
            r = orc.apply_operator(orc.cnot(), s)
            want = np.zeros((2, 2), dtype=complex)
            want[a, b ^ a] = 1
            assert np.allclose(r, want)
    s00 = orc.zeros_state(2)
    assert np.allclose(orc.apply_operator(orc.hadamard(), s00, [0]).reshape(-1), np.array([1, 0, 1, 0]) / np.sqrt(2))
    assert np.allclose(orc.apply_operator(orc.hadamard(), s00, [1]).reshape(-1), np.array([1, 1, 0, 0]) / np.sqrt(2))
