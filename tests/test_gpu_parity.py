"""GPU parity tests: the CUDA path (through the Python drop-in API and the C ABI of libqcs.so)
against the CPU oracle on the same seeded inputs, and against golden outputs of the reference.
Tolerances (BASELINE.json north_star): complex128 1e-12, complex64 1e-5, relative to the largest
amplitude; measurement outcomes identical for identical uniform draws."""
import copy
import ctypes
import os

import numpy as np
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import _lib
from qcircuits_b200._device import make_ops
from tests.util import check_close, TOL, oracle_apply, rand_state_vec, rand_unitary, rel_err

pytestmark = pytest.mark.gpu

DTYPES = [np.complex128, np.complex64]
# 0 = cp.async ring (padded layout), 1 = TMA bulk-copy ring (QCS_TEST_MODES narrows the set, e.g. "0")
MODES = [int(m) for m in os.environ.get('QCS_TEST_MODES', '0,1').split(',')]


@pytest.fixture(autouse=True)
def _default_mode():
    yield
    _lib.load().qcs_set_tile_mode(0)


def _state(vec, dtype):
    return qc.State.from_column_vector(np.asarray(vec), dtype=dtype)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_one_qubit_gate_on_every_bit_position(dtype, mode):
    _lib.load().qcs_set_tile_mode(mode)
    rng = np.random.default_rng(10 + mode)
    for n in (1, 2, 3, 5, 9, 10, 11, 12, 13, 14, 17):
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        for q in range(n):
            U = rand_unitary(rng, 1)
            got = qc.Operator.from_matrix(U)(s, qubit_indices=[q])
            want = oracle_apply(U, vec, [q])
            check_close(got.to_column_vector(), want, dtype, 'test_one_qubit_gate_on_every_bit_position')
        # the argument is left untouched (tests.py:366-385)
        check_close(s.to_column_vector(), vec, dtype, 'test_one_qubit_gate_on_every_bit_position')


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_multi_qubit_gates_random_positions(dtype, mode):
    _lib.load().qcs_set_tile_mode(mode)
    rng = np.random.default_rng(20 + mode)
    for n in (2, 3, 4, 6, 10, 12, 15, 18):
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        for k in range(1, min(n, 4) + 1):
            for _ in range(4):
                qi = [int(q) for q in rng.permutation(n)[:k]]
                U = rand_unitary(rng, k)
                got = qc.Operator.from_matrix(U)(s, qubit_indices=qi)
                want = oracle_apply(U, vec, qi)
                check_close(got.to_column_vector(), want, dtype, 'test_multi_qubit_gates_random_positions')


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_large_state_high_and_low_bits(dtype, mode):
    """Many tiles per CTA, tile bits far above the chunk: n = 21."""
    _lib.load().qcs_set_tile_mode(mode)
    rng = np.random.default_rng(30)
    n = 21
    vec = rand_state_vec(rng, n)
    s = _state(vec, dtype)
    for qi in ([0], [20], [10], [11], [0, 20], [20, 0], [5, 12, 19], [0, 1, 2], [20, 19, 18, 17], [3, 20, 9, 14]):
        U = rand_unitary(rng, len(qi))
        got = qc.Operator.from_matrix(U)(s, qubit_indices=qi)
        want = oracle_apply(U, vec, qi)
        check_close(got.to_column_vector(), want, dtype, 'test_large_state_high_and_low_bits')


@pytest.mark.parametrize('dtype', DTYPES)
def test_structured_gates(dtype):
    """Controlled / diagonal structure is recognised on the host and executed as predicates."""
    rng = np.random.default_rng(40)
    n = 13
    vec = rand_state_vec(rng, n)
    s = _state(vec, dtype)
    Rk = np.exp(0.2j) * qc.RotationZ(0.9)
    gates = [qc.CNOT(), qc.ControlledU(qc.PauliZ()), qc.Toffoli(), qc.Swap(), qc.SqrtSwap(), qc.RotationZ(1.1),
             qc.ControlledU(Rk), qc.ControlledU(qc.ControlledU(qc.Hadamard())), qc.PiBy8(), qc.Phase(),
             qc.U_f(lambda a, b: a | b, 3), qc.ControlledU(qc.ControlledU(qc.ControlledU(qc.PauliX()))),
             qc.CNOT().swap_qubits(0, 1)]
    for g in gates:
        k = g.rank // 2
        for _ in range(6):
            qi = [int(q) for q in rng.permutation(n)[:k]]
            got = g(s, qubit_indices=qi)
            want = oracle_apply(g.to_matrix(), vec, qi)
            check_close(got.to_column_vector(), want, dtype, 'test_structured_gates')


@pytest.mark.parametrize('dtype', DTYPES)
def test_wide_operators_use_the_dense_path(dtype):
    rng = np.random.default_rng(50)
    for n, k in ((5, 5), (7, 5), (9, 6), (11, 7), (8, 8)):
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        qi = [int(q) for q in rng.permutation(n)[:k]]
        U = rand_unitary(rng, k)
        got = qc.Operator.from_matrix(U)(s, qubit_indices=qi)
        want = oracle_apply(U, vec, qi)
        check_close(got.to_column_vector(), want, dtype, 'test_wide_operators_use_the_dense_path')


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_non_unitary_operators_and_lazy_norm_chain(dtype, mode):
    """Every Operator(State) renormalises (operators.py:275-276): projectors and arbitrary matrices
    give unit states, and the deferred scale survives a chain of passes."""
    _lib.load().qcs_set_tile_mode(mode)
    rng = np.random.default_rng(60)
    n = 11
    vec = rand_state_vec(rng, n)
    s = _state(vec, dtype)
    want = vec
    for step in range(6):
        k = 1 + step % 3
        qi = [int(q) for q in rng.permutation(n)[:k]]
        A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        s = qc.Operator.from_matrix(A)(s, qubit_indices=qi)
        want = oracle_apply(A, want, qi)
        check_close(s.to_column_vector(), want, dtype, 'test_non_unitary_operators_and_lazy_norm_chain')
    probs = s.probabilities
    assert abs(probs.sum() - 1.0) < 1e-6 and probs.shape == (2,) * n
    P0 = qc.Operator.from_matrix(np.array([[1, 0], [0, 0]]))
    got = P0(s, qubit_indices=[4])
    check_close(got.to_column_vector(), oracle_apply(np.array([[1, 0], [0, 0]]), want, [4]), dtype, 'test_non_unitary_operators_and_lazy_norm_chain')


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_fused_pass_matches_sequential_oracle(dtype, mode):
    """qcs_apply_fused: a run of gates in ONE pass == the same gates one at a time."""
    lib = _lib.load()
    lib.qcs_set_tile_mode(mode)
    rng = np.random.default_rng(70 + mode)
    low = 10 if dtype == np.complex64 else 9
    for n in (6, 12, 16, 19):
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        # two high bits (if any) + everything low is fair game for dense ops; diagonals/controls anywhere
        high = [int(b) for b in rng.permutation(np.arange(low, n))[:2]] if n > low else []
        allowed = list(range(min(low, n))) + high
        plans, want = [], vec
        for _ in range(14):
            kind = rng.integers(4)
            if kind == 0:        # dense 1..3 qubits on allowed bits
                k = int(rng.integers(1, 4))
                bits = [int(b) for b in rng.permutation(allowed)[:k]]
                g = qc.Operator.from_matrix(rand_unitary(rng, k))
            elif kind == 1:      # diagonal anywhere
                bits = [int(b) for b in rng.permutation(n)[:2]]
                g = qc.ControlledU(np.exp(0.3j) * qc.RotationZ(float(rng.uniform(0, 6))))
            elif kind == 2:      # CNOT: control anywhere, target allowed
                tgt = int(rng.choice(allowed))
                ctl = int(rng.choice([b for b in range(n) if b != tgt]))
                bits = [ctl, tgt]
                g = qc.CNOT()
            else:
                bits = [int(rng.integers(n))]
                g = qc.RotationZ(float(rng.uniform(0, 6)))
            plans.append((g._plan(), bits, False))
            want = orc.apply_matrix_flat(g.to_matrix(), want, bits)
        ops, keep = make_ops(plans)
        got = s._dv.apply_ops(ops, keep)
        check_close(got.to_host(), want, dtype, 'test_fused_pass_matches_sequential_oracle')
        # in place (dst == src) gives the same result
        dv = s._dv.clone()
        dv2 = dv.apply_ops(ops, keep, out=dv)
        check_close(dv2.to_host(), want, dtype, 'test_fused_pass_matches_sequential_oracle')


def test_fused_pass_rejects_too_many_high_bits():
    lib = _lib.load()
    s = qc.zeros(24, dtype=np.complex64)
    plans = [(qc.Hadamard()._plan(), [b], False) for b in range(10, 24)]
    ops, keep = make_ops(plans)
    with pytest.raises(_lib.QcsError):
        s._dv.apply_ops(ops, keep)
    assert b'tile' in lib.qcs_last_error()


@pytest.mark.parametrize('dtype', DTYPES)
def test_golden_apply_cases(golden, dtype):
    for c in range(int(golden['apply_count'])):
        vin, M = golden['apply%d_in' % c], golden['apply%d_op' % c]
        qi = [int(q) for q in golden['apply%d_qi' % c]]
        got = qc.Operator.from_matrix(M)(_state(vin, dtype), qubit_indices=qi)
        check_close(got.to_column_vector(), golden['apply%d_out' % c], dtype, 'test_golden_apply_cases')
    s = _state(golden['proj_in'], dtype)
    got = qc.Operator.from_matrix(np.array([[1, 0], [0, 0]]))(s, qubit_indices=[2])
    check_close(got.to_column_vector(), golden['proj_out'], dtype, 'test_golden_apply_cases')
    got = qc.Operator.from_matrix(golden['nonunitary_op'])(s, qubit_indices=[4, 1])
    check_close(got.to_column_vector(), golden['nonunitary_out'], dtype, 'test_golden_apply_cases')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('n,depth', [(3, 4), (5, 6), (8, 6), (10, 8), (12, 4)])
def test_golden_layered_circuits(golden, dtype, n, depth):
    mk = {'H': lambda t: qc.Hadamard(), 'RX': qc.RotationX, 'RZ': qc.RotationZ, 'CNOT': lambda t: qc.CNOT(),
          'CZ': lambda t: qc.ControlledU(qc.PauliZ())}
    state = qc.zeros(n, dtype=dtype)
    for name, theta, qs in orc.layered_circuit(n, depth, seed=1234):
        state = mk[name](theta)(state, qubit_indices=list(qs))
    tol = TOL[np.dtype(dtype)]
    assert rel_err(state.to_column_vector(), golden['layered_n%d_d%d_amps' % (n, depth)]) < tol
    assert np.max(np.abs(state.probabilities.reshape(-1) - golden['layered_n%d_d%d_probs' % (n, depth)])) < tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_golden_measurements_with_numpy_global_stream(golden, dtype):
    """np.random.seed(s) reproduces the reference's outcomes and post-measurement states."""
    tol = TOL[np.dtype(dtype)]
    for c in range(int(golden['meas_count'])):
        vin = golden['meas%d_in' % c]
        qi = golden['meas%d_qi' % c]
        n = int(round(np.log2(vin.size)))
        if qi.ndim == 0:
            arg = None if int(qi) < 0 else int(qi)
        else:
            arg = [int(q) for q in qi]
        remove = bool(golden['meas%d_remove' % c])
        want_bits, want_out = golden['meas%d_bits' % c], golden['meas%d_out' % c]
        # (a) through NumPy's global stream, exactly like the reference
        s = _state(vin, dtype)
        seed = 1000 + c % 3
        np.random.seed(seed)
        bits = s.measure(qubit_indices=arg, remove=remove)
        # (b) through a caller-supplied uniform
        s2 = _state(vin, dtype)
        bits2 = s2.measure(qubit_indices=arg, remove=remove, u=float(golden['meas%d_u' % c]))
        for b, st in ((bits, s), (bits2, s2)):
            if isinstance(arg, int):
                assert isinstance(b, int) and b == int(want_bits)
            else:
                assert isinstance(b, tuple) and b == tuple(int(x) for x in want_bits), c
            assert rel_err(np.asarray(st._t).reshape(-1), want_out.reshape(-1)) < tol, c
            assert st.rank == (n - (1 if isinstance(arg, int) else (n if arg is None else len(arg))) if remove else n)


@pytest.mark.parametrize('dtype', DTYPES)
def test_marginals_and_collapse_random_subsets(dtype):
    rng = np.random.default_rng(80)
    tol = TOL[np.dtype(dtype)]
    for n in (4, 9, 13, 16):
        vec = rand_state_vec(rng, n)
        for m in (1, 2, 3, min(n, 11), n):
            qi = [int(q) for q in rng.permutation(n)[:m]]
            s = _state(vec, dtype)
            ps = s.marginal_probabilities(qi)
            want_ps, _, _ = orc.measurement_probabilities(vec.reshape([2] * n), qi)
            assert np.max(np.abs(ps - want_ps)) < tol
            for remove in (False, True):
                u = float(rng.uniform())
                s = _state(vec, dtype)
                bits = s.measure(qi, remove=remove, u=u)
                want_bits, want_t = orc.measure(vec.reshape([2] * n), qi, u, remove=remove)
                if orc.sample_outcome(want_ps, u) == orc.sample_outcome(ps, u):
                    assert tuple(bits) == tuple(want_bits)
                    assert rel_err(np.asarray(s._t).reshape(-1), np.asarray(want_t).reshape(-1)) < tol
                # measuring again gives the same answer (tests.py FastMeasurementTests)
                if not remove:
                    assert s.measure(qi, u=float(rng.uniform())) == bits


@pytest.mark.parametrize('dtype', DTYPES)
def test_state_arithmetic_and_tensor_product(golden, dtype):
    tol = TOL[np.dtype(dtype)]
    a, b = _state(golden['kron_a'], dtype), _state(golden['kron_b'], dtype)
    assert rel_err((a * b).to_column_vector(), golden['kron_ab']) < tol
    assert rel_err((b ** 3).to_column_vector(), golden['kron_pow3']) < tol
    ar = _state(golden['kron_a'].reshape([2] * 3)[::-1].reshape(-1), dtype)
    assert rel_err((a + ar).to_column_vector(), golden['add_ab']) < tol
    assert abs(a.dot(ar) - golden['dot_ab']) < tol
    assert rel_err((a - ar).to_column_vector(), golden['kron_a'] - ar.to_column_vector()) < tol
    assert rel_err((2 * a).to_column_vector(), 2 * golden['kron_a']) < tol
    assert rel_err((a * (0.5 + 0.5j)).to_column_vector(), (0.5 + 0.5j) * golden['kron_a']) < tol
    assert rel_err((a / 4).to_column_vector(), golden['kron_a'] / 4) < tol
    assert rel_err((-a).to_column_vector(), -golden['kron_a']) < tol
    # (zeros + ones)/sqrt(2) == positive_superposition (tests.py:652-705)
    z, o = qc.zeros(1, dtype=dtype), qc.ones(1, dtype=dtype)
    assert rel_err(((z + o) / np.sqrt(2)).to_column_vector(), qc.positive_superposition(1, dtype=dtype).to_column_vector()) < tol
    # lazily normalised operands combine with their scales applied
    h = qc.Hadamard()(z)
    assert rel_err((h * h).to_column_vector(), np.full(4, 0.5)) < tol
    assert rel_err((h + h).to_column_vector(), np.full(2, np.sqrt(2))) < tol
    s = qc.State(np.array([3.0, 4.0]), dtype=dtype)
    s.renormalize_()
    assert rel_err(s.to_column_vector(), np.array([0.6, 0.8])) < tol
    assert abs(s.dot(s) - 1) < tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_factories_and_bit_order(dtype):
    """Qubit 0 is the most significant bit of the column vector (tests.py:811-821)."""
    assert np.array_equal(qc.zeros(3, dtype=dtype).to_column_vector(), np.eye(8)[0])
    assert np.array_equal(qc.ones(3, dtype=dtype).to_column_vector(), np.eye(8)[7])
    assert np.array_equal(qc.bitstring(1, 0, 0, dtype=dtype).to_column_vector(), np.eye(8)[4])
    assert np.array_equal(qc.bitstring(0, 1, 1, dtype=dtype)._t, np.eye(8)[3].reshape(2, 2, 2))
    s = qc.bitstring(0, 1, 1, 0, dtype=dtype)
    assert s[0, 1, 1, 0] == 1 and s.shape == (2, 2, 2, 2) and s.rank == 4
    assert s.measure() == (0, 1, 1, 0)
    assert s.measure(2) == 1 and s.measure([3, 1]) == (0, 1)
    assert s.measure(1, remove=True) == 1 and s.rank == 3
    b = qc.bell_state(0, 1, dtype=dtype)
    check_close(b.to_column_vector(), np.array([0, 1, 1, 0]) / np.sqrt(2), dtype, 'test_factories_and_bit_order')
    q = qc.qubit(theta=0.7, phi=0.2, dtype=dtype)
    check_close(q.to_column_vector(), [np.cos(0.35), np.sin(0.35) * np.exp(0.2j)], dtype, 'test_factories_and_bit_order')
    with pytest.raises(ValueError):
        qc.zeros(0)
    with pytest.raises(ValueError):
        qc.State.from_column_vector(np.ones(3))
    with pytest.raises(ValueError):
        s.measure([])
    with pytest.raises(ValueError):
        s.measure([0, 0])
    with pytest.raises(ValueError):
        s.measure([5])
    with pytest.raises(AssertionError):
        qc.State(np.array([1.0, 1.0]), dtype=dtype).probabilities


@pytest.mark.parametrize('dtype', DTYPES)
def test_permute_and_swap_qubits(dtype):
    rng = np.random.default_rng(90)
    tol = TOL[np.dtype(dtype)]
    for n in (2, 5, 9):
        vec = rand_state_vec(rng, n)
        axes = [int(a) for a in rng.permutation(n)]
        s = _state(vec, dtype)
        ret = s.permute_qubits(axes)
        assert ret is s
        assert rel_err(s._t, np.transpose(vec.reshape([2] * n), axes)) < tol
        s.permute_qubits(axes, inverse=True)
        assert rel_err(s.to_column_vector(), vec) < tol
        s.swap_qubits(0, n - 1)
        assert rel_err(s._t, np.swapaxes(vec.reshape([2] * n), 0, n - 1)) < tol
        c = copy.deepcopy(s)
        c.swap_qubits(0, n - 1)
        assert rel_err(c.to_column_vector(), vec) < tol and rel_err(s._t, np.swapaxes(vec.reshape([2] * n), 0, n - 1)) < tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_density_operator_against_reference(golden, dtype):
    tol = TOL[np.dtype(dtype)]
    states = [_state(v, dtype) for v in golden['rho_states']]
    rho = qc.DensityOperator.from_ensemble(states, golden['rho_ps'])
    assert rho.rank == 6
    assert rel_err(rho._t.reshape(-1), golden['rho_t']) < tol
    assert np.max(np.abs(rho.probabilities.reshape(-1) - golden['rho_probs'])) < tol
    U2 = qc.Operator.from_matrix(golden['rho_op'])
    rho2 = U2(rho, qubit_indices=[2, 0])
    assert isinstance(rho2, qc.DensityOperator)
    assert rel_err(rho2._t.reshape(-1), golden['rho_applied']) < tol
    assert rel_err(rho._t.reshape(-1), golden['rho_t']) < tol          # argument untouched
    assert rel_err(qc.Hadamard()(rho, qubit_indices=[1])._t.reshape(-1), golden['rho_h1']) < tol
    assert rel_err(rho.reduced_density_operator([0, 2])._t.reshape(-1), golden['rho_reduced_02']) < tol
    for c in range(int(golden['rhomeas_count'])):
        qi = golden['rhomeas%d_qi' % c]
        arg = None if qi.ndim == 0 else [int(q) for q in qi]
        remove = bool(golden['rhomeas%d_remove' % c])
        r = copy.deepcopy(rho2)
        np.random.seed(2000 + c)
        bits = r.measure(qubit_indices=arg, remove=remove)
        assert bits == tuple(int(b) for b in golden['rhomeas%d_bits' % c])
        assert rel_err(np.asarray(r._t).reshape(-1), golden['rhomeas%d_out' % c]) < tol
    # state / density consistency (tests.py:1146-1190): same outcome -> same post-measurement state
    s = states[0]
    r = s.density_operator()
    u = 0.37
    sb = copy.deepcopy(s)
    bits_s = sb.measure([1, 2], u=u)
    bits_r = r.measure([1, 2], u=u)
    assert bits_s == bits_r
    assert rel_err(r._t.reshape(-1), sb.density_operator()._t.reshape(-1)) < tol
    # purify round trip (host eig; tests.py ReducedDensityAndPurificationTests)
    pure = rho.purify()
    assert rel_err(pure.reduced_density_operator([0, 1, 2])._t.reshape(-1), golden['rho_t']) < 1e-4


@pytest.mark.parametrize('dtype', DTYPES)
def test_density_operator_larger_system_vs_oracle(dtype):
    rng = np.random.default_rng(100)
    tol = TOL[np.dtype(dtype)]
    n = 6
    vecs = [rand_state_vec(rng, n) for _ in range(2)]
    rho = qc.DensityOperator.from_ensemble([_state(v, dtype) for v in vecs], [0.25, 0.75])
    want = orc.from_ensemble([v.reshape([2] * n) for v in vecs], [0.25, 0.75])
    for k, qi in ((1, [5]), (1, [0]), (2, [3, 1]), (3, [0, 5, 2]), (4, [1, 2, 3, 4]), (5, [4, 0, 2, 5, 1])):
        U = rand_unitary(rng, k)
        rho = qc.Operator.from_matrix(U)(rho, qubit_indices=qi)
        want = orc.apply_operator_density(orc.op_from_matrix(U), want, qi)
        assert rel_err(rho._t.reshape(-1), want.reshape(-1)) < tol, qi
    rho.swap_qubits(1, 4)
    want_sw = np.swapaxes(np.swapaxes(want, 2, 8), 3, 9)
    assert rel_err(rho._t.reshape(-1), want_sw.reshape(-1)) < tol
    ps, _ = rho._measurement_probabilities([4, 0])
    assert np.max(np.abs(ps - orc.density_measurement_probabilities(want_sw, [4, 0]))) < tol


def test_grover_reference_example_d5(golden):
    """examples/grover_algorithm.py (config 1, scaled down): composed operators on the host, the
    6-qubit Grover iterate applied through the dense path, seeded measurement."""
    d = 5
    answers = golden['grover_d5_answers']

    def f(*bits):
        return answers[sum(v * 2 ** i for i, v in enumerate(bits))]

    Oracle = qc.U_f(f, d=d + 1)
    H_d = qc.Hadamard(d)
    N = 2 ** d
    proj = np.zeros((N, N))
    proj[0, 0] = 1
    Inversion = H_d((2 * qc.Operator.from_matrix(proj) - qc.Identity(d))(H_d))
    Grover = Inversion(Oracle, qubit_indices=range(d))
    assert np.max(np.abs(Grover.to_matrix() - golden['grover_d5_op'])) < 1e-12
    state = qc.zeros(d) * qc.ones(1)
    state = (H_d * qc.Hadamard())(state)
    for _ in range(4):
        state = Grover(state)
    assert rel_err(state.to_column_vector(), golden['grover_d5_amps']) < 1e-11
    np.random.seed(5)
    bits = state.measure(qubit_indices=range(d))
    assert bits == tuple(int(b) for b in golden['grover_d5_bits'])
    assert f(*bits) == 1


def test_c_abi_direct_call_and_error_codes():
    """Call qcs_apply_gate through ctypes with raw device pointers, as a foreign host would."""
    import torch
    lib = _lib.load()
    n = 12
    rng = np.random.default_rng(110)
    vec = rand_state_vec(rng, n)
    src = torch.from_numpy(vec.view(np.float64).copy()).cuda()
    dst = torch.empty_like(src)
    norm = torch.zeros(1, dtype=torch.float64, device='cuda')
    ws = torch.zeros(int(lib.qcs_workspace_bytes()), dtype=torch.uint8, device='cuda')
    U = np.ascontiguousarray(rand_unitary(rng, 2))
    bitpos = (ctypes.c_int32 * 2)(11, 0)
    rc = lib.qcs_apply_gate(dst.data_ptr(), src.data_ptr(), n, _lib.QCS_C128, U.ctypes.data, 2, bitpos, None,
                            norm.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    got = dst.cpu().numpy().view(np.complex128)
    want = orc.apply_matrix_flat(U, vec, [11, 0], renorm=False)
    assert rel_err(got, want) < 1e-12
    assert abs(float(norm.item()) - 1.0) < 1e-12
    bad = (ctypes.c_int32 * 2)(3, 3)
    assert lib.qcs_apply_gate(dst.data_ptr(), src.data_ptr(), n, _lib.QCS_C128, U.ctypes.data, 2, bad, None, None,
                              ws.data_ptr(), 0) == -1
    assert b'repeated' in lib.qcs_last_error()
    assert lib.qcs_apply_gate(dst.data_ptr(), src.data_ptr(), n, 9, U.ctypes.data, 2, bitpos, None, None,
                              ws.data_ptr(), 0) == -1
    assert lib.qcs_apply_gate(None, src.data_ptr(), n, _lib.QCS_C128, U.ctypes.data, 2, bitpos, None, None,
                              ws.data_ptr(), 0) == -1


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('dtype', DTYPES)
def test_circuit_run_fused_matches_oracle(dtype, mode):
    from qcircuits_b200.workloads import layered_circuit
    _lib.load().qcs_set_tile_mode(mode)
    tol = TOL[np.dtype(dtype)]
    for n, depth in ((5, 4), (14, 6), (18, 5)):
        circ = layered_circuit(n, depth, seed=1234)
        want = orc.run_circuit(orc.zeros_state(n), orc.layered_circuit(n, depth, seed=1234), flat=True).reshape(-1)
        s0 = qc.zeros(n, dtype=dtype)
        for max_ops in (4, 24, 64):
            got = circ.run(s0, fuse=True, max_ops=max_ops)
            assert rel_err(got.to_column_vector(), want) < tol, (n, max_ops)
        got = circ.run(s0, fuse=False)
        assert rel_err(got.to_column_vector(), want) < tol
        basis0 = np.zeros(2 ** n)
        basis0[0] = 1
        assert np.array_equal(s0.to_column_vector(), basis0)     # argument untouched
        passes, gates = circ.stats(dtype)
        assert gates == len(orc.layered_circuit(n, depth)) and passes < gates


def test_circuit_run_host_buffers():
    import torch
    from qcircuits_b200.workloads import layered_circuit
    n = 15
    circ = layered_circuit(n, 5, seed=7)
    want = orc.run_circuit(orc.zeros_state(n), orc.layered_circuit(n, 5, seed=7), flat=True).reshape(-1)
    host_in = torch.zeros(2 ** n, dtype=torch.complex64).pin_memory()
    host_in[0] = 1
    host_out = torch.empty(2 ** n, dtype=torch.complex64).pin_memory()
    circ.run_host(host_in, host_out)
    assert rel_err(host_out.numpy(), want) < 4e-5


@pytest.mark.parametrize('dtype,n', [(np.complex64, 30), (np.complex128, 28)])
def test_full_size_properties(dtype, n):
    """BASELINE sizes, where the oracle cannot run: size-independent properties.
    (a) H on every qubit of |0..0> is the uniform superposition; (b) a circuit followed by its inverse
    returns |0..0>; (c) fused and unfused schedules agree; (d) marginals of a product state factorise."""
    from qcircuits_b200.circuit import Circuit
    from qcircuits_b200.workloads import layered_gate_list, operator_for
    tol = 1e-4 if dtype == np.complex64 else 1e-11
    H = qc.Hadamard()
    c = Circuit(n)
    for q in range(n):
        c.append(H, [q])
    s = c.run(qc.zeros(n, dtype=dtype))
    ps = s.marginal_probabilities([0, n // 2, n - 1])
    assert np.max(np.abs(ps - 1 / 8)) < tol
    sample = s._dv.buf[:64].cpu().numpy().view(dtype)
    scale = 1 / np.sqrt(float(s._dv.norm[0].item()))
    assert np.max(np.abs(sample * scale - 2.0 ** (-n / 2))) < tol * 2.0 ** (-n / 2)
    del s
    # (b) + (c): two layers, then the inverse gates in reverse order
    gates = layered_gate_list(n, 2, seed=3)
    fwd = Circuit(n)
    both = Circuit(n)
    for name, theta, qs in gates:
        fwd.append(operator_for(name, theta), list(qs))
        both.append(operator_for(name, theta), list(qs))
    for name, theta, qs in reversed(gates):
        both.append(operator_for(name, theta).adj, list(qs))
    z = qc.zeros(n, dtype=dtype)
    back = both.run(z)
    ps = back.marginal_probabilities(list(range(0, n, 3)))
    assert abs(ps[0] - 1.0) < tol and abs(ps.sum() - 1.0) < tol
    amp0 = back._dv.buf[:2].cpu().numpy().view(dtype)[0] / np.sqrt(float(back._dv.norm[0].item()))
    assert abs(abs(amp0) - 1.0) < tol
    del back
    a = fwd.run(z, fuse=True)
    b = fwd.run(z, fuse=False)
    qi = [0, 5, n - 1, n - 2, 11]
    assert np.max(np.abs(a.marginal_probabilities(qi) - b.marginal_probabilities(qi))) < tol
    assert abs(a.dot(b) - 1.0) < tol
    # measurement at full size: outcome is consistent with the marginals and collapses the state
    bits = a.measure(qi, u=0.5)
    assert a.marginal_probabilities(qi)[int(''.join(map(str, bits)), 2)] > 1 - tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_u_f_is_applied_as_an_index_permutation(dtype):
    """U_f(state) through the permutation kernel (qcs_apply_uf) == the oracle applying the full permutation matrix;
    the operator tensor is not built on that path, and everything that needs the tensor still gets the reference's."""
    rng = np.random.default_rng(90)
    for d, n in ((2, 2), (2, 7), (3, 9), (5, 5), (6, 13), (9, 14), (11, 11)):
        table = rng.integers(0, 2, size=2 ** (d - 1))
        f = lambda *bits, table=table: int(table[int(''.join(map(str, bits)), 2)])
        op = qc.U_f(f, d)
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        qi = [int(q) for q in rng.permutation(n)[:d]]
        got = op(s, qubit_indices=qi)
        assert op._tensor is None                      # the 2^(2d) tensor was never materialised
        cols = np.arange(2 ** d)
        M = np.zeros((2 ** d, 2 ** d))
        M[(cols & ~1) | ((cols & 1) ^ table[cols >> 1]), cols] = 1
        want = oracle_apply(M, vec, qi)
        check_close(got.to_column_vector(), want, dtype, 'test_u_f_is_applied_as_an_index_permutation')
        check_close(s.to_column_vector(), vec, dtype, 'test_u_f_is_applied_as_an_index_permutation')   # argument untouched
    # a lazily scaled (non-unit buffer) argument, then the tensor paths: composition and density operators
    op = qc.U_f(lambda a, b: a ^ b, 3)
    s = qc.Operator.from_matrix(np.array([[2.0, 0], [0, 0.5]]))(_state(rand_state_vec(rng, 4), dtype), qubit_indices=[1])
    want = oracle_apply(op.to_matrix(), s.to_column_vector(), [3, 0, 2])
    check_close(op(s, qubit_indices=[3, 0, 2]).to_column_vector(), want, dtype, 'test_u_f_is_applied_as_an_index_permutation')
    comp = qc.Hadamard(3)(qc.U_f(lambda a, b: a & b, 3))
    assert np.max(np.abs(comp.to_matrix() - qc.Hadamard(3).to_matrix() @ qc.U_f(lambda a, b: a & b, 3).to_matrix())) < 1e-14
    rho = _state(rand_state_vec(rng, 3), dtype).density_operator()
    out = qc.U_f(lambda a, b: a | b, 3)(rho)
    U = qc.U_f(lambda a, b: a | b, 3).to_matrix()
    want_rho = U @ rho.to_matrix() @ U.conj().T
    assert np.max(np.abs(out.to_matrix() - want_rho)) < TOL[np.dtype(dtype)]


@pytest.mark.parametrize('dtype', DTYPES)
def test_operator_relabelled_after_use_is_reanalysed(dtype):
    """apply, swap_qubits / permute_qubits (in place, reference operators.py:125-170), apply again: the second
    application must see the new tensor, not a cached plan or device matrix"""
    rng = np.random.default_rng(91)
    n = 9
    vec = rand_state_vec(rng, n)
    s = _state(vec, dtype)
    op = qc.CNOT()
    check_close(op(s, qubit_indices=[2, 5]).to_column_vector(), oracle_apply(op.to_matrix(), vec, [2, 5]), dtype, 'relabel')
    op.swap_qubits(0, 1)
    check_close(op(s, qubit_indices=[2, 5]).to_column_vector(), oracle_apply(op.to_matrix(), vec, [2, 5]), dtype, 'relabel')
    big = qc.Operator.from_matrix(rand_unitary(rng, 6))
    qi = [0, 8, 3, 4, 1, 6]
    check_close(big(s, qubit_indices=qi).to_column_vector(), oracle_apply(big.to_matrix(), vec, qi), dtype, 'relabel')
    big.permute_qubits([5, 3, 0, 1, 2, 4])
    check_close(big(s, qubit_indices=qi).to_column_vector(), oracle_apply(big.to_matrix(), vec, qi), dtype, 'relabel')


def test_ensemble_of_mixed_amplitude_types_is_rejected():
    a = qc.zeros(3, dtype=np.complex64)
    b = qc.zeros(3, dtype=np.complex128)
    with pytest.raises(TypeError):
        qc.DensityOperator.from_ensemble([a, b], [0.5, 0.5])


@pytest.mark.parametrize('dtype', DTYPES)
def test_circuit_whose_pass_ends_in_a_dense_gate(dtype):
    """advisor repro: H, RX, then a 3-qubit unitary / U_f last in its pass -- program order must survive scheduling"""
    from qcircuits_b200.circuit import Circuit
    rng = np.random.default_rng(92)
    n = 12
    for last in (qc.Operator.from_matrix(rand_unitary(rng, 3)), qc.U_f(lambda a, b: a ^ b, 3)):
        c = Circuit(n)
        c.append(qc.Hadamard(), [0])
        c.append(qc.RotationX(0.3), [1])
        c.append(last, [0, 1, 2])
        vec = rand_state_vec(rng, n)
        want = vec
        for g in c.gates:
            want = orc.apply_matrix_flat(g.plan.full, want, g.bits, renorm=False)
        got = c.run(_state(vec, dtype)).to_column_vector()
        check_close(got, want / np.linalg.norm(want), dtype, 'test_circuit_whose_pass_ends_in_a_dense_gate')
