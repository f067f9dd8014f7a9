"""The large-state code paths of the streaming kernels (segment-decomposed marginals and bit permutation,
vectorised tensor product and probabilities) against NumPy on states big enough to reach them."""
import numpy as np
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import _lib
from tests.util import check_close, TOL, rand_state_vec, rel_err

pytestmark = pytest.mark.gpu
DTYPES = [np.complex128, np.complex64]


def _state(vec, dtype):
    return qc.State.from_column_vector(np.asarray(vec), dtype=dtype)


@pytest.mark.parametrize('dtype', DTYPES)
def test_marginals_of_large_states_every_bit_class(dtype):
    """Measured bits inside the per-thread part (0..8), the per-step part (9..11) and the segment part (>= 12),
    few and many outcomes (shared-memory histogram up to 2^10, global atomics above)."""
    rng = np.random.default_rng(5)
    tol = TOL[np.dtype(dtype)]
    for n in (15, 18, 20):
        vec = rand_state_vec(rng, n)
        s = _state(vec, dtype)
        t = np.abs(vec.reshape([2] * n)) ** 2
        cases = [[n - 1], [0], [n - 10], [n - 12], [n - 13], [0, n - 1], [n - 9, n - 11, n - 13, 0],
                 list(range(n - 1, n - 11, -1)), list(range(0, 10)), [int(q) for q in rng.permutation(n)[:7]],
                 [int(q) for q in rng.permutation(n)[:12]], [int(q) for q in rng.permutation(n)[:n - 1]]]
        for qi in cases:
            ps = s.marginal_probabilities(qi)
            rest = tuple(a for a in range(n) if a not in qi)
            want = np.transpose(t.sum(axis=rest) if rest else t, np.argsort(np.argsort(qi))).reshape(-1)
            assert np.max(np.abs(ps - want)) < tol, (n, qi)
            assert abs(ps.sum() - 1) < 10 * tol


def test_marginals_of_a_real_probability_vector():
    """QCS_F64 input (DensityOperator marginals run on the diagonal probabilities)."""
    import torch
    from qcircuits_b200._device import DeviceVector
    rng = np.random.default_rng(6)
    n = 16
    p = rng.uniform(size=2 ** n)
    p /= p.sum()
    dev = torch.from_numpy(p).cuda()
    dv = DeviceVector(_lib.alloc_amplitudes(1, np.complex128), None, 1, np.complex128)
    for bits in ([n - 1, 0, 7], [3], list(range(4, 15)), [15, 14, 13, 12, 11]):
        got = dv.marginals(bits, real_input=dev)
        t = p.reshape([2] * n)
        axes = [n - 1 - b for b in bits]
        rest = tuple(a for a in range(n) if a not in axes)
        want = np.transpose(t.sum(axis=rest), np.argsort(np.argsort(axes))).reshape(-1)
        assert np.max(np.abs(got - want)) < 1e-13, bits


@pytest.mark.parametrize('dtype', DTYPES)
def test_permute_qubits_of_large_states(dtype):
    rng = np.random.default_rng(7)
    tol = TOL[np.dtype(dtype)]
    for n in (15, 17, 19):
        vec = rand_state_vec(rng, n)
        perms = [[int(a) for a in rng.permutation(n)] for _ in range(3)]
        keep_last = [int(a) for a in rng.permutation(n - 1)] + [n - 1]       # flat bit 0 stays in place
        swap_ends = list(range(n))
        swap_ends[0], swap_ends[n - 1] = n - 1, 0
        # flat bit 0 (the last qubit) traded with near and far bits: the 2 x 2 block form of the complex64 kernel
        moved = []
        for other in (n - 2, n - 3, n - 8, n - 11, n - 13, 1):
            axes = list(range(n))
            axes[n - 1], axes[other] = other, n - 1
            moved.append(axes)
        cyc = [n - 1] + list(range(n - 1))                                   # every bit moves up by one
        for axes in perms + [keep_last, swap_ends, list(range(n)), cyc] + moved:
            s = _state(vec, dtype)
            s.permute_qubits(axes)
            assert rel_err(s._t, np.transpose(vec.reshape([2] * n), axes)) < tol, (n, axes)
            s.permute_qubits(axes, inverse=True)
            assert rel_err(s.to_column_vector(), vec) < tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_tensor_product_and_probabilities_of_large_states(dtype):
    rng = np.random.default_rng(8)
    tol = TOL[np.dtype(dtype)]
    for na, nb in ((9, 8), (16, 1), (1, 16), (12, 5)):
        a, b = rand_state_vec(rng, na), rand_state_vec(rng, nb)
        got = _state(a, dtype) * _state(b, dtype)
        want = np.kron(a, b)
        assert rel_err(got.to_column_vector(), want) < tol
        probs = got.probabilities
        assert probs.shape == (2,) * (na + nb)
        assert np.max(np.abs(probs.reshape(-1) - np.abs(want) ** 2)) < tol
        assert abs(got.dot(got) - 1) < 10 * tol


@pytest.mark.parametrize('dtype', DTYPES)
def test_partial_trace_against_the_oracle(dtype):
    """DensityOperator.reduced_density_operator on the device (qcs_partial_trace) vs the oracle's restatement
    of density_operator.py:127-138, for mixed states, any retain order, few and many traced qubits."""
    rng = np.random.default_rng(9)
    tol = TOL[np.dtype(dtype)]
    for d in (2, 4, 7):
        states = [rand_state_vec(rng, d).reshape([2] * d) for _ in range(3)]
        ps = rng.uniform(size=3)
        ps /= ps.sum()
        rho_t = orc.from_ensemble(states, ps)
        rho = qc.DensityOperator(rho_t, dtype=dtype)
        cases = [[0], [d - 1], list(range(d)), list(range(d - 1, -1, -1))]
        if d >= 4:
            cases += [[2, 0], [1, 3, 2], [int(q) for q in rng.permutation(d)[:d - 1]]]
        for retain in cases:
            got = rho.reduced_density_operator(retain)
            want = orc.reduced_tensor(rho_t, retain)
            assert got.rank == 2 * len(retain)
            assert rel_err(got._t.reshape(-1), np.asarray(want).reshape(-1)) < tol, (d, retain)
            assert abs(got.probabilities.sum() - 1) < 1e-4
    # a pure state through State.reduced_density_operator (state.py:285-305)
    psi = rand_state_vec(rng, 9)
    got = qc.State.from_column_vector(psi, dtype=dtype).reduced_density_operator([8, 3])
    want = orc.reduced_tensor(orc.outer_product(psi.reshape([2] * 9)), [8, 3])
    assert rel_err(got._t.reshape(-1), np.asarray(want).reshape(-1)) < tol


def test_operator_composition_on_the_device():
    """A(B) for an 8-qubit B runs as a gate pass over B's out-wires (Operator._compose_on_device); compared
    with the oracle's restatement of operators.py:225-278 for tile-engine operators (1-3 qubits, controlled,
    diagonal) and dense ones (5 and 8 qubits)."""
    from qcircuits_b200 import operators as ops_mod
    from tests.util import rand_unitary
    rng = np.random.default_rng(10)
    d = 8
    assert 2 * d >= ops_mod.DEVICE_COMPOSE_MIN_BITS
    B = qc.Operator.from_matrix(rand_unitary(rng, d))
    cases = [(qc.Hadamard(), [3]), (qc.CNOT(), [6, 1]), (qc.RotationZ(0.7), [0]), (qc.Toffoli(), [7, 2, 4]),
             (qc.Operator.from_matrix(rand_unitary(rng, 2)), [5, 0]),
             (qc.Operator.from_matrix(rand_unitary(rng, 5)), [4, 0, 7, 2, 1]),
             (qc.Operator.from_matrix(rand_unitary(rng, d)), None)]
    for A, qi in cases:
        got = A(B, qubit_indices=qi)
        want = orc.compose_operators(A._t, B._t, qi)
        assert isinstance(got, qc.Operator) and got.rank == 2 * d
        assert rel_err(got._t.reshape(-1), np.asarray(want).reshape(-1)) < 1e-12, qi
    # B itself is untouched (operators.py: "returns a new object")
    assert rel_err(B.to_matrix().reshape(-1), B.to_matrix().reshape(-1)) == 0


def test_device_resident_operator_chain_and_adjoint():
    """compositions and adjoints of large operators stay on the device (Operator._device_only) until
    somebody reads the tensor; a chain A(B(C)).adj applied to a State never downloads.  Compared with the oracle's
    restatement of operators.py:95-114 (adj) and :225-278 (composition), and with the matrix adjoint for both
    amplitude types of the bit-permutation kernels (qcs_conj_permute_bits)."""
    from qcircuits_b200 import operators as ops_mod
    from qcircuits_b200._device import DeviceVector
    from tests.util import rand_unitary
    rng = np.random.default_rng(21)
    d = 8
    A = qc.Operator.from_matrix(rand_unitary(rng, d))
    B = qc.Operator.from_matrix(rand_unitary(rng, d))
    H3 = qc.Operator.from_matrix(rand_unitary(rng, 3))
    AB = A(B)
    assert AB._device_only and AB.rank == 2 * d and AB.shape == (2,) * (2 * d)
    chain = H3(AB, qubit_indices=[5, 0, 2])
    assert chain._device_only
    adj = chain.adj
    assert adj._device_only
    want_chain = orc.compose_operators(H3._t, orc.compose_operators(A._t, B._t, None), [5, 0, 2])
    Mc = qc.Operator(np.asarray(want_chain)).to_matrix()
    # applied to a state without ever materialising the host tensor (dense path, matrix permuted on the device)
    vec = rand_state_vec(rng, d)
    got = adj(qc.State.from_column_vector(vec))
    assert adj._device_only and chain._device_only
    want = np.conj(Mc.T) @ vec
    assert rel_err(got.to_column_vector(), want / np.linalg.norm(want)) < 1e-12
    # on a density operator: U rho U^dagger needs the conjugated matrix, also made on the device
    rho = qc.State.from_column_vector(vec).density_operator()
    got_rho = chain(rho)
    assert chain._device_only
    out = Mc @ vec
    assert rel_err(got_rho.to_matrix().reshape(-1), np.outer(out, np.conj(out)).reshape(-1)) < 1e-12
    # reading the tensor downloads it, and it is the reference's
    assert rel_err(adj._t.reshape(-1), np.asarray(orc.op_adjoint(np.asarray(want_chain))).reshape(-1)) < 1e-12
    assert not adj._device_only
    assert rel_err(adj.to_matrix().reshape(-1), np.conj(Mc.T).reshape(-1)) < 1e-12
    # adjoint of a host operator of the same size goes through the device too; involution
    assert rel_err(A.adj.adj._t.reshape(-1), A._t.reshape(-1)) == 0
    # the kernel itself, both amplitude types, bit 0 moved (pair swap) and bit 0 in place, small and large
    for dtype in (np.complex64, np.complex128):
        for n in (6, 16, 20):
            x = rand_state_vec(rng, n)
            dv = DeviceVector.from_host(x.reshape([2] * n), dtype)
            for src_bit in ([b ^ 1 for b in range(n)], [0] + [int(b) for b in 1 + rng.permutation(n - 1)]):
                got = dv.permuted(src_bit, conj=True).to_host()
                i = np.arange(2 ** n)
                j = np.zeros_like(i)
                for b in range(n):
                    j |= ((i >> b) & 1) << src_bit[b]
                ref = np.conj(x.astype(dtype)[j])
                assert np.array_equal(got, ref.astype(np.complex128)), (dtype, n)


@pytest.mark.parametrize('dtype', DTYPES)
def test_measuring_more_qubits_than_one_chunk(dtype):
    """State.measure of more than MEASURE_CHUNK_BITS qubits (e.g. the default measure() of a large state) draws
    the outcome hierarchically; it must be the outcome of the reference's flat rule for the same uniform, and
    the collapsed state must match."""
    from qcircuits_b200.state import MEASURE_CHUNK_BITS
    rng = np.random.default_rng(13)
    tol = TOL[np.dtype(dtype)]
    n = MEASURE_CHUNK_BITS + 4
    vec = rand_state_vec(rng, n)
    vec[rng.uniform(size=vec.size) < 0.5] = 0          # make intervals of very different widths
    vec /= np.linalg.norm(vec)
    for qi, remove in ((list(range(n)), False), ([int(q) for q in rng.permutation(n)[:n - 1]], True),
                       ([int(q) for q in rng.permutation(n)[:MEASURE_CHUNK_BITS + 1]], False)):
        u = float(rng.uniform())
        s = _state(vec, dtype)
        bits = s.measure(qi, remove=remove, u=u)
        want_ps, _, _ = orc.measurement_probabilities(vec.reshape([2] * n), qi)
        want_bits, want_t = orc.measure(vec.reshape([2] * n), qi, u, remove=remove)
        o = orc.sample_outcome(want_ps, u)
        cdf = np.cumsum(want_ps) / want_ps.sum()
        edge = min(abs(u - cdf[o]), abs(u - (cdf[o - 1] if o else 0.0)))
        if edge > 1e-6:                                    # (complex64 marginals differ from the oracle's by ~1e-7)
            assert tuple(bits) == tuple(want_bits), (qi, remove)
            assert rel_err(np.asarray(s._t).reshape(-1), np.asarray(want_t).reshape(-1)) < tol
    # default measure(): all qubits, global NumPy stream -- deterministic under a seed, state collapses to a basis state
    s = _state(vec, dtype)
    np.random.seed(11)
    bits = s.measure()
    assert len(bits) == n
    idx = int(''.join(str(b) for b in bits), 2)
    assert abs(vec[idx]) > 0
    assert abs(abs(s.to_column_vector()[idx]) - 1) < 10 * tol
    np.random.seed(11)
    assert _state(vec, dtype).measure() == bits


@pytest.mark.parametrize('dtype', DTYPES)
def test_controlled_two_qubit_unitaries_small_and_many_controls(dtype):
    """The shared-memory dense op still serves two-target matrices where the register arm cannot: tiny states
    (every tile bit is a register bit) and more than two controls."""
    from tests.util import rand_unitary
    rng = np.random.default_rng(14)
    tol = TOL[np.dtype(dtype)]

    def controlled(U, nc):
        k = int(np.log2(U.shape[0]))
        M = np.eye(2 ** (k + nc), dtype=np.complex128)
        M[-2 ** k:, -2 ** k:] = U
        return M

    for n, nc in ((3, 1), (4, 1), (4, 2), (5, 3), (13, 3), (13, 1)):
        vec = rand_state_vec(rng, n)
        M = controlled(rand_unitary(rng, 2), nc)
        qi = [int(q) for q in rng.permutation(n)[:2 + nc]]
        got = qc.Operator.from_matrix(M)(_state(vec, dtype), qubit_indices=qi)
        want = orc.apply_matrix_flat(M, vec, [n - 1 - q for q in qi], renorm=True)
        assert rel_err(got.to_column_vector(), want) < tol, (n, nc, qi)
