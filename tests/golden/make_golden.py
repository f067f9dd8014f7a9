"""Generate golden vectors from the UNMODIFIED reference (grey-area/qcircuits v0.6.0).

Run in the build container only (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``qcircuits`` from /root/reference, runs seeded inputs through the reference's
public API and writes ``tests/golden/golden_*.npz``.  Those files are what pins
``oracle/qc_oracle.py`` (tests/test_oracle_golden.py) and, through the GPU parity
tests, the CUDA path.
"""
import os
import sys
import warnings

import numpy as np
from scipy.stats import unitary_group

warnings.filterwarnings('ignore')
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
import qcircuits as qc  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.qc_oracle import layered_circuit  # noqa: E402  (only the gate-list generator)


def random_state(rng, n):
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    v /= np.linalg.norm(v)
    return qc.State(v.reshape([2] * n))


def random_unitary(rng, k):
    M = unitary_group.rvs(2 ** k, random_state=rng)
    return qc.Operator.from_matrix(M)


def ref_gate(name, theta):
    if name == 'H':
        return qc.Hadamard()
    if name == 'RX':
        return qc.RotationX(theta)
    if name == 'RZ':
        return qc.RotationZ(theta)
    if name == 'CNOT':
        return qc.CNOT()
    if name == 'CZ':
        return qc.ControlledU(qc.PauliZ())
    raise KeyError(name)


def main():
    out = {}

    # 1. layered circuits (SURVEY 8d workload) ------------------------------------------
    for n, depth in [(3, 4), (5, 6), (8, 6), (10, 8), (12, 4)]:
        state = qc.zeros(n)
        for name, theta, qs in layered_circuit(n, depth, seed=1234):
            state = ref_gate(name, theta)(state, qubit_indices=list(qs))
        out['layered_n%d_d%d_amps' % (n, depth)] = state.to_column_vector()
        out['layered_n%d_d%d_probs' % (n, depth)] = state.probabilities.reshape(-1)

    # 2. random unitaries on random qubit subsets (order matters) ------------------------
    rng = np.random.RandomState(7)
    case = 0
    for n in [1, 2, 4, 6, 9]:
        for k in range(1, min(n, 4) + 1):
            for _ in range(2):
                qi = [int(q) for q in rng.permutation(n)[:k]]
                s = random_state(rng, n)
                U = random_unitary(rng, k)
                r = U(s, qubit_indices=qi)
                out['apply%d_in' % case] = s.to_column_vector()
                out['apply%d_op' % case] = U.to_matrix()
                out['apply%d_qi' % case] = np.array(qi)
                out['apply%d_out' % case] = r.to_column_vector()
                case += 1
    out['apply_count'] = np.array(case)

    # 3. non-unitary operators get renormalised ------------------------------------------
    s = random_state(rng, 5)
    P0 = qc.Operator.from_matrix(np.array([[1, 0], [0, 0]]))
    r = P0(s, qubit_indices=[2])
    out['proj_in'] = s.to_column_vector()
    out['proj_out'] = r.to_column_vector()
    A = qc.Operator.from_matrix(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    r = A(s, qubit_indices=[4, 1])
    out['nonunitary_op'] = A.to_matrix()
    out['nonunitary_out'] = r.to_column_vector()

    # 4. measurement with the global NumPy stream ----------------------------------------
    case = 0
    for n, qi, remove in [(1, 0, False), (3, [1], False), (3, [1], True), (5, [3, 0], False),
                          (5, [3, 0], True), (5, None, False), (6, [5, 2, 4], True),
                          (6, 2, True), (7, [0, 1, 2, 3, 4, 5, 6], False), (8, [7, 0, 3], False)]:
        for seed in (0, 1, 2):
            s = random_state(rng, n)
            v_in = s.to_column_vector()
            np.random.seed(1000 + seed)
            u = np.random.random_sample()
            np.random.seed(1000 + seed)
            bits = s.measure(qubit_indices=qi, remove=remove)
            out['meas%d_in' % case] = v_in
            out['meas%d_qi' % case] = np.array(-1 if qi is None else qi)
            out['meas%d_remove' % case] = np.array(remove)
            out['meas%d_u' % case] = np.array(u)
            out['meas%d_bits' % case] = np.array(bits)
            out['meas%d_out' % case] = s.to_column_vector() if s.rank > 0 else np.array([s._t])
            case += 1
    out['meas_count'] = np.array(case)

    # 5. tensor products / scalar arithmetic ---------------------------------------------
    a, b = random_state(rng, 3), random_state(rng, 2)
    out['kron_a'] = a.to_column_vector()
    out['kron_b'] = b.to_column_vector()
    out['kron_ab'] = (a * b).to_column_vector()
    out['kron_pow3'] = (b ** 3).to_column_vector()
    out['add_ab'] = (a + qc.State(a._t[::-1])).to_column_vector()
    out['dot_ab'] = np.array(a.dot(qc.State(a._t[::-1])))

    # 6. density operators ----------------------------------------------------------------
    n = 3
    states = [random_state(rng, n) for _ in range(3)]
    ps = np.array([0.5, 0.3, 0.2])
    rho = qc.DensityOperator.from_ensemble(states, ps)
    out['rho_states'] = np.stack([s.to_column_vector() for s in states])
    out['rho_ps'] = ps
    out['rho_t'] = rho._t.reshape(-1)
    out['rho_probs'] = rho.probabilities.reshape(-1)
    U2 = random_unitary(rng, 2)
    rho2 = U2(rho, qubit_indices=[2, 0])
    out['rho_op'] = U2.to_matrix()
    out['rho_applied'] = rho2._t.reshape(-1)
    H = qc.Hadamard()
    out['rho_h1'] = H(rho, qubit_indices=[1])._t.reshape(-1)
    out['rho_reduced_02'] = rho.reduced_density_operator([0, 2])._t.reshape(-1)
    case = 0
    for qi, remove in [([1], False), ([1], True), ([2, 0], False), ([2, 0], True), (None, False)]:
        r = qc.DensityOperator(rho2._t)
        np.random.seed(2000 + case)
        u = np.random.random_sample()
        np.random.seed(2000 + case)
        bits = r.measure(qubit_indices=qi, remove=remove)
        out['rhomeas%d_qi' % case] = np.array(-1 if qi is None else qi)
        out['rhomeas%d_remove' % case] = np.array(remove)
        out['rhomeas%d_u' % case] = np.array(u)
        out['rhomeas%d_bits' % case] = np.array(bits)
        out['rhomeas%d_out' % case] = np.asarray(r._t).reshape(-1)
        case += 1
    out['rhomeas_count'] = np.array(case)

    # 7. operator composition and Grover (config 1 scaled down) ---------------------------
    A1, A2 = random_unitary(rng, 2), random_unitary(rng, 3)
    out['compose_a'] = A1.to_matrix()
    out['compose_b'] = A2.to_matrix()
    out['compose_ab'] = A1(A2, qubit_indices=[2, 0]).to_matrix()

    d = 5
    answers = np.zeros(2 ** d, dtype=np.int32)
    answers[11] = 1

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
    iters = int(round(np.arccos(np.sqrt(1 / N)) / (2 * np.arcsin(np.sqrt(1 / N)))))
    for _ in range(iters):
        state = Grover(state)
    out['grover_d5_answers'] = answers
    out['grover_d5_op'] = Grover.to_matrix()
    out['grover_d5_amps'] = state.to_column_vector()
    np.random.seed(5)
    out['grover_d5_u'] = np.array(np.random.random_sample())
    np.random.seed(5)
    out['grover_d5_bits'] = np.array(state.measure(qubit_indices=range(d)))

    # 8. gate factory tensors --------------------------------------------------------------
    for name, op in [('I', qc.Identity()), ('X', qc.PauliX()), ('Y', qc.PauliY()), ('Z', qc.PauliZ()),
                     ('H', qc.Hadamard()), ('H3', qc.Hadamard(3)), ('S', qc.Phase()), ('T', qc.PiBy8()),
                     ('RX', qc.RotationX(0.7)), ('RY', qc.RotationY(1.3)), ('RZ', qc.RotationZ(2.1)),
                     ('R', qc.Rotation([0.6, 0.0, 0.8], 0.9)), ('SqrtNot', qc.SqrtNot()), ('CNOT', qc.CNOT()),
                     ('Toffoli', qc.Toffoli()), ('Swap', qc.Swap()), ('SqrtSwap', qc.SqrtSwap()),
                     ('CZ', qc.ControlledU(qc.PauliZ())), ('CCH', qc.ControlledU(qc.ControlledU(qc.Hadamard()))),
                     ('Uf', qc.U_f(lambda a, b: a ^ b, 3))]:
        out['gate_' + name] = op._t.reshape(-1)

    np.savez_compressed(os.path.join(HERE, 'golden_v1.npz'), **out)
    print('wrote', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'golden_v1.npz')), 'bytes')


if __name__ == '__main__':
    main()
