"""CPU tests of the host-side mirror of the reference interface: gate factories, operator algebra and
composition (host NumPy by design), matrix structure analysis, and the C-ABI surface of libqcs.so
(load + every declared symbol; no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import _gates, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gate_factories_match_reference(golden):
    ops = {'I': qc.Identity(), 'X': qc.PauliX(), 'Y': qc.PauliY(), 'Z': qc.PauliZ(), 'H': qc.Hadamard(),
           'H3': qc.Hadamard(3), 'S': qc.Phase(), 'T': qc.PiBy8(), 'RX': qc.RotationX(0.7), 'RY': qc.RotationY(1.3),
           'RZ': qc.RotationZ(2.1), 'R': qc.Rotation([0.6, 0.0, 0.8], 0.9), 'SqrtNot': qc.SqrtNot(),
           'CNOT': qc.CNOT(), 'Toffoli': qc.Toffoli(), 'Swap': qc.Swap(), 'SqrtSwap': qc.SqrtSwap(),
           'CZ': qc.ControlledU(qc.PauliZ()), 'CCH': qc.ControlledU(qc.ControlledU(qc.Hadamard())),
           'Uf': qc.U_f(lambda a, b: a ^ b, 3)}
    for name, op in ops.items():
        assert np.array_equal(op._t.reshape(-1), golden['gate_' + name]), name
        assert op._t.dtype == np.complex128


def test_matrix_round_trip_and_adjoint():
    rng = np.random.default_rng(0)
    for d in (1, 2, 3):
        M = rng.normal(size=(2 ** d, 2 ** d)) + 1j * rng.normal(size=(2 ** d, 2 ** d))
        op = qc.Operator.from_matrix(M)
        assert op.rank == 2 * d and op.shape == (2,) * (2 * d)
        assert np.array_equal(op.to_matrix(), M)
        assert np.array_equal(op.adj.to_matrix(), M.conj().T)
        assert np.array_equal(op._t, orc.op_from_matrix(M))
    with pytest.raises(ValueError):
        qc.Operator.from_matrix(np.zeros((2, 3)))
    with pytest.raises(ValueError):
        qc.Operator.from_matrix(np.zeros((3, 3)))


def test_operator_composition_matches_reference(golden):
    A = qc.Operator.from_matrix(golden['compose_a'])
    B = qc.Operator.from_matrix(golden['compose_b'])
    AB = A(B, qubit_indices=[2, 0])
    assert isinstance(AB, qc.Operator)
    assert np.max(np.abs(AB.to_matrix() - golden['compose_ab'])) < 1e-12
    # full-width composition is a matrix product
    C = qc.Operator.from_matrix(golden['compose_a'])
    assert np.allclose(A(C).to_matrix(), golden['compose_a'] @ golden['compose_a'])


def test_operator_algebra_and_tensor_product():
    X, Z, H = qc.PauliX(), qc.PauliZ(), qc.Hadamard()
    assert np.allclose(((X + Z) / np.sqrt(2)).to_matrix(), H.to_matrix())
    assert np.allclose((2 * X - X).to_matrix(), X.to_matrix())
    assert np.allclose((-X).to_matrix(), -X.to_matrix())
    assert np.allclose((X * Z).to_matrix(), np.kron(X.to_matrix(), Z.to_matrix()))
    assert np.allclose((H ** 3).to_matrix(), np.kron(np.kron(H.to_matrix(), H.to_matrix()), H.to_matrix()))
    op = qc.CNOT() * qc.Hadamard()
    before = op.to_matrix().copy()
    op.swap_qubits(0, 2).swap_qubits(0, 2)
    assert np.array_equal(op.to_matrix(), before)
    op.permute_qubits([2, 0, 1]).permute_qubits([2, 0, 1], inverse=True)
    assert np.array_equal(op.to_matrix(), before)


def test_argument_validation_is_the_references():
    H2 = qc.Hadamard(2)
    I3 = qc.Identity(3)
    with pytest.raises(ValueError):
        I3(H2)                      # operator larger than system
    with pytest.raises(ValueError):
        H2(I3)                      # too-large system without indices
    with pytest.raises(ValueError):
        H2(I3, qubit_indices=[0, 0])
    with pytest.raises(ValueError):
        H2(I3, qubit_indices=[-1, 0])
    with pytest.raises(ValueError):
        H2(I3, qubit_indices=[0, 3])
    with pytest.raises(ValueError):
        H2(I3, qubit_indices=[0])
    with pytest.raises(ValueError):
        qc.Rotation([1, 1, 0], 0.1)
    with pytest.raises(ValueError):
        qc.U_f(lambda: 0, 1)
    with pytest.raises(RuntimeError):
        qc.U_f(lambda a: 2, 2)


def _rebuild(plan):
    """matrix described by a GatePlan, on the original operator qubits"""
    k = plan.k
    red = np.diag(plan.matrix) if plan.diagonal else plan.matrix
    full = np.zeros((2 ** k, 2 ** k), dtype=complex)
    for col in range(2 ** k):
        bits = [(col >> (k - 1 - q)) & 1 for q in range(k)]
        if not all(bits[c] for c in plan.controls):
            full[col, col] = 1
            continue
        x = 0
        for q in plan.targets:
            x = (x << 1) | bits[q]
        for y in range(red.shape[0]):
            out = list(bits)
            for j, q in enumerate(plan.targets):
                out[q] = (y >> (len(plan.targets) - 1 - j)) & 1
            row = 0
            for b in out:
                row = (row << 1) | b
            full[row, col] += red[y, x]
    return full


def test_structure_analysis_is_exact():
    rng = np.random.default_rng(1)
    U = orc.op_to_matrix(orc.rotation_x(0.3))
    cases = [qc.CNOT(), qc.ControlledU(qc.PauliZ()), qc.Toffoli(), qc.Swap(), qc.SqrtSwap(), qc.RotationZ(0.4),
             qc.Hadamard(2), qc.ControlledU(qc.ControlledU(qc.RotationX(0.3))), qc.Identity(3), qc.PiBy8(),
             qc.ControlledU(np.exp(0.3j) * qc.RotationZ(0.7)), qc.U_f(lambda a, b: a & b, 3),
             qc.Operator.from_matrix(rng.normal(size=(8, 8)))]
    for op in cases:
        plan = _gates.analyze(op.to_matrix())
        assert np.array_equal(_rebuild(plan), op.to_matrix())
    p = _gates.analyze(qc.CNOT().to_matrix())
    assert (p.controls, p.targets, p.diagonal) == ([0], [1], False)
    p = _gates.analyze(qc.Toffoli().to_matrix())
    assert (sorted(p.controls), p.targets) == ([0, 1], [2])
    p = _gates.analyze(qc.ControlledU(qc.PauliZ()).to_matrix())
    assert p.diagonal and len(p.targets) == 1
    # a reversed CNOT (control on the second qubit) is recognised too
    p = _gates.analyze(qc.CNOT().swap_qubits(0, 1).to_matrix())
    assert (p.controls, p.targets) == ([1], [0])
    assert U.shape == (2, 2)


def test_libqcs_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'qcs.h')).read()
    declared = set(re.findall(r'\b(qcs_[a-z0-9_]+)\s*\(', header))
    declared.discard('qcs_op')
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    bound = _lib.load()
    assert bound.qcs_version() >= 100
    assert bound.qcs_workspace_bytes() > 0
    low, tile = ctypes.c_int(), ctypes.c_int()
    assert bound.qcs_tile_limits(_lib.QCS_C64, ctypes.byref(low), ctypes.byref(tile)) == 0
    assert (low.value, tile.value) == (10, 12)          # 8 KB preferred contiguous run, 32 KB tiles
    assert bound.qcs_tile_limits(_lib.QCS_C128, ctypes.byref(low), ctypes.byref(tile)) == 0
    assert (low.value, tile.value) == (9, 11)
    assert bound.qcs_tile_limits(7, None, None) < 0 and b'dtype' in bound.qcs_last_error()


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        qc.zeros(3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        qc.State(np.array([1.0, 0.0]))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'qcircuits_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+\S*oracle', text, re.M), f
                assert 'qc_oracle' not in text and 'oracle/' not in text, f


def test_workload_generator_matches_the_oracles():
    from qcircuits_b200.workloads import layered_gate_list
    for n, depth in ((3, 4), (12, 5), (30, 40)):
        assert layered_gate_list(n, depth, 1234) == orc.layered_circuit(n, depth, 1234)


def test_chunked_sampling_equals_the_flat_inverse_cdf():
    """State.measure of many qubits samples hierarchically (state._sample_chunked); the outcome must be the
    index the reference's flat np.random.choice rule gives for the same uniform."""
    from qcircuits_b200.state import _sample_chunked
    rng = np.random.default_rng(12)
    m = 11
    for trial in range(40):
        p = rng.uniform(size=2 ** m) ** 3
        if trial % 3 == 0:
            p[rng.uniform(size=2 ** m) < 0.7] = 0.0          # sparse distributions (basis-like states)
        p /= p.sum()
        t = p.reshape([2] * m)
        order = [int(b) for b in rng.permutation(m)]         # measured bit positions, first = most significant
        axes = [m - 1 - b for b in order]
        flat = np.transpose(t, axes).reshape(-1)
        state = {'t': t.copy()}

        def conditional(done_bits, done_outcome, chunk_bits):
            if done_bits:
                sel = [slice(None)] * m
                for j, b in enumerate(done_bits):
                    sel[m - 1 - b] = (done_outcome >> (len(done_bits) - 1 - j)) & 1
                kept = np.zeros_like(state['t'])
                kept[tuple(sel)] = state['t'][tuple(sel)]
                state['t'] = kept / kept.sum()
            ax = [m - 1 - b for b in chunk_bits]
            rest = tuple(a for a in range(m) if a not in ax)
            return np.transpose(state['t'].sum(axis=rest), np.argsort(np.argsort(ax))).reshape(-1)

        u = float(rng.uniform())
        cdf = np.cumsum(flat)
        cdf /= cdf[-1]
        want = int(np.searchsorted(cdf, u, side='right'))
        got = _sample_chunked(conditional, order, u, 4)
        assert got == want, (trial, got, want)
    # u = None consumes exactly one uniform of the global stream, like np.random.choice
    np.random.seed(5)
    expect_u = np.random.random_sample()
    after = np.random.random_sample()
    np.random.seed(5)
    state = {'t': np.full([2] * 6, 1.0 / 64)}
    m = 6
    got = _sample_chunked(lambda db, do, cb: np.full(2 ** len(cb), 1.0 / 2 ** len(cb)), list(range(6)), None, 3)
    assert got == int(expect_u * 64)
    assert np.random.random_sample() == after


def test_operator_caches_follow_the_tensor():
    """permute_qubits / swap_qubits rewrite Operator._t in place (reference operators.py:125-170); the cached structure
    analysis and device matrices must not survive that (the reference recomputes from _t on every call)."""
    from qcircuits_b200 import _gates
    op = qc.CNOT()
    before = op._plan()
    assert before.controls == [0] and before.targets == [1]
    op.swap_qubits(0, 1)
    after = op._plan()
    assert after is not before and after.controls == [1] and after.targets == [0]
    want = _gates.analyze(op.to_matrix())
    assert after.controls == want.controls and after.targets == want.targets
    tof = qc.Toffoli()
    tof._plan()
    tof.permute_qubits([2, 0, 1])
    assert sorted(tof._plan().controls) == sorted(_gates.analyze(tof.to_matrix()).controls)
    assert tof._plan().targets == _gates.analyze(tof.to_matrix()).targets
    h = qc.Hadamard()
    h._plan()
    h._t = qc.PauliX()._t            # direct assignment
    assert np.array_equal(h._plan().full, qc.PauliX().to_matrix())
    import copy
    c = copy.deepcopy(tof)
    assert np.array_equal(c._t, tof._t) and c._plan().targets == tof._plan().targets
