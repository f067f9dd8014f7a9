"""CPU tests of the fusion scheduler: a schedule is a reordering of the gate list that only commutes
gates past each other when that is exact, and every pass fits one tile."""
import numpy as np
import pytest
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200.circuit import Circuit
from tests.util import rand_state_vec, rand_unitary


def _layered(n, depth, seed=1234):
    mk = {'H': lambda t: qc.Hadamard(), 'RX': qc.RotationX, 'RZ': qc.RotationZ, 'CNOT': lambda t: qc.CNOT(),
          'CZ': lambda t: qc.ControlledU(qc.PauliZ())}
    c = Circuit(n)
    for name, theta, qs in orc.layered_circuit(n, depth, seed):
        c.append(mk[name](theta), list(qs))
    return c


def _run_oracle(circ, order, vec):
    n = circ.n
    for i in order:
        g = circ.gates[i]
        vec = orc.apply_matrix_flat(g.plan.full, vec, g.bits)
    return vec


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_schedule_is_a_valid_reordering(dtype):
    rng = np.random.default_rng(5)
    n = 14
    c = _layered(n, 6)
    # a few awkward extras: wide gate, 3-qubit dense on high bits, reversed CNOT, swaps
    c.append(qc.Operator.from_matrix(rand_unitary(rng, 5)), [0, 3, 6, 9, 12])
    c.append(qc.Operator.from_matrix(rand_unitary(rng, 3)), [0, 1, 2])
    c.append(qc.CNOT(), [7, 2])
    c.append(qc.Swap(), [0, 13])
    c.append(qc.Toffoli(), [1, 12, 5])
    for name, theta, qs in orc.layered_circuit(n, 2, seed=9):
        c.append({'H': qc.Hadamard(), 'RX': qc.RotationX(theta), 'RZ': qc.RotationZ(theta), 'CNOT': qc.CNOT(),
                  'CZ': qc.ControlledU(qc.PauliZ())}[name], list(qs))
    passes = c.schedule(dtype, fuse=True, max_ops=16)
    order = [i for p in passes for i in p]
    assert sorted(order) == list(range(len(c)))
    assert all(len(p) <= 16 for p in passes)
    assert len(passes) < len(c) / 2                      # fusion actually happens
    vec = rand_state_vec(rng, n)
    want = _run_oracle(c, range(len(c)), vec)
    got = _run_oracle(c, order, vec)
    assert np.max(np.abs(got - want)) < 1e-12
    # every pass fits a tile: for some contiguous low run L (the engine shrinks it down from `low`), the
    # dense targets above bit L number at most tile - L
    low, tile = (10, 12) if dtype == np.complex64 else (9, 11)
    for p in passes:
        if len(p) == 1:
            continue
        nd = 0
        for i in p:
            nd |= c.gates[i].nd_mask
        assert any(bin(nd >> L).count('1') <= tile - L for L in range(low - 6, low + 1))


def test_headline_circuit_fuses_well():
    c = _layered(30, 40)
    assert len(c) == 1780
    passes = c.schedule(np.complex64, fuse=True, max_ops=24)
    assert sorted(i for p in passes for i in p) == list(range(1780))
    assert len(passes) < 500
    unfused = c.schedule(np.complex64, fuse=False)
    assert len(unfused) == 1780


def test_single_qubit_merge_preserves_the_unitary():
    from qcircuits_b200.circuit import merge_single_qubit_gates
    rng = np.random.default_rng(8)
    n = 10
    c = _layered(n, 8)
    c.append(qc.PiBy8(), [3])
    c.append(qc.Phase(), [3])
    c.append(qc.Toffoli(), [3, 4, 5])
    c.append(qc.Hadamard(), [3])
    c.append(qc.Hadamard(), [3])          # H H = identity: dropped
    c.append(qc.Operator.from_matrix(rand_unitary(rng, 3)), [1, 3, 8])
    merged = merge_single_qubit_gates(c.gates)
    assert len(merged) < len(c.gates)        # RZ.RZ, S.T, H.H merge; H.RZ stays two cheap ops
    vec = rand_state_vec(rng, n)
    want = vec
    for g in c.gates:
        want = orc.apply_matrix_flat(g.plan.full, want, g.bits, renorm=False)
    got = vec
    for g in merged:
        if g.plan.full is not None:
            got = orc.apply_matrix_flat(g.plan.full, got, g.bits, renorm=False)
    assert np.max(np.abs(got - want)) < 1e-12


def test_swap_expansion_preserves_the_unitary():
    from qcircuits_b200.circuit import expand_swaps
    rng = np.random.default_rng(11)
    n = 9
    c = Circuit(n)
    c.append(qc.Hadamard(), [2])
    c.append(qc.Swap(), [2, 7])
    c.append(qc.RotationX(0.4), [7])
    c.append(qc.Swap(), [0, 1])
    c.append(qc.CNOT(), [1, 8])
    ex = expand_swaps(c.gates)
    assert len(ex) == len(c.gates) + 4
    vec = rand_state_vec(rng, n)
    want, got = vec, vec
    for g in c.gates:
        want = orc.apply_matrix_flat(g.plan.full, want, g.bits, renorm=False)
    for g in ex:
        got = orc.apply_matrix_flat(g.plan.full, got, g.bits, renorm=False)
    assert np.max(np.abs(got - want)) < 1e-14


def _replay(c, dtype, **kw):
    """the scheduled order of the circuit's gates, replayed through the oracle"""
    from qcircuits_b200.circuit import schedule_passes
    rng = np.random.default_rng(1)
    vec = rand_state_vec(rng, c.n)
    want, got = vec, vec
    for g in c.gates:
        want = orc.apply_matrix_flat(g.plan.full, want, g.bits, renorm=False)
    passes = schedule_passes(c.gates, c.n, dtype, **kw)
    assert sorted(i for p in passes for i in p) == list(range(len(c.gates)))
    for p in passes:
        for i in p:
            got = orc.apply_matrix_flat(c.gates[i].plan.full, got, c.gates[i].bits, renorm=False)
    return float(np.max(np.abs(got - want)))


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('tc', ['0', '1'])
def test_tail_trimming_never_reorders_non_commuting_gates(dtype, tc, monkeypatch):
    """A pass that ENDS in a dense op (3-qubit unitary, U_f) used to report the previous register sweep as its
    tail; the trimmed gates then ran after the dense gate.  Scheduled order must equal program order up to
    commuting gates, for every tail_trim."""
    monkeypatch.setenv('QCS_TC', tc)
    rng = np.random.default_rng(21)
    for trial in range(6):
        n = 12
        c = Circuit(n)
        c.append(qc.Hadamard(), [0])
        c.append(qc.RotationX(0.3), [1])
        if trial % 2:
            c.append(qc.U_f(lambda a, b: a ^ b, 3), [0, 1, 2])
        else:
            c.append(qc.Operator.from_matrix(rand_unitary(rng, 3)), [0, 1, 2])
        for _ in range(trial):
            c.append(qc.Hadamard(), [int(rng.integers(n))])
            c.append(qc.CNOT(), [int(q) for q in rng.permutation(n)[:2]])
            c.append(qc.Operator.from_matrix(rand_unitary(rng, 3)), [int(q) for q in rng.permutation(n)[:3]])
        for tail_trim in (0, 4, 8, 12):
            assert _replay(c, dtype, tail_trim=tail_trim) < 1e-12, (trial, tail_trim)


def test_leftover_last_pass_goes_back_into_the_pass_before():
    """the hand-back of a short last sweep has nobody to hand to at the end of the circuit: a last pass of a few
    leftover ops is merged into the pass before it when the planner holds the union (order still exact)"""
    rng = np.random.default_rng(8)
    for n, depth in ((14, 5), (16, 9), (13, 12)):
        c = _layered(n, depth, seed=77 + n)
        passes = c.schedule(np.complex64, fuse=True, max_ops=128)
        order = [i for p in passes for i in p]
        assert sorted(order) == list(range(len(c)))
        vec = rand_state_vec(rng, n)
        assert np.max(np.abs(_run_oracle(c, order, vec) - _run_oracle(c, range(len(c)), vec))) < 1e-12
        if len(passes) >= 2 and len(passes[-1]) <= 8:
            # a tiny last pass survives only if the planner refused the union
            import ctypes
            from qcircuits_b200 import _lib
            from qcircuits_b200._device import make_ops
            merged = passes[-2] + passes[-1]
            ops, keep = make_ops([(c.gates[i].plan, c.gates[i].bits, False) for i in merged])
            stats = (ctypes.c_int32 * 128)()
            assert len(merged) > 128 or _lib.load().qcs_plan_stats(n, _lib.QCS_C64, ops, len(ops), stats) != 0
