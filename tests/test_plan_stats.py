"""CPU-only: the host planner of the tile engine (qcs_plan_stats plans a pass without launching it)."""
import ctypes

import numpy as np
import pytest

import qcircuits_b200  # noqa: F401  (package import must work without a GPU)
from qcircuits_b200 import _gates, _lib
from qcircuits_b200._device import make_ops

H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
CNOT = np.eye(4, dtype=np.complex128)[[0, 1, 3, 2]]
CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)


@pytest.fixture(autouse=True)
def register_sweeps_only(monkeypatch):
    """These tests pin the REGISTER planner (the interpreter's arm choice); complex64 passes with dense blocks
    otherwise go to the tensor-core planner, which tests/test_tc_plan.py covers."""
    monkeypatch.setenv('QCS_TC', '0')


def _stats(n, dtype, gates):
    lib = _lib.load()
    ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
    st = (ctypes.c_int32 * 128)()
    _lib.check(lib.qcs_plan_stats(n, _lib.dtype_code(dtype), ops, len(ops), st))
    return list(st)


def test_cheapest_forms_are_chosen():
    rx = lambda t: np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]])
    st = _stats(20, np.complex64, [(H, [0]), (rx(0.7), [1]), (CNOT, [0, 1]), (CZ, [0, 1]), (rx(3.0), [2])])
    arms = {a: st[8 + a] for a in range(100) if st[8 + a]}
    # register bits 0, 1, 2 (+ one filler): H -> HAD on position 0, RX -> IMAGU on position 1, CNOT(control bit 0,
    # target bit 1) -> CX, CZ -> SGN2 on the pair (0, 1), RX(3.0) (|tan| > 2) -> IMAG on position 2, then END
    assert arms == {0: 1, 24 + 1: 1, 40 + 1 * 3 + 0: 1, 80: 1, 16 + 2: 1, 94: 1}
    assert st[1] == 1 and st[2] == 6          # one sweep, five ops + END
    assert st[5] == 1                         # deferred real scalars (1/sqrt2, cos) -> real pass-level factor


def test_controls_outside_the_tile_become_predicates():
    st = _stats(24, np.complex64, [(CNOT, [23, 3]), (CZ, [22, 3])])
    arms = {a for a in range(100) if st[8 + a]}
    assert any(8 <= a < 12 for a in arms)      # REAL_P: X under a tile-level predicate
    assert any(64 <= a < 68 for a in arms)     # SGN1_P
    assert st[4] == 2                          # two distinct tile-level predicates


def test_tile_shrinks_its_contiguous_run_for_many_high_targets():
    # five dense targets above bit 10: the planner gives up contiguous low bits (L = 7, H = 5 in a 32 KB tile)
    st = _stats(26, np.complex64, [(H, [b]) for b in (12, 14, 16, 18, 20)])
    assert st[0] == 12
    lib = _lib.load()
    ops, keep = make_ops([(_gates.analyze(H), [b], False) for b in range(10, 26)])
    out = (ctypes.c_int32 * 128)()
    assert lib.qcs_plan_stats(26, 0, ops, len(ops), out) == -3      # QCS_ERR_UNSUPPORTED: too many high bits


def test_groups_the_planner_cannot_hold_are_split():
    """A pass the scheduler accepts by count but whose register-op table would overflow (general 4-qubit
    diagonal gates: 15 phase patterns each) is split until every piece plans."""
    import qcircuits_b200 as qc
    from qcircuits_b200.circuit import Circuit, build_passes, merge_single_qubit_gates, schedule_passes
    rng = np.random.default_rng(3)
    n = 16
    c = Circuit(n)
    for _ in range(60):
        d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=16))
        c.append(qc.Operator.from_matrix(np.diag(d)), [int(q) for q in rng.permutation(n)[:4]])
    gates = merge_single_qubit_gates(c.gates)
    groups = schedule_passes(gates, n, np.complex64, max_ops=128)
    passes = build_passes(gates, groups, n, np.complex64)
    assert sum(cnt for _, _, cnt in passes) == 60
    assert len(passes) > len(groups)                 # 60 x 15 patterns do not fit one 420-entry table
    st = (ctypes.c_int32 * 128)()
    for ops, keep, cnt in passes:
        assert _lib.load().qcs_plan_stats(n, 0, ops, len(ops), st) == 0


def test_controlled_two_qubit_unitary_on_a_tiny_state_plans():
    """Every tile bit of a 3-qubit state is a register bit, so the control of a controlled two-qubit unitary
    cannot stay a thread bit: the op must fall back to the shared-memory dense op instead of stalling the
    planner."""
    rng = np.random.default_rng(4)
    A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    U, _ = np.linalg.qr(A)
    CU = np.eye(8, dtype=np.complex128)
    CU[4:, 4:] = U
    for n in (3, 4, 5, 12, 20):
        for dtype in (np.complex64, np.complex128):
            st = _stats(n, dtype, [(CU, [n - 1, 1, 0]), (H, [0]), (CU, [0, 2, 1])])
            assert st[2] >= 1 or st[3] >= 1          # planned: register ops and / or shared-memory dense ops
            assert st[3] + sum(st[8 + a] for a in range(95, 107)) == 2
