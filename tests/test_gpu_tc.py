"""Tensor-core sweeps (tile_kernel_tc: tcgen05.mma kind::f16, two-term fp16 split of both operands, accumulators in
tensor memory) against the oracle.  complex64 only; the bar is the north star's 1e-5 relative to the largest amplitude."""
import ctypes

import numpy as np
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import _gates, _lib
from qcircuits_b200._device import make_ops
from tests.util import rand_state_vec, rand_unitary, rel_err

pytestmark = pytest.mark.gpu

H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
CNOT = np.eye(4, dtype=np.complex128)[[0, 1, 3, 2]]
CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
TOL = 1e-5


def _rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


def _rz(t):
    return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])


def _ctrl(U):
    d = 2 * U.shape[0]
    M = np.eye(d, dtype=np.complex128)
    M[d // 2:, d // 2:] = U
    return M


def _run(n, vec, gates):
    """one fused pass on the device + the same gates through the oracle; returns (got, want, stats)"""
    lib = _lib.load()
    plans = [(_gates.analyze(M), bits, False) for M, bits in gates]
    ops, keep = make_ops(plans)
    stats = (ctypes.c_int32 * 128)()
    _lib.check(lib.qcs_plan_stats(n, _lib.QCS_C64, ops, len(ops), stats))
    s = qc.State.from_column_vector(vec, dtype=np.complex64)
    got = s._dv.apply_ops(ops, keep).to_host()
    want = vec
    for M, bits in gates:
        want = orc.apply_matrix_flat(M, want, bits, renorm=False)
    return got, want / np.linalg.norm(want), list(stats)


def test_one_block_on_every_kind_of_bit_set():
    """a run of gates inside five tile bits = one dense block; low, high and scattered bit sets"""
    rng = np.random.default_rng(1)
    worst = 0.0
    for n, block in ((12, [0, 1, 2, 3, 4]), (12, [7, 8, 9, 10, 11]), (14, [0, 3, 5, 9, 11]), (16, [1, 2, 6, 10, 11]),
                     (20, [0, 1, 2, 10, 11]), (20, [4, 5, 6, 7, 8]), (13, [2, 4, 6, 8, 10])):
        vec = rand_state_vec(rng, n)
        gates = []
        for _ in range(16):
            k = int(rng.integers(1, 4))
            bits = [int(b) for b in rng.permutation(block)[:k]]
            gates.append((rand_unitary(rng, k), bits))
        got, want, st = _run(n, vec, gates)
        assert st[7] & 255 == 1 and st[7] >> 8 == len(gates), st[:8]
        err = rel_err(got, want)
        worst = max(worst, err)
        assert err < TOL, (n, block, err)
    print('worst error / tolerance: %.3f' % (worst / TOL))


def test_blocks_with_register_ops_before_them_and_several_sweeps():
    """layered-circuit-like passes: H / RX / RZ on every tile bit, CNOT / CZ bricks, controls and diagonal gates that
    reach outside the tile; several tensor-core sweeps and plain register sweeps in one pass"""
    rng = np.random.default_rng(2)
    worst = 0.0
    for trial in range(30):
        n = int(rng.choice([12, 13, 16, 18, 21]))
        tile = list(range(10)) + ([int(b) for b in rng.permutation(np.arange(10, n))[:2]] if n > 11 else [10, 11][:n - 10])
        vec = rand_state_vec(rng, n)
        gates = []
        for _ in range(int(rng.integers(8, 90))):
            kind = int(rng.integers(9))
            if kind <= 2:
                gates.append(((H, _rx(rng.uniform(0, 6)), _rz(rng.uniform(0, 6)))[kind], [int(rng.choice(tile))]))
            elif kind == 3:
                a = int(rng.integers(len(tile) - 1))
                gates.append((CNOT, [tile[a], tile[a + 1]] if rng.integers(2) else [tile[a + 1], tile[a]]))
            elif kind == 4:
                gates.append((CZ, [int(b) for b in rng.permutation(n)[:2]]))                 # diagonal: anywhere
            elif kind == 5:
                tgt = int(rng.choice(tile))
                ctl = int(rng.choice([b for b in range(n) if b != tgt]))                      # control anywhere
                gates.append((_ctrl(rand_unitary(rng, 1)), [ctl, tgt]))
            elif kind == 6:
                gates.append((rand_unitary(rng, 2), [int(b) for b in rng.permutation(tile)[:2]]))
            elif kind == 7:
                gates.append((rand_unitary(rng, 3), [int(b) for b in rng.permutation(tile)[:3]]))
            else:
                gates.append((_ctrl(_rz(rng.uniform(0, 6))), [int(b) for b in rng.permutation(n)[:2]]))
        got, want, st = _run(n, vec, gates)
        err = rel_err(got, want)
        worst = max(worst, err)
        assert err < TOL, (trial, n, err, st[:8])
    print('worst error / tolerance: %.3f' % (worst / TOL))


def test_tensor_core_pass_equals_register_pass(monkeypatch):
    """the same pass through the register-sweep kernel (QCS_TC=0) and the tensor-core kernel"""
    rng = np.random.default_rng(3)
    n = 18
    vec = rand_state_vec(rng, n)
    gates = []
    for layer in range(6):
        for b in range(12):
            gates.append(((H, _rx(rng.uniform(0, 6)), _rz(rng.uniform(0, 6)))[int(rng.integers(3))], [b]))
        for b in range(layer % 2, 11, 2):
            gates.append((CNOT if layer % 4 < 2 else CZ, [b + 1, b]))
    gates = gates[:120]
    got_tc, want, st = _run(n, vec, gates)
    assert st[7] & 255 >= 2
    monkeypatch.setenv('QCS_TC', '0')
    got_reg, _, st0 = _run(n, vec, gates)
    assert st0[7] == 0
    assert rel_err(got_tc, want) < TOL and rel_err(got_reg, want) < TOL
    assert rel_err(got_tc, got_reg) < TOL


def test_non_unitary_ops_stay_off_the_tensor_cores_and_lazy_norm():
    """the fp16 operands are ranged with the norm of the state, which only unitaries keep: a pass with a projector or an
    arbitrary matrix runs on the register-sweep kernel (fp32) and renormalises as in the reference
    (operators.py:275-276) through the squared norm the pass emits; the same pass with unitaries takes the tensor cores"""
    rng = np.random.default_rng(4)
    n = 15
    vec = rand_state_vec(rng, n)
    gates = [(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)), [3]), (CNOT, [3, 4]),
             (np.array([[1, 0], [0, 0]], dtype=np.complex128), [5]), (rand_unitary(rng, 2), [4, 6]), (H, [7])]
    got, want, st = _run(n, vec, gates)
    assert st[7] & 255 == 0
    assert rel_err(got, want) < TOL
    assert abs(np.vdot(got, got) - 1.0) < 1e-5
    gates[0] = (rand_unitary(rng, 1), [3])
    gates[2] = (_rz(0.7), [5])
    got, want, st = _run(n, vec, gates)
    assert st[7] & 255 >= 1
    assert rel_err(got, want) < TOL
    assert abs(np.vdot(got, got) - 1.0) < 1e-5


def test_ranging_of_the_fp16_operands():
    """states whose norm is far from one (the caller gave no lazy norm: the library measures it), a basis state (one
    amplitude carries everything) and a state that is tiny almost everywhere: the error bar stays relative to the
    largest amplitude"""
    rng = np.random.default_rng(9)
    n = 16
    gates = []
    for _ in range(24):
        a, b = (int(v) for v in rng.choice(5, size=2, replace=False))
        gates.append((rand_unitary(rng, 2), [3 + a, 3 + b]))
    lib = _lib.load()
    plans = [(_gates.analyze(M), bits, False) for M, bits in gates]
    ops, keep = make_ops(plans)
    stats = (ctypes.c_int32 * 128)()
    _lib.check(lib.qcs_plan_stats(n, _lib.QCS_C64, ops, len(ops), stats))
    assert stats[7] & 255 >= 1
    t = _lib.torch()
    basis = np.zeros(2 ** n, dtype=np.complex128)
    basis[12345] = 1.0
    spiky = rand_state_vec(rng, n) * 1e-6
    spiky[777] = 0.8
    for name, vec in (('x 1e-12', rand_state_vec(rng, n) * 1e-12), ('x 3e8', rand_state_vec(rng, n) * 3e8),
                      ('basis state', basis), ('one large amplitude', spiky)):
        want = vec.copy()
        for M, bits in gates:
            want = orc.apply_matrix_flat(M, want, bits, renorm=False)
        buf = t.from_numpy(vec.astype(np.complex64).view(np.float32)).cuda()
        ws = _lib.workspace()
        _lib.check(lib.qcs_apply_fused(buf.data_ptr(), buf.data_ptr(), n, _lib.QCS_C64, ops, len(ops), None, None,
                                       ws.data_ptr(), _lib.stream_ptr()))
        got = buf.cpu().numpy().view(np.complex64).astype(np.complex128)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        print('%s: error / tolerance %.3f' % (name, err / TOL))
        assert err < TOL, name


def test_bounded_entry_point_ranges_with_the_callers_bound():
    """qcs_apply_fused_bounded (include/qcs.h): a caller that keeps no lazy norm hands over a bound of the squared norm --
    the sharded engine passes the norm of the WHOLE state to every shard.  Exact, loose (x 64) and per-shard-small cases
    (a shard that holds 2^-12 of the norm) stay within the tolerance relative to the largest amplitude of the state."""
    rng = np.random.default_rng(13)
    n = 16
    gates = []
    for _ in range(30):
        a, b = (int(v) for v in rng.choice(6, size=2, replace=False))
        gates.append((rand_unitary(rng, 2), [2 + a, 2 + b]))
    lib = _lib.load()
    ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
    stats = (ctypes.c_int32 * 128)()
    _lib.check(lib.qcs_plan_stats(n, _lib.QCS_C64, ops, len(ops), stats))
    assert stats[7] & 255 >= 1
    t = _lib.torch()
    vec = rand_state_vec(rng, n)
    for name, scale, bound2 in (('exact bound', 1.0, 1.0), ('loose bound', 1.0, 64.0), ('small shard', 2.0 ** -6, 1.0)):
        v = vec * scale
        want = v.copy()
        for M, bits in gates:
            want = orc.apply_matrix_flat(M, want, bits, renorm=False)
        buf = t.from_numpy(v.astype(np.complex64).view(np.float32)).cuda()
        bound = t.tensor([bound2, 0.0], dtype=t.float64, device='cuda')
        _lib.check(lib.qcs_apply_fused_bounded(buf.data_ptr(), buf.data_ptr(), n, _lib.QCS_C64, ops, len(ops), None, None,
                                               bound.data_ptr(), _lib.workspace().data_ptr(), _lib.stream_ptr()))
        got = buf.cpu().numpy().view(np.complex64).astype(np.complex128)
        # the bar of the sharded state: relative to the largest amplitude of the WHOLE state (norm 1 over 2^16 amplitudes)
        err = np.max(np.abs(got - want)) / np.max(np.abs(vec))
        print('%s: error / tolerance %.4f' % (name, err / TOL))
        assert err < TOL, name


def test_both_group_layouts_of_the_kernel(monkeypatch):
    """four groups of 128 threads (default: a thread owns a whole row, two halves) and two groups of 256 (a 2-stage
    ring each) run the same plan: both against the oracle, and equal to each other up to the order in which the
    squared-norm partials are added (the lazy norm differs in its last bits)"""
    rng = np.random.default_rng(21)
    n = 17
    vec = rand_state_vec(rng, n)
    gates = []
    for block in ([0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [2, 4, 6, 8, 10], [7, 8, 9, 10, 11]):
        for _ in range(6):
            a, b = (int(v) for v in rng.choice(block, size=2, replace=False))
            gates.append((rand_unitary(rng, 2), [a, b]))
        gates.append((_ctrl(_rx(0.4)), [14, block[1]]))          # a control outside the tile: register op
        gates.append((_rz(0.9), [16]))
    results = {}
    for groups in ('4', '2'):
        monkeypatch.setenv('QCS_TC_GROUPS', groups)
        got, want, st = _run(n, vec, gates)
        assert st[7] & 255 >= 3
        assert rel_err(got, want) < TOL
        results[groups] = got
    assert rel_err(results['4'], results['2']) < 2e-7


def test_layered_circuit_through_the_scheduler():
    """Circuit.run on the headline workload's generator at 20 qubits (depth 40, 1 180 gates) vs the oracle"""
    from qcircuits_b200.workloads import layered_circuit, layered_gate_list, operator_for
    n, depth = 20, 40
    circ = layered_circuit(n, depth, seed=1234)
    out = circ.run(qc.zeros(n, dtype=np.complex64))
    vec = np.zeros(2 ** n, dtype=np.complex128)
    vec[0] = 1.0
    cache = {}
    for name, theta, qs in layered_gate_list(n, depth, 1234):
        M = cache.get((name, theta))
        if M is None:
            M = cache[(name, theta)] = operator_for(name, theta).to_matrix()
        vec = orc.apply_matrix_flat(M, vec, [n - 1 - q for q in qs], renorm=False)
    err = rel_err(out.to_column_vector(), vec / np.linalg.norm(vec))
    print('layered 20 qubits depth 40: error / tolerance %.3f' % (err / TOL))
    assert err < TOL
    lib = _lib.load()
    stats = (ctypes.c_int32 * 128)()
    ntc = 0
    for item in circ._compiled(np.complex64, True, 128):
        if item[0] == 'tile':
            _lib.check(lib.qcs_plan_stats(n, _lib.QCS_C64, item[1], len(item[1]), stats))
            ntc += stats[7] & 255
    assert ntc > 0


def test_thirty_qubit_circuit_and_its_inverse_is_the_identity():
    """full size: 30 qubits complex64 (8 GiB), circuit followed by its inverse returns |0...0>"""
    from qcircuits_b200.circuit import Circuit
    from qcircuits_b200.workloads import layered_gate_list, operator_for
    t = _lib.torch()
    if t.cuda.mem_get_info()[0] < 20 * 2 ** 30:
        pytest.skip('needs 20 GiB of device memory')
    n, depth = 30, 6
    gl = layered_gate_list(n, depth, 99)
    c = Circuit(n)
    for name, theta, qs in gl:
        c.append(operator_for(name, theta), list(qs))
    for name, theta, qs in reversed(gl):
        c.append(operator_for(name, theta).adj, list(qs))
    out = c.run(qc.zeros(n, dtype=np.complex64))
    p = out._dv.marginals(list(range(n - 1, n - 17, -1)))      # the 16 most significant bits
    assert abs(p[0] - 1.0) < 2e-5
    p2 = out._dv.marginals(list(range(15, -1, -1)))
    assert abs(p2[0] - 1.0) < 2e-5
