"""CPU-only: the tensor-core planner (qcs_plan_blocks plans a pass without launching it).

A tensor-core sweep multiplies a run of gates into one dense 2^5 x 2^5 block; the host builds the two fp16 operand
images the kernel copies to shared memory.  Here the images are decoded and compared with the product of the same
gates computed by the oracle (oracle/qc_oracle.py, apply_matrix_flat) on five qubits."""
import ctypes

import numpy as np

import qcircuits_b200  # noqa: F401
from qcircuits_b200 import _gates, _lib
from qcircuits_b200._device import make_ops
from tests.util import rand_unitary
from oracle import qc_oracle as orc

H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
CNOT = np.eye(4, dtype=np.complex128)[[0, 1, 3, 2]]
CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
TOFFOLI = np.eye(8, dtype=np.complex128)[[0, 1, 2, 3, 4, 5, 7, 6]]


def plan_blocks(n, gates):
    lib = _lib.load()
    ops, keep = make_ops([(_gates.analyze(M), bits, False) for M, bits in gates])
    stats = (ctypes.c_int32 * 128)()
    images = np.zeros(8 * 8192, dtype=np.uint8)
    info = np.full(8 * 8, -1, dtype=np.int32)
    rc = lib.qcs_plan_blocks(n, ops, len(ops), stats, images.ctypes.data, info.ctypes.data)
    return rc, list(stats), images, info.reshape(8, 8)


def decode(images, slot):
    """(hi + lo) complex 32 x 32 block of a slot, from the fp16 K-major no-swizzle operand images (8 KB per slot: element
    (row n, column k) of an image at (n % 8) * 16 + (n // 8) * 512 + (k // 8) * 128 + (k % 8) * 2 bytes)."""
    img = images[slot * 8192:(slot + 1) * 8192].view(np.float16).astype(np.float64)
    parts = []
    for half in range(2):
        M = np.zeros((64, 32))
        for r in range(64):
            for c in range(32):
                M[r, c] = img[half * 2048 + ((r % 8) * 16 + (r // 8) * 512 + (c // 8) * 128 + (c % 8) * 2) // 2]
        parts.append(M)
    hi, lo = parts
    full = hi + lo
    return full[:32] + 1j * full[32:], hi[:32] + 1j * hi[32:]


def oracle_block(gates, block_bits):
    """Product of the gates as a matrix over the block index (bit j of the index = flat bit block_bits[j])."""
    pos = {b: j for j, b in enumerate(block_bits)}
    U = np.eye(32, dtype=np.complex128)
    for M, bits in gates:
        cols = [orc.apply_matrix_flat(M, U[:, c].copy(), [pos[b] for b in bits], renorm=False) for c in range(32)]
        U = np.stack(cols, axis=1)
    return U


def test_block_matrix_of_a_five_bit_run_matches_the_oracle():
    rng = np.random.default_rng(11)
    n = 16
    for trial in range(12):
        block = sorted(rng.choice(12, size=5, replace=False).tolist())
        gates = []
        for _ in range(int(rng.integers(6, 20))):
            kind = int(rng.integers(7))
            if kind == 0:
                gates.append((H, [int(rng.choice(block))]))
            elif kind == 1:
                gates.append((rand_unitary(rng, 1), [int(rng.choice(block))]))
            elif kind == 2:
                gates.append((CNOT, [int(b) for b in rng.choice(block, size=2, replace=False)]))
            elif kind == 3:
                gates.append((CZ, [int(b) for b in rng.choice(block, size=2, replace=False)]))
            elif kind == 4:
                gates.append((rand_unitary(rng, 2), [int(b) for b in rng.choice(block, size=2, replace=False)]))
            elif kind == 5:
                gates.append((TOFFOLI, [int(b) for b in rng.choice(block, size=3, replace=False)]))
            else:
                d = np.exp(2j * np.pi * rng.uniform(size=8))
                gates.append((np.diag(d), [int(b) for b in rng.choice(block, size=3, replace=False)]))
        rc, stats, images, info = plan_blocks(n, gates)
        assert rc == 0
        assert stats[7] & 255 == 1 and stats[7] >> 8 == len(gates)      # one block holds every op
        assert stats[1] == 1 and info[0][5] == 0                          # one sweep, no register ops before it
        block_bits = [int(b) for b in info[0][:5]]
        assert sorted(block_bits) == block
        assert block_bits[:4] == sorted(block_bits[:4])                   # register bits ascending, thread bit last
        U, U_hi = decode(images, 0)
        want = oracle_block(gates, block_bits)
        assert np.max(np.abs(U - want)) < 3e-7                            # hi + lo carries ~21 bits
        assert np.max(np.abs(U_hi - want)) < 6e-4                         # the hi image alone is fp16


def test_ops_outside_the_block_run_as_register_ops_or_wait():
    rng = np.random.default_rng(5)
    n = 20
    # bits 0..4 form the block; CZ(4, 5) and the RX on bit 7 are diagonal / outside: CZ runs as a register op of
    # the same sweep, the gate on bit 7 needs another sweep
    rx = lambda t: np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]])
    gates = [(H, [0]), (CNOT, [0, 1]), (rand_unitary(rng, 2), [2, 3]), (CZ, [4, 5]), (H, [4]), (CNOT, [3, 4]), (rx(0.3), [7])]
    rc, stats, images, info = plan_blocks(n, gates)
    assert rc == 0
    ntc, nblk = stats[7] & 255, stats[7] >> 8
    assert ntc >= 1 and nblk >= 5
    assert stats[1] == 2


def test_a_pass_without_any_dense_block_is_left_to_the_register_kernel():
    # controls outside the tile: nothing can join a block
    gates = [(CNOT, [23, 3]), (CZ, [22, 3]), (CNOT, [21, 2])]
    rc, stats, images, info = plan_blocks(24, gates)
    assert rc == -3


def test_a_non_unitary_op_keeps_the_pass_off_the_tensor_cores():
    # the fp16 operands are ranged with the norm of the state, which only unitaries keep
    rng = np.random.default_rng(2)
    proj = np.array([[1, 0], [0, 0]], dtype=complex)
    gates = [(H, [0]), (CNOT, [0, 1]), (rand_unitary(rng, 2), [2, 3]), (proj, [1]), (H, [4])]
    rc, stats, images, info = plan_blocks(20, gates)
    assert rc == -3
    gates[3] = (2.0 * np.eye(2, dtype=complex), [1])
    assert plan_blocks(20, gates)[0] == -3
    gates[3] = (np.diag([1.0, 1j]), [1])
    assert plan_blocks(20, gates)[0] == 0


def test_headline_circuit_goes_mostly_through_blocks():
    from qcircuits_b200.workloads import layered_circuit
    lib = _lib.load()
    circ = layered_circuit(30, 40, seed=1234)
    comp = circ._compiled(np.complex64, True, 128)
    stats = (ctypes.c_int32 * 128)()
    in_blocks = total = 0
    for item in comp:
        _lib.check(lib.qcs_plan_stats(30, 0, item[1], len(item[1]), stats))
        assert (stats[7] & 255) <= 8
        in_blocks += stats[7] >> 8
        total += len(item[1])
    assert in_blocks > 0.7 * total
