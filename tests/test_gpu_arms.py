"""Every arm of the register-op interpreter (csrc/gen_regops.py) against the oracle.

Random fused passes are built from gate kinds chosen to reach each arm family -- H (HAD), real matrices
(REAL / REALU), RX / Y (IMAG / IMAGU), arbitrary and non-unitary 2x2 (GEN), the same under controls that fall
on register bits (CX, GENM) or elsewhere (the "_P" entries), phases and signs on one or two register bits
(PH1 / SGN1 / PH2 / SGN2, plain and predicated), thread-uniform phases (PHT), general diagonal patterns
(PHM) and dense two-qubit matrices on a pair of register bits (GEN2, plain and predicated).
`qcs_plan_stats` reports which arms each pass uses; the test asserts that the union covers the whole jump
table for the amplitude type, so no arm can be wrong without a parity failure here."""
import ctypes

import numpy as np
import pytest

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import _gates, _lib
from qcircuits_b200._device import make_ops
from tests.util import TOL, check_close, rand_state_vec, rand_unitary, rel_err

pytestmark = pytest.mark.gpu

# (family, first arm id, members per register position / pair) -- the generated table of csrc/gen_regops.py
FAMILIES = [('HAD', 0, 'j'), ('REAL', 4, 'j'), ('REAL_P', 8, 'j'), ('REALU', 12, 'j'), ('IMAG', 16, 'j'),
            ('IMAG_P', 20, 'j'), ('IMAGU', 24, 'j'), ('GEN', 28, 'j'), ('GEN_P', 32, 'j'), ('GENM', 36, 'j'),
            ('CX', 40, 'jc'), ('PH1', 52, 'j'), ('PH1_P', 56, 'j'), ('SGN1', 60, 'j'), ('SGN1_P', 64, 'j'),
            ('PH2', 68, 'p'), ('PH2_P', 74, 'p'), ('SGN2', 80, 'p'), ('SGN2_P', 86, 'p'), ('PHT', 92, '1'),
            ('PHM', 93, '1'), ('END', 94, '1'), ('GEN2', 95, 'p'), ('GEN2_P', 101, 'p')]
PAIRS = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def expected_arms(rb):
    want = set()
    for _, base, kind in FAMILIES:
        if kind == 'j':
            want |= {base + j for j in range(rb)}
        elif kind == 'jc':
            want |= {base + j * 3 + (cb if cb < j else cb - 1) for j in range(rb) for cb in range(rb) if cb != j}
        elif kind == 'p':
            want |= {base + i for i, (a, b) in enumerate(PAIRS) if b < rb}
        else:
            want.add(base)
    return want


def _ctrl(U, ncontrols=1):
    k = int(round(np.log2(U.shape[0])))
    d = 2 ** (k + ncontrols)
    M = np.eye(d, dtype=np.complex128)
    M[d - 2 ** k:, d - 2 ** k:] = U
    return M


def _rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


def _ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def random_gate(rng):
    """(matrix, number of qubits)"""
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.diag([1.0 + 0j, -1.0])
    ph = lambda: np.diag([1.0, np.exp(1j * rng.uniform(0, 2 * np.pi))])
    kind = int(rng.integers(28))
    if kind == 0: return H
    if kind == 1: return X
    if kind == 2: return rng.normal(size=(2, 2)).astype(np.complex128)              # real, non-unitary: REAL
    if kind == 3: return _ry(rng.uniform(-1.5, 1.5))                                  # REALU
    if kind == 4: return _ry(rng.uniform(2.6, 3.6))                                   # |tan| > 2: REAL
    if kind == 5: return _rx(rng.uniform(-1.5, 1.5))                                  # IMAGU
    if kind == 6: return _rx(rng.uniform(2.6, 3.6)) if rng.integers(2) else Y         # IMAG
    if kind == 7: return rand_unitary(rng, 1)                                          # GEN
    if kind == 8: return rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))        # GEN, non-unitary
    if kind == 9: return _ctrl(X)                                                      # CX / REAL_P
    if kind == 10: return _ctrl(_rx(rng.uniform(0, 6)))                                # IMAG_P / GENM
    if kind == 11: return _ctrl(rand_unitary(rng, 1))                                  # GEN_P / GENM
    if kind == 12: return _ctrl(X, 2)                                                  # Toffoli
    if kind == 13: return _ctrl(rand_unitary(rng, 1), 2)
    if kind == 14: return np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, size=2)))      # RZ-like: PH1 / PHT
    if kind == 15: return Z                                                            # SGN1 / PHT
    if kind == 16: return _ctrl(Z)                                                     # CZ: SGN2 / SGN1_P / PHT
    if kind == 17: return _ctrl(ph())                                                  # PH2 / PH1_P / PHT
    if kind == 18: return _ctrl(Z, 2)                                                  # CCZ: SGN2_P / PHM
    if kind == 19: return _ctrl(ph(), 2)                                               # PH2_P / PHM
    if kind == 20: return np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, size=4)))      # 2-qubit diagonal: PHM ...
    if kind == 21: return np.diag(rng.uniform(0.5, 1.5, size=2)).astype(np.complex128) # real non-unitary diagonal
    if kind == 22: return _ctrl(np.diag([1.0 + 0j, rng.uniform(0.5, 1.5)]))            # controlled real factor
    if kind == 23: return _ctrl(H)                                                     # REAL_P / GENM
    if kind == 24: return rand_unitary(rng, 2)                                         # GEN2 (or the shared-memory dense op)
    if kind == 25: return _ctrl(rand_unitary(rng, 2))                                  # GEN2_P
    if kind == 26: return rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))       # GEN2, non-unitary
    return _ctrl(rand_unitary(rng, 2), 2)                                              # GEN2_P, two controls


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_every_interpreter_arm_matches_the_oracle(dtype, monkeypatch):
    monkeypatch.setenv('QCS_TC', '0')      # the register planner; tensor-core sweeps: tests/test_gpu_tc.py
    rng = np.random.default_rng(4242)
    n = 14
    lib = _lib.load()
    rb = 4 if dtype == np.complex64 else 3
    want_arms = expected_arms(rb)
    seen = set()
    stats = (ctypes.c_int32 * 128)()
    worst = 0.0
    for trial in range(400):
        vec = rand_state_vec(rng, n)
        ref = vec.copy()
        plans = []
        # concentrate the gates on a few bits so that several ops share a sweep's register bits
        hot = [int(b) for b in rng.permutation(n)[:int(rng.integers(3, 7))]]
        for _ in range(int(rng.integers(4, 20))):
            M = random_gate(rng)
            k = int(round(np.log2(M.shape[0])))
            pool = hot if (rng.integers(4) and len(hot) >= k) else list(range(n))
            bits = [int(b) for b in rng.permutation(pool)[:k]]
            plans.append((_gates.analyze(M), bits, False))
            ref = orc.apply_matrix_flat(M, ref, bits, renorm=False)
        ops, keep = make_ops(plans)
        code = _lib.dtype_code(dtype)
        rc = lib.qcs_plan_stats(n, code, ops, len(ops), stats)
        if rc != 0:          # (more high dense bits than one tile holds: not a pass the scheduler would build)
            continue
        seen |= {a for a in range(112) if stats[8 + a]}
        s = qc.State.from_column_vector(vec, dtype=dtype)
        out = s._dv.apply_ops(ops, keep)
        got = out.to_host()                          # normalised by the fused squared norm
        ref_n = ref / np.linalg.norm(ref)
        err = check_close(got, ref_n, dtype, 'test_every_interpreter_arm_matches_the_oracle')
        worst = max(worst, err)
        if seen >= want_arms and trial >= 60:
            break
    missing = sorted(want_arms - seen)
    assert not missing, 'interpreter arms never exercised: %s' % missing
