"""CPU oracle for the qcircuits state-vector hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm (grey-area/qcircuits
v0.6.0).  It is the checker for the CUDA path, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``qcircuits_b200/`` imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified reference
(imported from /root/reference in the build container) on seeded inputs and committed
its outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors, and against the known answers printed in the
reference's ``docs/source/tutorial.rst`` (:184-211, :223-228, :260-279, :505-518).

Two formulations are given for gate application:

* ``apply_operator``      -- the reference's own route: transpose view, ``np.tensordot``,
                             defensive copy, inverse transpose, 4-pass renormalise.
                             Same memory passes as the reference, so it doubles as the
                             CPU baseline ("kind": "port").
* ``apply_matrix_flat``   -- an independent flat-index form (bit ``n-1-q`` of the C-order
                             index is qubit ``q``), used for sizes where the rank-n route
                             is too slow and as a cross-check of the first.

Tensor conventions (reference ``qcircuits/tensors.py:17-18``, ``operators.py:69-73``):
a d-qubit state is a complex128 array of shape ``[2]*d`` (qubit 0 = axis 0 = most
significant bit of the flat index); a d-qubit operator is ``[2]*2d`` with axes
interleaved ``[out0, in0, out1, in1, ...]``.
"""

import numpy as np

C128 = np.complex128


# --------------------------------------------------------------------------------------
# operator <-> matrix            (reference operators.py:37-93)
# --------------------------------------------------------------------------------------
def op_from_matrix(M):
    """2^d x 2^d matrix -> interleaved rank-2d tensor (operators.py:56-73)."""
    M = np.asarray(M, dtype=C128)
    if M.ndim != 2 or M.shape[0] != M.shape[1]:
        raise ValueError('The matrix should be square.')
    d = int(round(np.log2(M.shape[0])))
    if 2 ** d != M.shape[0]:
        raise ValueError('The matrix dimension should be a power of 2.')
    order = np.empty(2 * d, dtype=int)
    order[0::2] = np.arange(d)
    order[1::2] = np.arange(d, 2 * d)
    return np.array(M.reshape([2] * (2 * d)).transpose(order), dtype=C128)


def op_to_matrix(t):
    """interleaved rank-2d tensor -> matrix (operators.py:75-79)."""
    d = t.ndim // 2
    order = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
    return t.transpose(order).reshape(2 ** d, 2 ** d)


def op_adjoint(t):
    """conjugate + swap each (out, in) axis pair (operators.py:109-114)."""
    d = t.ndim
    order = np.empty(d, dtype=int)
    order[0::2] = np.arange(1, d, 2)
    order[1::2] = np.arange(0, d, 2)
    return np.array(np.conj(t).transpose(order), dtype=C128)


# --------------------------------------------------------------------------------------
# gate tensors                   (reference operators.py:328-853)
# --------------------------------------------------------------------------------------
def _power(t1, d):
    t = t1
    for _ in range(d - 1):
        t = np.tensordot(t, t1, axes=0)          # tensors.py:97-99
    return np.array(t, dtype=C128)


def identity(d=1):
    return _power(np.array([[1, 0], [0, 1]], dtype=C128), d)          # operators.py:350


def pauli_x(d=1):
    return _power(np.array([[0, 1], [1, 0]], dtype=C128), d)          # operators.py:378


def pauli_y(d=1):
    return _power(np.array([[0, -1j], [1j, 0]], dtype=C128), d)       # operators.py:405


def pauli_z(d=1):
    return _power(np.array([[1, 0], [0, -1]], dtype=C128), d)         # operators.py:432


def hadamard(d=1):
    # operators.py:459-461  (1/sqrt(2) multiplies the +-1 entries)
    return _power(1 / np.sqrt(2) * np.array([[1, 1], [1, -1]], dtype=C128), d)


def phase(d=1):
    return _power(np.array([[1, 0], [0, 1j]], dtype=C128), d)         # operators.py:488


def pi_by_8(d=1):
    return _power(np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]], dtype=C128), d)  # :515


def rotation(v, theta):
    """cos(t/2) I - i sin(t/2) (vx X + vy Y + vz Z)   (operators.py:547-548)."""
    v = np.asarray(v, dtype=float)
    return np.array(np.cos(theta / 2) * identity()
                    - 1j * np.sin(theta / 2) * (v[0] * pauli_x() + v[1] * pauli_y() + v[2] * pauli_z()),
                    dtype=C128)


def rotation_x(theta):
    return rotation([1., 0., 0.], theta)


def rotation_y(theta):
    return rotation([0., 1., 0.], theta)


def rotation_z(theta):
    return rotation([0., 0., 1.], theta)


def controlled(u):
    """|0><0| (x) I + |1><1| (x) U on d+1 qubits, control first (operators.py:797-806)."""
    d = u.ndim // 2 + 1
    t = np.zeros([2] * (2 * d), dtype=C128)
    t[:, 0, ...] = identity(d)[:, 0, ...]
    t[:, 1, ...] = np.tensordot(identity(), u, axes=0)[:, 1, ...]
    return t


def cnot():
    return controlled(pauli_x())      # same tensor as operators.py:675-682 (checked in tests)


def cz():
    return controlled(pauli_z())


def swap():
    M = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=C128)
    return op_from_matrix(M)          # operators.py:734-741


# --------------------------------------------------------------------------------------
# State arithmetic               (reference state.py:74-97, 187-203)
# --------------------------------------------------------------------------------------
def dot(a, b):
    return np.sum(np.conj(a) * b)                                  # state.py:89


def renormalize(t):
    """four passes, like the reference: conj, multiply, sum, divide (state.py:97)."""
    return t / np.sqrt(np.real(dot(t, t)))


def probabilities(t):
    return np.real(np.conj(t) * t)                                 # state.py:200


def tensor_product(a, b):
    return np.array(np.tensordot(a, b, axes=0), dtype=C128)       # tensors.py:76


# --------------------------------------------------------------------------------------
# gate application               (reference operators.py:225-278)
# --------------------------------------------------------------------------------------
def check_qubit_indices(op_qubits, d, qubit_indices):
    """argument validation with the reference's exception types (operators.py:234-253)."""
    if op_qubits > d:
        raise ValueError('operator larger than system')
    if qubit_indices is None:
        if op_qubits < d:
            raise ValueError('qubit indices required')
        return list(range(d))
    qubit_indices = [int(q) for q in qubit_indices]
    if len(set(qubit_indices)) != len(qubit_indices):
        raise ValueError('repeated qubit index')
    if min(qubit_indices) < 0 or max(qubit_indices) >= d:
        raise ValueError('qubit index out of range')
    if len(qubit_indices) != op_qubits:
        raise ValueError('qubit index count does not match operator')
    return qubit_indices


def apply_operator(op_t, psi_t, qubit_indices=None, renorm=True):
    """Reference route for Operator(State) (operators.py:255-276).

    Operator qubit j acts on state qubit ``qubit_indices[j]``.  The result is a fresh
    array; the inputs are untouched.
    """
    d = psi_t.ndim
    k = op_t.ndim // 2
    qi = check_qubit_indices(k, d, qubit_indices)
    rest = sorted(set(range(d)) - set(qi))
    perm = qi + rest
    arg = np.transpose(psi_t, perm)                                # view  (:259)
    res = np.tensordot(op_t, arg, (list(range(1, 2 * k, 2)), list(range(k))))   # (:261)
    res = np.array(res, dtype=C128)                                # defensive copy (:273)
    res = np.transpose(res, np.argsort(perm))                      # (:274)
    if renorm:
        res = renormalize(res)                                     # (:275-276)
    return res


def apply_matrix_flat(M, vec, bitpos, renorm=True):
    """Independent flat-index form of the same contraction.

    ``vec`` has 2^n entries in C order; ``bitpos[j]`` is the flat bit (0 = least
    significant) that operator qubit j acts on, operator qubit 0 being the most
    significant bit of M's row/column index.  For a state, bitpos[j] = n-1-qubit_indices[j].
    """
    vec = np.asarray(vec, dtype=C128)
    n = int(round(np.log2(vec.size)))
    k = len(bitpos)
    M = np.asarray(M, dtype=C128).reshape(2 ** k, 2 ** k)
    idx = np.arange(vec.size, dtype=np.int64)
    mask = 0
    for b in bitpos:
        mask |= 1 << b
    base = idx[(idx & mask) == 0]                                  # one entry per group
    offs = np.zeros(2 ** k, dtype=np.int64)
    for x in range(2 ** k):
        for j, b in enumerate(bitpos):
            if (x >> (k - 1 - j)) & 1:
                offs[x] |= 1 << b
    gathered = vec[base[None, :] + offs[:, None]]                  # (2^k, groups)
    out = np.empty_like(vec)
    out[base[None, :] + offs[:, None]] = M @ gathered
    if renorm:
        out = out / np.sqrt(np.real(np.sum(np.conj(out) * out)))
    return out


# --------------------------------------------------------------------------------------
# measurement                    (reference state.py:262-269, 307-387)
# --------------------------------------------------------------------------------------
def measurement_probabilities(psi_t, qubit_indices):
    """marginals over the listed qubits; first listed qubit = MSB of the outcome (state.py:263-267)."""
    qi = list(qubit_indices)
    rest = list(set(range(psi_t.ndim)) - set(qi))                  # ascending in CPython
    perm = qi + rest
    amps = np.transpose(psi_t, perm)
    ps = np.reshape(np.real(amps * np.conj(amps)), (2 ** len(qi), -1)).sum(axis=1)
    return ps, amps, perm


def sample_outcome(ps, u):
    """What ``np.random.choice(len(ps), p=ps)`` (state.py:364) does with its one uniform ``u``:
    legacy RandomState.choice = searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')."""
    cdf = np.cumsum(np.asarray(ps, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side='right'))


def outcome_bits(outcome, m):
    return tuple((outcome >> i) & 1 for i in range(m - 1, -1, -1))  # state.py:365


def measure(psi_t, qubit_indices, u, remove=False):
    """Returns (bits, new_tensor).  ``u`` is the uniform draw the reference would have
    taken from NumPy's global stream.  Follows state.py:333-387."""
    if abs(np.sum(probabilities(psi_t)) - 1.0) > 1e-4:
        raise RuntimeError('Vector is not a unit vector.')
    qi = list(qubit_indices)
    if qi == []:
        raise ValueError('Must measure at least one qubit.')
    if min(qi) < 0 or max(qi) >= psi_t.ndim:
        raise ValueError('qubit index out of range')
    if len(qi) != len(set(qi)):
        raise ValueError('repeated qubit index')
    ps, amps, perm = measurement_probabilities(psi_t, qi)
    outcome = sample_outcome(ps, u)
    bits = outcome_bits(outcome, len(qi))
    collapsed = amps[bits] / np.sqrt(ps[outcome])                  # :368-369
    if remove:
        new_t = np.array(collapsed, dtype=C128)                    # :374
    else:
        full = np.zeros_like(amps)
        full[bits] = collapsed
        new_t = np.array(np.transpose(full, np.argsort(perm)), dtype=C128)   # :376-378
    if new_t.ndim > 0:
        new_t = renormalize(new_t)                                 # :385
    return bits, new_t


# --------------------------------------------------------------------------------------
# density operators              (reference density_operator.py, operators.py:317-321)
# --------------------------------------------------------------------------------------
def outer_product(psi_t):
    """|psi><psi| in interleaved layout (density_operator.py:121-125)."""
    res = np.tensordot(np.conj(psi_t), psi_t, axes=0)
    d = res.ndim
    order = [v for pair in zip(range(d // 2, d), range(0, d // 2)) for v in pair]
    return np.array(res.transpose(order), dtype=C128)


def from_ensemble(states, ps):
    t = None
    for p, s in zip(ps, states):                                   # density_operator.py:106-115
        o = p * outer_product(s)
        t = o if t is None else t + o
    return t


def _apply_operator_to_operator(op_t, arg_t, qubit_indices):
    """Operator applied to the even (out) axes of an operator-shaped tensor
    (operators.py:226-228, 259-274)."""
    d = arg_t.ndim // 2
    k = op_t.ndim // 2
    qi = check_qubit_indices(k, d, qubit_indices)
    rest = sorted(set(range(d)) - set(qi))
    perm = qi + rest
    paired = [a for q in perm for a in (2 * q, 2 * q + 1)]
    arg = np.transpose(arg_t, paired)
    res = np.tensordot(op_t, arg, (list(range(1, 2 * k, 2)), list(range(0, 2 * k, 2))))
    # tensordot leaves [out_0..out_{k-1}, in_0..in_{k-1}, rest...]; re-interleave (:266-271)
    order = np.empty(2 * d, dtype=int)
    order[0:2 * k:2] = np.arange(k)
    order[1:2 * k:2] = np.arange(k, 2 * k)
    order[2 * k:] = np.arange(2 * k, 2 * d)
    res = np.array(np.transpose(res, order), dtype=C128)
    inv = np.argsort(perm)
    paired_inv = [a for q in inv for a in (2 * q, 2 * q + 1)]
    return np.transpose(res, paired_inv)


def compose_operators(op_t, arg_t, qubit_indices=None):
    """Operator(Operator) composition A(B) (operators.py:225-278 with an OperatorBase arg)."""
    return np.array(_apply_operator_to_operator(op_t, arg_t, qubit_indices), dtype=C128)


def apply_operator_density(op_t, rho_t, qubit_indices=None):
    """U rho U^dagger via adj -> apply -> adj -> apply (operators.py:317-321); no renormalise."""
    step = _apply_operator_to_operator(op_t, op_adjoint(rho_t), qubit_indices)
    step = op_adjoint(np.array(step, dtype=C128))
    return np.array(_apply_operator_to_operator(op_t, step, qubit_indices), dtype=C128)


def density_probabilities(rho_t):
    n = rho_t.ndim // 2
    return np.real(np.diag(op_to_matrix(rho_t))).reshape([2] * n)  # density_operator.py:60


def reduced_tensor(rho_t, retain):
    """partial trace keeping ``retain`` (density_operator.py:127-138)."""
    n = rho_t.ndim // 2
    retain = list(retain)
    traced = list(set(range(n)) - set(retain))
    perm = traced + retain
    paired = [a for q in perm for a in (2 * q, 2 * q + 1)]
    t = np.transpose(rho_t, paired)
    d1, d2 = len(traced), len(retain)
    D1 = 2 ** d1
    order = list(range(0, 2 * d1, 2)) + list(range(1, 2 * d1, 2)) + list(range(2 * d1, 2 * n))
    t = np.transpose(t, order).reshape([D1, D1] + [2] * (2 * d2))
    return t[range(D1), range(D1), ...].sum(axis=0)


def density_measurement_probabilities(rho_t, qubit_indices):
    t = reduced_tensor(rho_t, list(qubit_indices))
    return np.real(np.diag(op_to_matrix(t)))                       # density_operator.py:145-146


def density_measure(rho_t, qubit_indices, u, remove=False):
    """density_operator.py:202-286; returns (bits, new_tensor)."""
    n = rho_t.ndim // 2
    qi = list(qubit_indices)
    rest = list(set(range(n)) - set(qi))
    ps = density_measurement_probabilities(rho_t, qi)
    outcome = sample_outcome(ps, u)
    bits = outcome_bits(outcome, len(qi))
    perm = qi + rest
    paired = [a for q in perm for a in (2 * q, 2 * q + 1)]
    t = np.transpose(rho_t, paired)
    new_t = t[tuple(np.repeat(bits, 2))] / ps[outcome]             # :268-270
    if not remove:
        basis = np.zeros([2] * len(qi), dtype=C128)
        basis[bits] = 1
        new_t = np.tensordot(outer_product(basis), new_t, axes=0)  # :275-278
        inv = np.argsort(perm)
        paired_inv = [a for q in inv for a in (2 * q, 2 * q + 1)]
        new_t = np.transpose(new_t, paired_inv)                    # :279
    return bits, np.array(new_t, dtype=C128)


# --------------------------------------------------------------------------------------
# synthetic workloads            
# --------------------------------------------------------------------------------------
def layered_circuit(n, depth, seed=1234):
    """Seeded random layered circuit: per layer one of H / RX(theta) / RZ(theta) on every
    qubit, then a brick of CNOT (even layer pairs) or CZ (odd layer pairs) on neighbours.
    Returns a list of (name, theta, qubits)."""
    rng = np.random.default_rng(seed)
    gates = []
    for layer in range(depth):
        for q in range(n):
            g = int(rng.integers(3))
            theta = float(rng.uniform(0, 2 * np.pi))
            gates.append((('H', 'RX', 'RZ')[g], theta, (q,)))
        two = 'CNOT' if (layer // 2) % 2 == 0 else 'CZ'
        for q in range(layer % 2, n - 1, 2):
            gates.append((two, 0.0, (q, q + 1)))
    return gates


def gate_tensor(name, theta=0.0):
    if name == 'H':
        return hadamard()
    if name == 'RX':
        return rotation_x(theta)
    if name == 'RZ':
        return rotation_z(theta)
    if name == 'CNOT':
        return cnot()
    if name == 'CZ':
        return cz()
    raise KeyError(name)


def zeros_state(n):
    t = np.zeros([2] * n, dtype=C128)
    t.flat[0] = 1
    return t


def run_circuit(psi_t, gates, flat=False):
    """Run a gate list through the reference route (or the flat route)."""
    n = psi_t.ndim
    if flat:
        v = psi_t.reshape(-1)
        for name, theta, qs in gates:
            v = apply_matrix_flat(op_to_matrix(gate_tensor(name, theta)), v, [n - 1 - q for q in qs])
        return v.reshape([2] * n)
    for name, theta, qs in gates:
        psi_t = apply_operator(gate_tensor(name, theta), psi_t, list(qs))
    return psi_t
