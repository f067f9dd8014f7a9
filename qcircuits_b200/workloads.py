"""Synthetic workloads of the benchmark configurations, as gate lists and Circuits."""
import numpy as np


def layered_gate_list(n, depth, seed=1234):
    """Seeded random layered circuit: per layer one of H / RX(theta) / RZ(theta) on every qubit
    (theta is always drawn), then a brick of CNOT (layer pairs 0, 2, ...) or CZ (pairs 1, 3, ...) on
    neighbouring qubits, offset alternating with the layer.  Returns [(name, theta, qubits)]."""
    rng = np.random.default_rng(seed)
    gates = []
    for layer in range(depth):
        for q in range(n):
            kind = int(rng.integers(3))
            theta = float(rng.uniform(0, 2 * np.pi))
            gates.append((('H', 'RX', 'RZ')[kind], theta, (q,)))
        name = 'CNOT' if (layer // 2) % 2 == 0 else 'CZ'
        gates.extend((name, 0.0, (q, q + 1)) for q in range(layer % 2, n - 1, 2))
    return gates


def operator_for(name, theta=0.0):
    from . import operators as ops
    if name == 'H':
        return ops.Hadamard()
    if name == 'RX':
        return ops.RotationX(theta)
    if name == 'RZ':
        return ops.RotationZ(theta)
    if name == 'CNOT':
        return ops.CNOT()
    if name == 'CZ':
        return ops.ControlledU(ops.PauliZ())
    raise KeyError(name)


def layered_circuit(n, depth, seed=1234):
    from .circuit import Circuit
    c = Circuit(n)
    cache = {}
    for name, theta, qs in layered_gate_list(n, depth, seed):
        key = (name, theta)
        op = cache.get(key)
        if op is None:
            op = cache[key] = operator_for(name, theta)
        c.append(op, list(qs))
    return c


def phase_estimation_gate_list(t, seed=1234):
    """BASELINE config 3, gate by gate: a t-qubit register estimates an eigenphase of a
    seeded random two-qubit unitary U held on qubits t, t+1.  H on the register, ControlledU(U^(2^(t-1-q)))
    on [q, t, t+1], then the inverse QFT as Swap / controlled-R_k^dagger / H gates.
    Returns ([(matrix or name, qubits)], eigenvector of U (4 amplitudes), its eigenphase in [0, 1))."""
    from scipy.stats import unitary_group
    U = unitary_group.rvs(4, random_state=seed)
    w, V = np.linalg.eig(U)
    vec, lam = V[:, 0] / np.linalg.norm(V[:, 0]), w[0] / abs(w[0])
    phase = (np.angle(lam) / (2 * np.pi)) % 1.0
    gates = [('H', (q,)) for q in range(t)]
    Vinv = np.linalg.inv(V)
    for q in range(t):
        power = 2 ** (t - 1 - q)
        Up = (V * (w / np.abs(w)) ** power) @ Vinv          # U^(2^(t-1-q)) through the eigendecomposition
        gates.append((('CU', Up), (q, t, t + 1)))
    for i in range(t // 2):
        gates.append(('SWAP', (i, t - 1 - i)))
    for q in range(t - 1, -1, -1):
        for k in range(t - q, 1, -1):
            gates.append((('CR', -2.0 * np.pi / 2 ** k), (q + k - 1, q)))
        gates.append(('H', (q,)))
    return gates, vec, phase


def phase_estimation_circuit(t, seed=1234):
    """(Circuit on t + 2 qubits, eigenvector, eigenphase) for phase_estimation_gate_list."""
    from . import operators as ops
    from .circuit import Circuit
    gates, vec, phase = phase_estimation_gate_list(t, seed)
    c = Circuit(t + 2)
    H, SW = ops.Hadamard(), ops.Swap()
    for g, qs in gates:
        if g == 'H':
            c.append(H, list(qs))
        elif g == 'SWAP':
            c.append(SW, list(qs))
        elif g[0] == 'CU':
            c.append(ops.ControlledU(ops.Operator.from_matrix(g[1])), list(qs))
        else:
            c.append(ops.ControlledU(ops.Operator.from_matrix(np.diag([1.0, np.exp(1j * g[1])]))), list(qs))
    return c, vec, phase
