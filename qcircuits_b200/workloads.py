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
