"""CPU tests of the N>1 path: the distributed schedule (local steps / exchanges / rank specialisation)
and the all-to-all exchange itself over gloo with world_size 2.  The kernels are replaced by a NumPy
engine built on the oracle -- test infrastructure, the product engine is CUDA-only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qcircuits_b200 as qc
from oracle import qc_oracle as orc
from qcircuits_b200 import dist as qd
from qcircuits_b200.circuit import Circuit
from qcircuits_b200.workloads import layered_circuit
from tests.util import rand_unitary


def plan_matrix(plan):
    """(full matrix, operator-qubit order) of a GatePlan: controls first, then targets"""
    t = len(plan.targets)
    red = np.diag(plan.matrix) if plan.diagonal else np.asarray(plan.matrix)
    k = len(plan.controls) + t
    full = np.eye(2 ** k, dtype=complex)
    full[-2 ** t:, -2 ** t:] = red
    return full, list(plan.controls) + list(plan.targets)


class NumpyEngine:
    def __init__(self, n_local, group=None):
        self.n_local = n_local
        self.vec = np.zeros(2 ** n_local, dtype=np.complex128)
        self.group = group
        self.launches = 0

    def set_basis(self, index):
        self.vec[:] = 0
        if index is not None:
            self.vec[index] = 1

    def run_local(self, rank_gates, max_ops):
        for g in rank_gates:
            M, order = plan_matrix(g.plan)
            self.vec = orc.apply_matrix_flat(M, self.vec, [g.bits[q] for q in order], renorm=False)

    def permute(self, src_bit):
        n = self.n_local
        idx = np.arange(2 ** n)
        j = np.zeros_like(idx)
        for b in range(n):
            j |= ((idx >> b) & 1) << src_bit[b]
        self.vec = self.vec[j]

    def exchange(self, group=None):
        src = torch.from_numpy(self.vec.view(np.float64).copy())
        out = torch.empty_like(src)
        dist.all_to_all_single(out, src, group=group)
        self.vec = out.numpy().view(np.complex128).copy()

    def norm2(self):
        return torch.tensor([float(np.sum(np.abs(self.vec) ** 2))], dtype=torch.float64)

    # the drop-in State surface (same contract as dist.CudaEngine)
    def load_host(self, vec):
        self.vec = np.array(vec, dtype=np.complex128).reshape(-1)

    def to_host(self):
        return self.vec.copy()

    def zero(self):
        self.vec[:] = 0

    def scale(self, factor):
        self.vec = self.vec * factor

    def collapse(self, bitpos, outcome, remove, scale):
        n, m = self.n_local, len(bitpos)
        idx = np.arange(2 ** n)
        x = np.zeros_like(idx)
        for p in bitpos:
            x = (x << 1) | ((idx >> p) & 1)
        keep = x == outcome
        if remove:
            rest = [b for b in range(n) if b not in set(bitpos)]
            sel = idx[keep]
            dst = np.zeros_like(sel)
            for j, b in enumerate(rest):
                dst |= ((sel >> b) & 1) << j
            out = np.zeros(2 ** (n - m), dtype=np.complex128)
            out[dst] = self.vec[sel] * scale
            self.vec, self.n_local = out, n - m
        else:
            self.vec = np.where(keep, self.vec * scale, 0)
        self.launches += 1

    def kron_low(self, vec):
        vec = np.asarray(vec).reshape(-1)
        self.vec = np.kron(self.vec, vec)
        self.n_local += int(round(np.log2(vec.size)))

    def clone(self):
        other = NumpyEngine(self.n_local, self.group)
        other.vec = self.vec.copy()
        return other

    def marginals(self, bitpos):
        n = self.n_local
        idx = np.arange(2 ** n)
        x = np.zeros_like(idx)
        for p in bitpos:
            x = (x << 1) | ((idx >> p) & 1)
        return torch.from_numpy(np.bincount(x, weights=np.abs(self.vec) ** 2, minlength=2 ** len(bitpos)))


def build_circuit(n, seed=3):
    rng = np.random.default_rng(seed)
    c = layered_circuit(n, 5, seed=1234)
    c.append(qc.Toffoli(), [0, n - 1, 3])
    c.append(qc.ControlledU(np.exp(0.4j) * qc.RotationZ(0.8)), [1, 0])
    c.append(qc.Swap(), [0, n - 1])                      # dense on both sides of the exchange
    c.append(qc.Operator.from_matrix(rand_unitary(rng, 2)), [2, 1])
    c.append(qc.RotationZ(0.3), [0])
    c.append(qc.CNOT(), [n - 1, 0])
    for q in range(n):
        c.append(qc.Hadamard(), [q])
    return c


def oracle_state(c):
    vec = np.zeros(2 ** c.n, dtype=complex)
    vec[0] = 1
    for g in c.gates:
        vec = orc.apply_matrix_flat(g.plan.full, vec, g.bits, renorm=False)
    return vec


def assemble(shards, phys, n, g):
    """shards[r] (physical layout) -> canonical column vector (qubit q on flat bit n-1-q)"""
    n_local = n - g
    full = np.concatenate(shards)
    idx = np.arange(2 ** n)
    src = np.zeros_like(idx)
    for q in range(n):
        src |= ((idx >> (n - 1 - q)) & 1) << phys[q]
    assert n_local + g == n
    return full[src]


@pytest.mark.parametrize('n,g', [(8, 1), (9, 2), (10, 3)])
def test_schedule_and_specialisation_single_process(n, g):
    """All P ranks simulated in one process; the exchange is the block transpose it stands for."""
    c = build_circuit(n)
    want = oracle_state(c)
    P, n_local = 2 ** g, n - g
    gates = [qd._LogicalGate(gt.plan, [n - 1 - b for b in gt.bits]) for gt in c.gates]
    steps, phys = qd.dist_schedule(gates, n, g)
    assert sum(1 for s in steps if s[0] == 'swap') >= 1
    engines = [NumpyEngine(n_local) for _ in range(P)]
    for r, e in enumerate(engines):
        e.set_basis(0 if r == 0 else None)
    from qcircuits_b200.circuit import _Gate
    for step in steps:
        if step[0] == 'local':
            for r, e in enumerate(engines):
                rg = []
                for idx, bits in step[1]:
                    sp = qd.specialize(gates[idx].plan, bits, n_local, r)
                    if sp is not None:
                        assert all(b < n_local for b in sp[1])
                        rg.append(_Gate(None, sp[1], plan=sp[0]))
                e.run_local(rg, 24)
        elif step[0] == 'lperm':
            for e in engines:
                e.permute(step[1])
        else:
            blocks = [e.vec.reshape(P, -1) for e in engines]
            for r, e in enumerate(engines):
                e.vec = np.concatenate([blocks[src][r] for src in range(P)])
    got = assemble([e.vec for e in engines], phys, n, g)
    assert np.max(np.abs(got - want)) < 1e-12


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        c = build_circuit(n)
        g = int(np.log2(world))
        st = qd.ShardedState(n, np.complex128, engine=NumpyEngine(n - g))
        st.set_zeros().run(c)
        total = st.norm2()
        marg = st.marginals([0, n - 1, 2])
        shards = [None] * world
        dist.all_gather_object(shards, st.engine.vec)
        if rank == 0:
            out.put((shards, st.phys, total, marg, st.exchanges))
    finally:
        dist.destroy_process_group()


def test_sharded_state_over_gloo_world_size_2():
    n, world = 8, 2
    c = build_circuit(n)
    want = oracle_state(c)
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    shards, phys, total, marg, exchanges = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = assemble(shards, phys, n, 1)
    assert np.max(np.abs(got - want)) < 1e-12
    assert abs(total - 1.0) < 1e-12 and exchanges >= 1
    ps, _, _ = orc.measurement_probabilities(want.reshape([2] * n), [0, n - 1, 2])
    assert np.max(np.abs(marg - ps)) < 1e-12


def test_headline_sharded_schedule_is_mostly_local():
    """36 qubits over 8 ranks: count the exchanges the schedule needs for the depth-40 circuit."""
    n, g = 36, 3
    c = layered_circuit(n, 40, seed=1234)
    gates = [qd._LogicalGate(gt.plan, [n - 1 - b for b in gt.bits]) for gt in c.gates]
    steps, phys = qd.dist_schedule(gates, n, g)
    swaps = sum(1 for s in steps if s[0] == 'swap')
    local = sum(len(s[1]) for s in steps if s[0] == 'local')
    assert local == len(c.gates)
    assert sorted(phys) == list(range(n))
    assert swaps <= 40
    print('exchanges for 36q depth 40 over 8 ranks:', swaps)


def _dropin_worker(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = int(np.log2(world))
        rng = np.random.default_rng(77)
        vec = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        vec /= np.linalg.norm(vec)
        results = {}
        st = qd.ShardedState.from_column_vector(vec, np.complex128, engine=NumpyEngine(n - g))
        assert st.rank == n and st.shape == (2,) * n
        results['roundtrip'] = st.to_column_vector()
        # Operator.__call__ leaves its argument alone and returns a new sharded state
        U = rand_unitary(np.random.default_rng(5), 2)
        st2 = qc.Operator.from_matrix(U)(st, qubit_indices=[0, n - 2])
        results['applied'] = st2.to_column_vector()
        results['untouched'] = st.to_column_vector()
        results['marg'] = st2.marginal_probabilities([n - 1, 0, 3])
        # measurement: same uniform on every rank; rank-bit and local qubits, keep and remove
        cases = []
        for qs, remove, u in (([0, 3], False, 0.37), ([n - 1], False, 0.81), ([0], True, 0.05), ([2, 0, n - 1], True, 0.66),
                              (None, False, 0.5)):
            s3 = st2.copy()
            bits = s3.measure(qs, remove=remove, u=u)
            cases.append((qs, remove, u, bits, s3.to_column_vector(), s3.rank))
        results['measure'] = cases
        # int argument -> int result; the global NumPy stream on rank 0 supplies the uniform when u is None
        np.random.seed(123)
        s4 = st2.copy()
        b = s4.measure(1)
        results['int_result'] = (b, s4.to_column_vector())
        # tensor products with an ordinary (small) state on either side
        phi = rng.normal(size=4) + 1j * rng.normal(size=4)
        phi /= np.linalg.norm(phi)

        results['phi'] = phi
        k1 = st2.copy()
        k1.engine.kron_low(phi)
        k1.phys = [p + 2 for p in st2.phys] + [1, 0]
        k1.n, k1.n_local = n + 2, st2.n_local + 2
        results['kron_right'] = k1.to_column_vector()
        k2 = st2.copy()
        k2.engine.kron_low(phi)
        k2.phys = [1, 0] + [p + 2 for p in st2.phys]
        k2.n, k2.n_local = n + 2, st2.n_local + 2
        results['kron_left'] = k2.to_column_vector()
        if rank == 0:
            out.put((vec, U, results))
    finally:
        dist.destroy_process_group()



@pytest.mark.parametrize('world', [2, 4])
def test_sharded_dropin_state_over_gloo(world):
    """from_column_vector / Operator.__call__ / marginals / measure (keep + remove, rank-bit + local qubits) /
    tensor product of a ShardedState against the oracle, world_size 2 and 4 on gloo."""
    n = 9
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dropin_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    vec, U, res = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.max(np.abs(res['roundtrip'] - vec)) < 1e-13
    assert np.max(np.abs(res['untouched'] - vec)) < 1e-13
    want = orc.apply_matrix_flat(U, vec, [n - 1 - 0, n - 1 - (n - 2)], renorm=True)
    assert np.max(np.abs(res['applied'] - want)) < 1e-12
    ps, _, _ = orc.measurement_probabilities(want.reshape([2] * n), [n - 1, 0, 3])
    assert np.max(np.abs(res['marg'] - ps)) < 1e-12
    for qs, remove, u, bits, post, rank_after in res['measure']:
        qi = list(range(n)) if qs is None else qs
        bits_want, post_want = orc.measure(want.reshape([2] * n), qi, u, remove=remove)
        assert tuple(bits) == tuple(bits_want), (qs, remove)
        assert rank_after == (n - len(qi) if remove else n)
        assert np.max(np.abs(post - post_want.reshape(-1))) < 1e-12, (qs, remove)
    b, post = res['int_result']
    np.random.seed(123)
    bits_want, post_want = orc.measure(want.reshape([2] * n), [1], float(np.random.random_sample()), remove=False)
    assert b == bits_want[0] and isinstance(b, (int, np.integer))
    assert np.max(np.abs(post - post_want.reshape(-1))) < 1e-12
    assert np.max(np.abs(res['kron_right'] - np.kron(want, res['phi']))) < 1e-12
    assert np.max(np.abs(res['kron_left'] - np.kron(res['phi'], want))) < 1e-12


def test_norm_bound_is_kept_by_unitary_circuits_only():
    """dist._all_unitary decides whether a sharded state keeps the device scalar that ranges the fp16 tensor-core
    operands (qcs_apply_fused_bounded): every gate unitary -> kept; a projector or a scaled gate -> measured again"""
    import qcircuits_b200 as qc
    from qcircuits_b200.circuit import Circuit
    from qcircuits_b200.dist import _all_unitary
    c = Circuit(6)
    c.append(qc.Hadamard(), [0])
    c.append(qc.CNOT(), [0, 3])
    c.append(qc.RotationZ(0.3), [5])
    assert _all_unitary(c)
    assert _all_unitary(c)                                   # (cached per gate count)
    c.append(qc.Operator.from_matrix(np.array([[1, 0], [0, 0]], dtype=complex)), [2])
    assert not _all_unitary(c)
    c2 = Circuit(4)
    c2.append(qc.Operator.from_matrix(2.0 * np.eye(2, dtype=complex)), [1])
    assert not _all_unitary(c2)
