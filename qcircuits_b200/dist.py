"""Sharded states: one process per GPU, the state split on its top log2(P) flat bits (north_star (4),
SURVEY.md 8e).

A state of n qubits over P = 2^g ranks: rank r owns the contiguous block of 2^(n-g) amplitudes whose
top g physical bits equal r.  Which *logical* qubit sits on which *physical* bit is tracked by
``ShardedState.phys`` (qubit -> physical bit), so that

* a gate whose non-diagonal targets are all on local physical bits runs with no communication, through
  the same fused tile passes as on one GPU.  Control bits and diagonal targets on rank bits become
  per-rank constants (``specialize``): the op is dropped, loses the control, or turns into a phase;
* when the next gates need a qubit that sits on a rank bit, the g rank bits are exchanged with the top g
  local bits by ONE NCCL all-to-all over NVLink (``torch.distributed.all_to_all_single`` on contiguous
  2^(n-2g)-amplitude blocks: no pack/unpack kernel), and only the map changes;
* the norm is reduced across ranks once, at the end (linear ops commute with the scalar).

``dist_schedule`` is pure host logic and is the same on every rank; it is exercised on CPU (gloo,
world_size 2) in tests/test_dist_cpu.py with a NumPy engine standing in for the kernels.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._gates import GatePlan
from .circuit import _Gate, schedule_passes


# ---------------------------------------------------------------------------------------------------
# host logic
# ---------------------------------------------------------------------------------------------------
def initial_layout(n, g):
    """qubit -> physical bit.  Qubits 0..g-1 start on the rank bits (canonical: bit n-1-q).  The top g
    LOCAL bits -- the ones that are traded away by the first exchange -- get the qubits from the far end
    of the register, so that the two sets that alternate on the rank bits block as little of the circuit
    as possible; everything else keeps its order."""
    n_local = n - g
    phys = [None] * n
    for q in range(g):
        phys[q] = n - 1 - q
    far = list(range(n - g, n)) if n - g >= g else []
    for i, q in enumerate(far):
        phys[q] = n_local - 1 - i
    free_bits = [b for b in range(n_local - 1, -1, -1) if b not in set(p for p in phys if p is not None)]
    rest = [q for q in range(n) if phys[q] is None]
    for q, b in zip(rest, free_bits):
        phys[q] = b
    return phys


def canonical_layout(n):
    return [n - 1 - q for q in range(n)]


def dist_schedule(gates, n, g, phys=None):
    """Split a gate list into steps that need no communication and exchanges.

    ``gates``: records with ``plan`` (GatePlan) and ``qubits`` (logical qubit of each operator qubit).
    Returns (steps, final_phys); a step is ('local', [(gate index, [physical bit per operator qubit])]),
    ('swap',) -- trade the g rank bits with the top g local bits -- or ('lperm', src_bit) -- a local bit
    permutation (dst bit b takes src bit src_bit[b]), only needed when one gate has non-diagonal targets
    in both halves of an exchange."""
    n_local = n - g
    phys = list(initial_layout(n, g) if phys is None else phys)
    steps = []
    pending = list(range(len(gates)))
    qmask = []
    for gate in gates:
        plan = gate.plan
        nd = 0 if plan.diagonal else sum(1 << gate.qubits[t] for t in plan.targets)
        d = sum(1 << gate.qubits[c] for c in plan.controls)
        if plan.diagonal:
            d |= sum(1 << gate.qubits[t] for t in plan.targets)
        qmask.append((nd, d))
    stuck = 0
    while pending:
        blk_nd = blk_d = 0
        now, rest = [], []
        for idx in pending:
            nd, d = qmask[idx]
            local = all(phys[q] < n_local for q in range(n) if (nd >> q) & 1)
            if local and (nd & (blk_nd | blk_d)) == 0 and (d & blk_nd) == 0:
                now.append((idx, [phys[q] for q in gates[idx].qubits]))
            else:
                rest.append(idx)
                blk_nd |= nd
                blk_d |= d
        if now:
            steps.append(('local', now))
            stuck = 0
        pending = rest
        if not pending:
            break
        if g == 0:
            raise RuntimeError('unschedulable gate on a single shard')
        stuck += 1
        if stuck >= 2:
            # the first pending gate has non-diagonal targets on both sides of the exchange: move its
            # local targets out of the top g local bits first
            nd, _ = qmask[pending[0]]
            need_local = [q for q in range(n) if (nd >> q) & 1 and phys[q] < n_local]
            top = set(range(n_local - g, n_local))
            movers = [q for q in need_local if phys[q] in top]
            safe = [b for b in range(n_local - g) if all(phys[q] != b or not ((nd >> q) & 1) for q in range(n))]
            if len(movers) > len(safe):
                raise RuntimeError('gate too wide for this sharding')
            src_bit = list(range(n_local))
            qubit_at = {phys[q]: q for q in range(n)}
            for q, b in zip(movers, safe):
                a = phys[q]
                other = qubit_at[b]
                src_bit[a], src_bit[b] = src_bit[b], src_bit[a]
                phys[q], phys[other] = b, a
                qubit_at[b], qubit_at[a] = q, other
            if movers:
                steps.append(('lperm', src_bit))
            elif stuck >= 4:
                raise RuntimeError('scheduler made no progress')
        steps.append(('swap',))
        qubit_at = {phys[q]: q for q in range(n)}
        for i in range(g):
            qa, qb = qubit_at[n_local + i], qubit_at[n_local - g + i]
            phys[qa], phys[qb] = n_local - g + i, n_local + i
    return steps, phys


def specialize(plan, bits, n_local, rank):
    """A gate as seen by one rank: controls / diagonal targets on rank bits are resolved with the rank's
    bit values.  Returns None (identity on this rank) or (GatePlan, local bits per operator qubit)."""
    def rank_bit(p):
        return (rank >> (p - n_local)) & 1

    controls = []
    for c in plan.controls:
        if bits[c] >= n_local:
            if not rank_bit(bits[c]):
                return None
        else:
            controls.append(c)
    targets = list(plan.targets)
    matrix = plan.matrix
    if plan.diagonal:
        diag = np.asarray(matrix).reshape([2] * len(targets))
        keep = []
        index = []
        for t in targets:
            if bits[t] >= n_local:
                index.append(rank_bit(bits[t]))
            else:
                index.append(slice(None))
                keep.append(t)
        diag = np.ascontiguousarray(diag[tuple(index)]).reshape(-1)
        targets = keep
        if not targets:
            # a pure phase on this rank (conditioned on the remaining local controls, if any)
            phase = complex(diag[0])
            if controls:
                t = controls.pop()
                targets, diag = [t], np.array([1.0, phase], dtype=np.complex128)
            else:
                if phase == 1.0:
                    return None
                any_bit = 0
                new_plan = GatePlan(1, [], [0], True, np.array([phase, phase], dtype=np.complex128), None)
                return new_plan, [any_bit]
        matrix = diag
    else:
        if any(bits[t] >= n_local for t in targets):
            raise ValueError('non-diagonal target on a rank bit')
    # renumber operator qubits compactly
    used = controls + targets
    remap = {q: i for i, q in enumerate(used)}
    new_plan = GatePlan(len(used), [remap[c] for c in controls], [remap[t] for t in targets], plan.diagonal,
                        np.ascontiguousarray(matrix), None)
    return new_plan, [bits[q] for q in used]


def _all_unitary(circuit):
    """every gate of the circuit is a unitary (cached on the circuit per gate count)"""
    key = len(circuit.gates)
    hit = circuit.__dict__.get('_unitary_cache')
    if hit is None or hit[0] != key:
        ok = True
        for gt in circuit.gates:
            M = gt.plan.full
            if M is None or np.abs(M.conj().T @ M - np.eye(M.shape[0])).max() > 1e-6:
                ok = False
                break
        hit = circuit.__dict__['_unitary_cache'] = (key, ok)
    return hit[1]


class _LogicalGate:
    __slots__ = ('plan', 'qubits')

    def __init__(self, plan, qubits):
        self.plan, self.qubits = plan, list(qubits)


# ---------------------------------------------------------------------------------------------------
# engines: what a rank does to its shard (CUDA in the product; tests plug in a NumPy checker)
# ---------------------------------------------------------------------------------------------------
class _PeerShard:
    """A shard allocated with qcs_dist_alloc (a plain cudaMalloc block, so that its CUDA IPC handle names exactly this
    buffer) and exposed to torch through __cuda_array_interface__; freed when the last tensor over it goes away."""

    def __init__(self, nreals, real_typestr, itemsize):
        ptr = ctypes.c_void_p()
        _lib.check(_lib.load().qcs_dist_alloc(ctypes.byref(ptr), nreals * itemsize))
        self.ptr = ptr.value
        self.__cuda_array_interface__ = {'shape': (nreals,), 'typestr': real_typestr, 'data': (self.ptr, False), 'version': 2}

    def __del__(self):
        try:
            _lib.load().qcs_dist_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def _exchange_mode():
    """'p2p' (default): the exchange is libqcs's own kernel pulling from the peers' HBM over NVLink (qcs_dist_swap_rank);
    'nccl': torch.distributed.all_to_all_single."""
    return os.environ.get('QCS_DIST_EXCHANGE', 'p2p')


class CudaEngine:
    """Shard + spare buffer on the current CUDA device; passes run through libqcs."""

    P2P_MIN_LOCAL_BITS = 13

    def __init__(self, n_local, dtype):
        self.n_local, self.dtype = n_local, np.dtype(dtype)
        self.code = _lib.dtype_code(dtype)
        self._p2p = _exchange_mode() == 'p2p'
        self._peer_map = {}        # local data_ptr -> [pointer to that buffer's counterpart on every rank]
        self._opened = []          # mappings to close before the next registration
        self._retired = []         # old shards that peers may still have mapped
        self.buf = self._alloc_shard(n_local)
        self.spare = self._alloc_shard(n_local)
        self.launches = 0
        self.exchange_kind = None
        # device scalar >= the squared norm of the WHOLE state (None: not known): ranges the fp16 operands of the
        # tensor-core sweeps (qcs_apply_fused_bounded) so that no pass has to measure its shard
        self.bound2 = None

    def _alloc_shard(self, n_local):
        if not self._p2p:
            return _lib.alloc_amplitudes(n_local, self.dtype)
        t = _lib.require_cuda()
        c64 = self.dtype == np.dtype(np.complex64)
        holder = _PeerShard(2 << n_local, '<f4' if c64 else '<f8', 4 if c64 else 8)
        return t.as_tensor(holder, device='cuda')

    def _retire(self, tensor):
        """keep a replaced shard alive until the peers have dropped their mappings (next registration)"""
        if self._p2p and tensor.data_ptr() in self._peer_map:
            self._retired.append(tensor)

    def set_basis(self, index):
        lib = _lib.load()
        if index is None:
            self.buf.zero_()
        else:
            _lib.check(lib.qcs_set_basis(self.buf.data_ptr(), self.n_local, self.code, index, _lib.stream_ptr()))
        self.bound2 = _lib.torch().ones(2, dtype=_lib.torch().float64, device='cuda')     # a basis state: norm 1

    def compile_local(self, rank_gates, max_ops):
        """schedule + plan the fused passes of a communication-free step once (ShardedState.run caches the result)"""
        from .circuit import build_passes
        groups = schedule_passes(rank_gates, self.n_local, self.dtype, max_ops=max_ops)
        return build_passes(rank_gates, groups, self.n_local, self.dtype)

    def run_compiled(self, passes):
        lib = _lib.load()
        ws = _lib.workspace().data_ptr()
        bound = None if self.bound2 is None else self.bound2.data_ptr()
        for ops, keep, _ in passes:
            _lib.check(lib.qcs_apply_fused_bounded(self.buf.data_ptr(), self.buf.data_ptr(), self.n_local, self.code, ops,
                                                   len(ops), None, None, bound, ws, _lib.stream_ptr()))
            self.launches += 1

    def run_local(self, rank_gates, max_ops):
        self.run_compiled(self.compile_local(rank_gates, max_ops))

    def permute(self, src_bit):
        _lib.check(_lib.load().qcs_permute_bits(self.spare.data_ptr(), self.buf.data_ptr(), self.n_local, self.code,
                                               _lib.int32_array(src_bit), _lib.stream_ptr()))
        self.buf, self.spare = self.spare, self.buf
        self.launches += 1

    def exchange(self, group=None):
        """all rank bits <-> the top g local bits: block p of rank r goes to block r of rank p"""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if self._p2p and self.n_local >= self.P2P_MIN_LOCAL_BITS and world <= 16 and self._register_peers(group):
            self._exchange_p2p(group, world)
            self.exchange_kind = 'p2p'
        else:
            dist.all_to_all_single(self.spare, self.buf, group=group)
            self.exchange_kind = 'nccl'
        self.buf, self.spare = self.spare, self.buf

    def _register_peers(self, group):
        """Map the peers' shard and spare buffers into this process (CUDA IPC), once per pair of buffers.  Collective;
        returns False on every rank if any rank cannot map (the exchange then falls back to NCCL for good)."""
        import torch.distributed as dist
        if self.buf.data_ptr() in self._peer_map and self.spare.data_ptr() in self._peer_map:
            return True
        t = _lib.require_cuda()
        lib = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        for m in self._opened:
            lib.qcs_ipc_close(ctypes.c_void_p(m))
        self._opened, self._peer_map = [], {}
        handles = (ctypes.c_ubyte * 128)()
        ok = 1
        for k, tensor in enumerate((self.buf, self.spare)):
            if lib.qcs_ipc_get_handle(ctypes.c_void_p(tensor.data_ptr()), ctypes.byref(handles, 64 * k)) != 0:
                ok = 0
        mine = t.tensor(list(handles) + [ok], dtype=t.uint8, device='cuda')
        everyone = t.empty(world * 129, dtype=t.uint8, device='cuda')
        dist.all_gather_into_tensor(everyone, mine, group=group)      # (also: every rank has closed its old mappings)
        self._retired = []
        table = everyone.cpu().numpy().reshape(world, 129)
        ok = int(table[:, 128].min())
        maps = {self.buf.data_ptr(): [0] * world, self.spare.data_ptr(): [0] * world}
        if ok:
            for k, tensor in enumerate((self.buf, self.spare)):
                for r in range(world):
                    if r == rank:
                        maps[tensor.data_ptr()][r] = tensor.data_ptr()
                        continue
                    raw = (ctypes.c_ubyte * 64)(*[int(x) for x in table[r, 64 * k:64 * k + 64]])
                    mapped = ctypes.c_void_p()
                    if lib.qcs_ipc_open(raw, ctypes.byref(mapped)) != 0:
                        ok = 0
                        break
                    self._opened.append(mapped.value)
                    maps[tensor.data_ptr()][r] = mapped.value
        flag = t.tensor([ok], dtype=t.int32, device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            for m in self._opened:
                lib.qcs_ipc_close(ctypes.c_void_p(m))
            self._opened, self._p2p = [], False
            return False
        self._peer_map = maps
        return True

    def release_peers(self, group=None):
        """Collective: drop every mapping of the peers' shards (CUDA IPC) so that the owners can really free them.  Call
        it before discarding a large sharded state inside a long-lived process group; the next exchange maps again."""
        import torch.distributed as dist
        lib = _lib.load()
        _lib.require_cuda().cuda.current_stream().synchronize()
        for m in self._opened:
            lib.qcs_ipc_close(ctypes.c_void_p(m))
        self._opened, self._peer_map = [], {}
        dist.barrier(group)
        self._retired = []

    def _exchange_p2p(self, group, world):
        import torch.distributed as dist
        t = _lib.require_cuda()
        g = int(round(np.log2(world)))
        rank = dist.get_rank(group)
        if not hasattr(self, '_flag'):
            self._flag = t.zeros(1, dtype=t.int32, device='cuda')
        push = int(os.environ.get('QCS_DIST_PUSH', '1')) != 0      # remote stores (682 GB/s) beat remote loads (627 GB/s)
        local, table = (self.buf, self.spare) if push else (self.spare, self.buf)
        peers = (ctypes.c_void_p * world)(*self._peer_map[table.data_ptr()])
        dist.all_reduce(self._flag, group=group)       # device-side barrier: every rank's shard is complete, its spare free
        _lib.check(_lib.load().qcs_dist_swap_rank(local.data_ptr(), peers, world, rank, self.n_local, self.code,
                                                  _lib.int32_array(list(range(g))),
                                                  _lib.int32_array([self.n_local - g + j for j in range(g)]), g,
                                                  1 if push else 0, _lib.stream_ptr()))
        dist.all_reduce(self._flag, group=group)       # ... and nobody rewrites a shard its peers are still reading
        self.launches += 1

    def norm2(self):
        out = _lib.alloc_scalar()
        _lib.check(_lib.load().qcs_abs2(None, self.buf.data_ptr(), self.n_local, self.code, None, out.data_ptr(),
                                        _lib.workspace().data_ptr(), _lib.stream_ptr()))
        return out[:1]

    # -- the drop-in State surface (measure, host transfer, tensor product) ----------------------------------------
    def load_host(self, vec):
        """this rank's 2^n_local amplitudes from a host array"""
        self.buf.copy_(_lib.upload(np.asarray(vec).reshape(-1), self.dtype))
        self.bound2 = None

    def to_host(self):
        return _lib.download(self.buf, self.dtype).astype(np.complex128)

    def zero(self):
        self.buf.zero_()
        self.bound2 = None

    def scale(self, factor):
        _lib.check(_lib.load().qcs_scale(self.buf.data_ptr(), self.buf.data_ptr(), self.n_local, self.code, float(factor), 0.0,
                                         None, None, None, _lib.stream_ptr()))
        self.bound2 = None

    def collapse(self, bitpos, outcome, remove, scale):
        """keep the amplitudes whose bits at bitpos equal outcome, times scale (qcs_collapse); remove drops those bits"""
        lib = _lib.load()
        m = len(bitpos)
        n_out = self.n_local - m if remove else self.n_local
        dst = self._alloc_shard(n_out) if remove else self.spare
        _lib.check(lib.qcs_collapse(dst.data_ptr(), self.buf.data_ptr(), self.n_local, self.code, _lib.int32_array(bitpos), m,
                                    int(outcome), 1 if remove else 0, float(scale), None, _lib.workspace().data_ptr(),
                                    _lib.stream_ptr()))
        if remove:
            self._retire(self.buf)
            self._retire(self.spare)
            self.buf, self.n_local = dst, n_out
            self.spare = self._alloc_shard(n_out)
        else:
            self.buf, self.spare = self.spare, self.buf
        self.launches += 1
        self.bound2 = None

    def kron_low(self, vec):
        """shard <- shard (x) vec: the new qubits become the lowest physical bits (qcs_kron)"""
        vec = np.asarray(vec).reshape(-1)
        nb = int(round(np.log2(vec.size)))
        small = _lib.upload(vec, self.dtype)
        dst = self._alloc_shard(self.n_local + nb)
        _lib.check(_lib.load().qcs_kron(dst.data_ptr(), self.buf.data_ptr(), self.n_local, small.data_ptr(), nb, self.code,
                                        None, None, None, None, _lib.stream_ptr()))
        self._retire(self.buf)
        self._retire(self.spare)
        self.n_local += nb
        self.buf = dst
        self.spare = self._alloc_shard(self.n_local)
        self.bound2 = None

    def clone(self):
        other = CudaEngine(self.n_local, self.dtype)
        other.buf.copy_(self.buf)
        other.bound2 = None if self.bound2 is None else self.bound2.clone()
        return other

    def marginals(self, bitpos):
        t = _lib.torch()
        out = t.empty(1 << len(bitpos), dtype=t.float64, device='cuda')
        ws = _lib.workspace()
        _lib.check(_lib.load().qcs_marginal_probs(out.data_ptr(), self.buf.data_ptr(), self.n_local, self.code,
                                                  _lib.int32_array(bitpos), len(bitpos), None, ws.data_ptr(),
                                                  ws.numel(), _lib.stream_ptr()))
        return out


class ShardedState:
    """This rank's part of an n-qubit state spread over ``world`` = 2^g ranks (one process per GPU).

    The drop-in surface of a sharded state (north_star (4): "states of 32 qubits or more sharded ... across 2/4/8
    GPUs") mirrors qc.State where the operation makes sense at that size: ``Operator.__call__(state, qubit_indices)``
    (operators.py:280-323), ``Circuit.run``, ``measure`` (state.py:307-387: per-rank marginal histogram -> all-reduce ->
    the SAME host inverse-CDF draw on every rank -> local collapse; a measured rank bit zeroes whole shards),
    marginal probabilities of chosen qubits, ``from_column_vector`` / ``to_column_vector`` (state.py:32-60,99-111) and
    the tensor product with an ordinary State (tensors.py:56-76).  Every method is collective: all ranks call it with
    the same arguments."""

    def __init__(self, n, dtype=np.complex64, rank=None, world=None, engine=None, group=None):
        import torch.distributed as dist
        self.proc_rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.g = int(round(np.log2(self.world)))
        if 2 ** self.g != self.world:
            raise ValueError('the number of ranks must be a power of two')
        self.n, self.n_local, self.dtype, self.group = n, n - self.g, np.dtype(dtype), group
        if self.n_local < 2 * self.g:
            raise ValueError('state too small to shard over this many ranks')
        self.engine = CudaEngine(self.n_local, dtype) if engine is None else engine
        self.phys = canonical_layout(n)
        self.norm = None
        self.exchanges = 0

    def set_zeros(self, layout=None):
        """|0...0>: index 0 lives on rank 0 under any layout."""
        self.phys = list(initial_layout(self.n, self.g) if layout is None else layout)
        self.engine.set_basis(0 if self.proc_rank == 0 else None)
        self.norm = None
        return self

    def run(self, circuit, max_ops=128):
        """Apply a Circuit (its gates are on logical qubits).  Collective: every rank calls it."""
        # the plan -- exchange schedule, rank-specialised gates, fused passes -- depends only on the circuit and the
        # layout it starts from: built once per (circuit, layout), like Circuit._plans on one GPU
        key = (id(circuit), len(circuit.gates), tuple(self.phys), max_ops, self.n_local)
        cache = self.__dict__.setdefault('_plans', {})
        plan = cache.get(key)
        if plan is None:
            gates = [_LogicalGate(gt.plan, [circuit.n - 1 - b for b in gt.bits]) for gt in circuit.gates]
            steps, phys = dist_schedule(gates, self.n, self.g, self.phys)
            compiled = []
            for step in steps:
                if step[0] == 'local':
                    rank_gates = []
                    for idx, bits in step[1]:
                        sp = specialize(gates[idx].plan, bits, self.n_local, self.proc_rank)
                        if sp is not None:
                            rank_gates.append(_Gate(None, sp[1], plan=sp[0]))
                    if rank_gates:
                        if hasattr(self.engine, 'compile_local'):
                            compiled.append(('passes', self.engine.compile_local(rank_gates, max_ops)))
                        else:
                            compiled.append(('gates', rank_gates))
                else:
                    compiled.append(step)
            plan = cache[key] = (compiled, phys, circuit)          # (the circuit is kept alive: its id is the key)
        compiled, phys, _ = plan
        self._ensure_bound()
        for step in compiled:
            if step[0] == 'passes':
                self.engine.run_compiled(step[1])
            elif step[0] == 'gates':
                self.engine.run_local(step[1], max_ops)
            elif step[0] == 'lperm':
                self.engine.permute(step[1])
            else:
                self.engine.exchange(self.group)
                self.exchanges += 1
        self.phys = list(phys)
        self.norm = None
        if getattr(self.engine, 'bound2', None) is not None and not _all_unitary(circuit):
            self.engine.bound2 = None              # the norm has changed: measured again before the next run
        return self

    def _ensure_bound(self):
        """the squared norm of the whole state as a device scalar on the engine (one reduction + all-reduce when it is not
        known; unitary circuits keep it): libqcs ranges the fp16 tensor-core operands with it"""
        eng = self.engine
        if not hasattr(eng, 'bound2') or eng.bound2 is not None:
            return
        import torch.distributed as dist
        part = eng.norm2().clone()
        dist.all_reduce(part, group=self.group)
        eng.bound2 = part.repeat(2)

    # -- construction / export ------------------------------------------------------------------------------------
    @property
    def rank(self):
        """number of qubits (Tensor.rank of the reference, tensors.py:41-54); the process index is ``proc_rank``"""
        return self.n

    @property
    def shape(self):
        return (2,) * self.n

    @classmethod
    def zeros(cls, n, dtype=np.complex64, **kw):
        return cls(n, dtype, **kw).set_zeros()

    @classmethod
    def from_column_vector(cls, vec, dtype=np.complex64, **kw):
        """State.from_column_vector (state.py:32-60) for a sharded state: ``vec`` is the full 2^n vector; every rank
        passes the same array and keeps its own contiguous block of 2^(n-g) amplitudes."""
        vec = np.asarray(vec).reshape(-1)
        n = int(round(np.log2(vec.size)))
        if 2 ** n != vec.size:
            raise ValueError('Vector length should be a power of 2.')
        st = cls(n, dtype, **kw)
        block = 2 ** st.n_local
        st.phys = canonical_layout(st.n)
        st.engine.load_host(vec[st.proc_rank * block:(st.proc_rank + 1) * block])
        return st

    def shards_on_host(self):
        """every rank's block, on every rank (for states that fit a host: tests, small registers)"""
        import torch.distributed as dist
        shards = [None] * self.world
        dist.all_gather_object(shards, self.engine.to_host(), group=self.group)
        return shards

    def to_column_vector(self):
        """the whole normalised vector in the reference's order (qubit 0 = most significant bit, state.py:99-111)"""
        n = self.n
        if n > 28:
            raise ValueError('to_column_vector gathers 2^n amplitudes on every host: use marginals / measure at this size')
        full = np.concatenate(self.shards_on_host())
        idx = np.arange(2 ** n)
        src = np.zeros_like(idx)
        for q in range(n):
            src |= ((idx >> (n - 1 - q)) & 1) << self.phys[q]
        out = full[src]
        return out / np.sqrt(np.sum(np.abs(out) ** 2))

    @property
    def probabilities(self):
        """|amplitude|^2 as an array of shape (2,) * n (state.py:187-203); only for states that fit a host"""
        v = self.to_column_vector()
        return (np.abs(v) ** 2).reshape([2] * self.n)

    def copy(self):
        other = ShardedState.__new__(ShardedState)
        other.__dict__.update(self.__dict__)
        other.phys = list(self.phys)
        other.engine = self.engine.clone()
        return other

    # -- gates -------------------------------------------------------------------------------------------------------
    def apply_(self, op, qubit_indices=None):
        """op on the listed qubits, in place (the out-of-place form is Operator.__call__)"""
        from .circuit import Circuit
        c = Circuit(self.n)
        c.append(op, qubit_indices)
        return self.run(c)

    def _apply_operator(self, op, qubit_indices):
        """Operator.__call__(state, qubit_indices) (operators.py:280-323): a NEW state, the argument untouched"""
        return self.copy().apply_(op, qubit_indices)

    def _with_factor(self, other, other_first):
        from .state import State
        if not isinstance(other, State):
            return NotImplemented
        nb = other.rank
        out = self.copy()
        out.engine.kron_low(other.to_column_vector())
        mine, theirs = [p + nb for p in self.phys], [nb - 1 - j for j in range(nb)]
        out.phys = theirs + mine if other_first else mine + theirs
        out.n, out.n_local = self.n + nb, self.n_local + nb
        return out

    def __mul__(self, other):
        """tensor product with an ordinary State (tensors.py:56-76): the result's qubits are this state's, then the
        other's; physically the new qubits take the lowest local bits of every shard (only the qubit map knows)"""
        return self._with_factor(other, False)

    def __rmul__(self, other):
        return self._with_factor(other, True)

    # -- measurement -------------------------------------------------------------------------------------------------
    def _raw_marginals(self, qubits):
        """unnormalised marginal sums over the listed logical qubits (first = MSB), all-reduced over the ranks"""
        import torch.distributed as dist
        t = _lib.torch()
        m = len(qubits)
        pos = [self.phys[q] for q in qubits]
        local = [(j, p) for j, p in enumerate(pos) if p < self.n_local]
        part = self.engine.marginals([p for _, p in local]) if local else self.engine.norm2().reshape(1)
        full = t.zeros(1 << m, dtype=t.float64, device=part.device)
        idx = np.zeros(1 << len(local), dtype=np.int64)
        fixed = 0
        for j, p in enumerate(pos):
            if p >= self.n_local and (self.proc_rank >> (p - self.n_local)) & 1:
                fixed |= 1 << (m - 1 - j)
        for x in range(1 << len(local)):
            o = fixed
            for k, (j, _) in enumerate(local):
                if (x >> (len(local) - 1 - k)) & 1:
                    o |= 1 << (m - 1 - j)
            idx[x] = o
        full[t.from_numpy(idx).to(part.device)] = part.to(full.dtype)
        dist.all_reduce(full, group=self.group)
        return full.cpu().numpy()

    def marginal_probabilities(self, qubit_indices):
        """normalised marginal distribution of the listed qubits (first = most significant bit); all ranks get it"""
        raw = self._raw_marginals([int(q) for q in qubit_indices])
        return raw / raw.sum()

    def measure(self, qubit_indices=None, remove=False, u=None):
        """Measure the listed qubits, collapsing the state in place (state.py:307-387).  int argument -> int, else a
        tuple; the first listed qubit is the most significant bit of the outcome.  One uniform is drawn per call --
        from NumPy's global stream on rank 0 (like np.random.choice at state.py:364) and broadcast, or ``u``."""
        import torch.distributed as dist
        int_arg = False
        if isinstance(qubit_indices, (int, np.integer)):
            qubit_indices, int_arg = [int(qubit_indices)], True
        if qubit_indices is None:
            qubit_indices = range(self.n)
        qubit_indices = [int(q) for q in qubit_indices]
        if qubit_indices == []:
            raise ValueError('Must measure at least one qubit.')
        if min(qubit_indices) < 0 or max(qubit_indices) >= self.n:
            raise ValueError('Trying to measure qubit index i not 0<=i<d, '
                             'where d is the rank of the state vector.')
        if len(qubit_indices) != len(set(qubit_indices)):
            raise ValueError('Qubit indices list contains repeated elements.')
        m = len(qubit_indices)
        if m > 24:
            raise ValueError('measure at most 24 qubits of a sharded state per call')
        box = [float(np.random.random_sample()) if u is None else float(u)]
        dist.broadcast_object_list(box, src=0, group=self.group)
        raw = self._raw_marginals(qubit_indices)
        cdf = np.cumsum(raw)
        cdf = cdf / cdf[-1]
        outcome = min(int(np.searchsorted(cdf, box[0], side='right')), len(cdf) - 1)
        bits = tuple((outcome >> i) & 1 for i in range(m - 1, -1, -1))
        if remove:
            self._make_local(qubit_indices)
        scale = 1.0 / np.sqrt(raw[outcome])            # collapse, /sqrt(p) and the trailing renormalise in one factor
        pos = [self.phys[q] for q in qubit_indices]
        local = [(j, p) for j, p in enumerate(pos) if p < self.n_local]
        alive = all(((self.proc_rank >> (p - self.n_local)) & 1) == bits[j] for j, p in enumerate(pos) if p >= self.n_local)
        if not alive:
            self.engine.zero()                          # this shard disagrees with a measured rank bit
        elif local:
            lo = 0
            for j, _ in local:
                lo = (lo << 1) | bits[j]
            self.engine.collapse([p for _, p in local], lo, bool(remove), scale)
        else:
            self.engine.scale(scale)
        if remove:                                      # (_make_local: every measured qubit sits on a local bit)
            gone = sorted(pos)
            keep = [q for q in range(self.n) if q not in set(qubit_indices)]
            self.phys = [self.phys[q] - sum(1 for b in gone if b < self.phys[q]) for q in keep]
            self.n -= m
            self.n_local -= m
        self.norm = None
        return bits[0] if int_arg else bits

    def _make_local(self, qubits):
        """bring the listed logical qubits onto local physical bits (removal shrinks every shard alike): park them below
        the top g local bits, then trade the rank bits for those top bits if any listed qubit sits on a rank bit"""
        if all(self.phys[q] < self.n_local for q in qubits):
            return
        if len(qubits) > self.n_local - 2 * self.g:
            raise ValueError('too many qubits to remove from a state sharded this way')
        top = set(range(self.n_local - self.g, self.n_local))
        qubit_at = {p: q for q, p in enumerate(self.phys)}
        listed = set(qubits)
        movers = [q for q in qubits if self.phys[q] in top]
        safe = [b for b in range(self.n_local - self.g) if qubit_at[b] not in listed]
        src_bit = list(range(self.n_local))
        for q, b in zip(movers, safe):
            a = self.phys[q]
            other = qubit_at[b]
            src_bit[a], src_bit[b] = src_bit[b], src_bit[a]
            self.phys[q], self.phys[other] = b, a
            qubit_at[b], qubit_at[a] = q, other
        if movers:
            self.engine.permute(src_bit)
        self.engine.exchange(self.group)
        self.exchanges += 1
        qubit_at = {p: q for q, p in enumerate(self.phys)}
        for i in range(self.g):
            qa, qb = qubit_at[self.n_local + i], qubit_at[self.n_local - self.g + i]
            self.phys[qa], self.phys[qb] = self.n_local - self.g + i, self.n_local + i

    def norm2(self):
        """Squared norm of the whole state (all-reduce of the per-rank sums)."""
        import torch.distributed as dist
        part = self.engine.norm2()
        dist.all_reduce(part, group=self.group)
        return float(part[0].item())

    def marginals(self, qubits):
        """Marginal distribution of the listed logical qubits (first = MSB), normalised; all ranks get it."""
        import torch.distributed as dist
        t = _lib.torch()
        m = len(qubits)
        pos = [self.phys[q] for q in qubits]
        local = [(j, p) for j, p in enumerate(pos) if p < self.n_local]
        if local:
            part = self.engine.marginals([p for _, p in local])
        else:
            part = self.engine.norm2().reshape(1)
        full = t.zeros(1 << m, dtype=t.float64, device=part.device)
        idx = np.zeros(1 << len(local), dtype=np.int64)
        for x in range(1 << len(local)):
            o = 0
            for k, (j, _) in enumerate(local):
                if (x >> (len(local) - 1 - k)) & 1:
                    o |= 1 << (m - 1 - j)
            for j, p in enumerate(pos):
                if p >= self.n_local and (self.proc_rank >> (p - self.n_local)) & 1:
                    o |= 1 << (m - 1 - j)
            idx[x] = o
        full[t.from_numpy(idx).to(part.device)] = part
        dist.all_reduce(full, group=self.group)
        out = full.cpu().numpy()
        return out / out.sum()
