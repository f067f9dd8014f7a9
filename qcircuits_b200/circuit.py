"""Circuit: a gate list with a fusion scheduler (north_star (1): "runs of adjacent
gates fuse into one pass").

``Operator.__call__`` is one pass over the state per gate.  A Circuit collects the same
``(operator, qubit_indices)`` pairs and reorders / groups them -- only across gates that commute --
into passes that each fit one tile of the libqcs tile engine (include/qcs.h, qcs_apply_fused):
every dense target of a pass lies in the low ``low_bits`` flat bits or in a small set of high bits
chosen per pass; diagonal gates and control bits ride along wherever they sit.  The result is
identical to applying the gates one by one (up to floating-point reassociation of the lazy
normalisation) and is checked against the oracle in tests/.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._device import DeviceVector, make_ops
from .state import State


class _Gate:
    __slots__ = ('plan', 'bits', 'nd_mask', 'd_mask', 'entries', 'wide', 'op')

    def __init__(self, op, bits, plan=None):
        plan = op._plan() if plan is None else plan
        self.op, self.plan, self.bits = op, plan, bits
        tmask = 0
        for q in plan.targets:
            tmask |= 1 << bits[q]
        cmask = 0
        for q in plan.controls:
            cmask |= 1 << bits[q]
        # bits the gate acts on non-diagonally / diagonally (controls are diagonal on their bit)
        self.nd_mask = 0 if plan.diagonal else tmask
        self.d_mask = cmask | (tmask if plan.diagonal else 0)
        t = len(plan.targets)
        self.entries = (2 << t) if plan.diagonal else (1 << (2 * t))
        self.wide = t > _lib.QCS_MAX_OP_QUBITS


def tile_limits(dtype):
    """(low_bits, tile_bits) of the tile engine for this amplitude dtype (qcs_tile_limits)."""
    low, tile = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.load().qcs_tile_limits(_lib.dtype_code(dtype), ctypes.byref(low), ctypes.byref(tile)))
    return low.value, tile.value


def schedule_passes(gates, n, dtype, max_ops=128, lookahead=512, high_bits=None, low_shrink=None, tail_trim=None):
    """Greedy fusion: repeatedly pick the tile that admits the most pending gates, where a gate may overtake
    an earlier one only if they commute (they share no bit on which either acts non-diagonally).  A tile is
    the low L flat bits plus a window of H consecutive higher bits with L + H = tile_bits; L ranges from the
    engine's preferred contiguous run down to ``low_shrink`` bits less (a shorter contiguous run buys a wider
    window: for nearest-neighbour circuits the gates a window absorbs grow with the square of its width; fused
    passes are compute-bound, so 128-byte contiguous runs cost nothing measurable -- QCS_LOW_SHRINK, default 6).
    ``gates`` are records with nd_mask / d_mask / entries / wide over the n flat bits of the (local) state.
    Returns lists of indices into ``gates``."""
    low, tile = tile_limits(dtype)
    if low_shrink is None:
        low_shrink = int(os.environ.get('QCS_LOW_SHRINK', '6'))
    if tail_trim is None:
        tail_trim = int(os.environ.get('QCS_TAIL_TRIM', '8'))
    block_cut = int(os.environ.get('QCS_BLOCK_CUT', '0')) != 0
    geoms = []
    if high_bits is not None:
        geoms.append((min(low, n), max(0, min(high_bits, n - min(low, n)))))
    else:
        for shrink in range(0, low_shrink + 1):
            L = min(low - shrink, n)
            geoms.append((L, max(0, min(tile - L, n - L))))
            if L == n:
                break
    L0 = geoms[0][0]
    pool_cap = 12288 // (8 if np.dtype(dtype) == np.complex64 else 16) - 4
    max_ops = min(max_ops, _lib.QCS_MAX_OPS)
    remaining = list(range(len(gates)))
    passes = []
    full = (1 << n) - 1
    candidates = []
    for L, hcap in geoms:
        low_mask = (1 << L) - 1
        if low_mask not in candidates:
            candidates.append(low_mask)
        for s in range(L, n - hcap + 1):
            mask = low_mask | (((1 << hcap) - 1) << s)
            if mask not in candidates:
                candidates.append(mask)

    def scan(tile_mask):
        blk_nd = blk_d = 0
        taken, entries = [], 0
        for idx in remaining[:lookahead]:
            g = gates[idx]
            ok = (not g.wide and (g.nd_mask & ~tile_mask) == 0 and (g.nd_mask & (blk_nd | blk_d)) == 0
                  and (g.d_mask & blk_nd) == 0 and len(taken) < max_ops and entries + g.entries <= pool_cap)
            if ok:
                taken.append(idx)
                entries += g.entries
            else:
                blk_nd |= g.nd_mask
                blk_d |= g.d_mask
                if blk_nd == full:
                    break
        return taken

    while remaining:
        first = gates[remaining[0]]
        if first.wide:                      # operator wider than the tile engine: its own dense pass
            passes.append([remaining.pop(0)])
            continue
        cands = candidates
        # the first pending gate must always be schedulable: its own high bits, padded upwards
        low_mask, hcap = (1 << L0) - 1, geoms[0][1]
        need = first.nd_mask & ~low_mask
        if need and bin(need).count('1') <= hcap:
            m, b = need, L0
            while bin(m).count('1') < hcap and b < n:
                m |= 1 << b
                b += 1
            cands = candidates + [low_mask | m]
        best = None
        for mask in cands:
            taken = scan(mask)
            if best is None or len(taken) > len(best):
                best = taken
        if not best:
            # a gate whose dense targets exceed the preferred tile even alone: libqcs shrinks the low run
            best = [remaining[0]]
        if tail_trim > 0 and len(best) > 1:
            cut = _cut_after_last_block(gates, best, n, dtype) if block_cut else None
            best = cut if cut is not None else _trim_tail_sweep(gates, best, n, dtype, tail_trim)
        passes.append(best)
        chosen = set(best)
        remaining = [i for i in remaining if i not in chosen]
    # The hand-back of a short last sweep pays when a LATER pass absorbs those ops.  At the end of the circuit there is
    # none: a last pass of a few leftover ops costs a whole trip through HBM (2.8 ms at 30 qubits) where one more sweep of
    # the pass before costs about 1 ms -- it goes back if the planner can hold the union in one pass.
    if (tail_trim > 0 and len(passes) >= 2 and len(passes[-1]) <= tail_trim and len(passes[-2]) + len(passes[-1]) <= max_ops
            and not any(gates[i].wide for i in passes[-1] + passes[-2])):
        merged = passes[-2] + passes[-1]
        stats = (ctypes.c_int32 * 128)()
        ops, keep = make_ops([(gates[i].plan, gates[i].bits, False) for i in merged])
        if _lib.load().qcs_plan_stats(n, _lib.dtype_code(dtype), ops, len(ops), stats) == 0:
            passes[-2:] = [merged]
    return passes


def _commutes_past_the_rest(gates, group, drop):
    """True if every op of ``group`` listed in ``drop`` (positions) may run after all kept ops that follow it in
    program order: no shared bit on which either acts non-diagonally."""
    for k in sorted(drop):
        a = gates[group[k]]
        for j in range(k + 1, len(group)):
            if j in drop:
                continue
            b = gates[group[j]]
            if (a.nd_mask & (b.nd_mask | b.d_mask)) or (a.d_mask & b.nd_mask):
                return False
    return True


def _cut_after_last_block(gates, group, n, dtype):
    """Passes of the tensor-core kernel hold at most four dense blocks.  What the planner cannot put into them runs
    as register sweeps AFTER the last block, at ~1.3 ms per sweep + 0.1 ms per op (30 qubits) where a block costs
    ~1.2 ms for ~20 ops: those ops are handed back to the scheduler -- they are last in the pass's execution order, so
    they may open the next pass instead, where they join its blocks.  Returns the shortened group, or None when the
    pass has no block (complex128, small states) or nothing runs after its last block."""
    lib = _lib.load()
    stats = (ctypes.c_int32 * 128)()
    order = (ctypes.c_int32 * len(group))()
    ops, keep = make_ops([(gates[i].plan, gates[i].bits, False) for i in group])
    if lib.qcs_plan_order(n, _lib.dtype_code(dtype), ops, len(ops), stats, order) != 0:
        return None
    if (stats[7] & 255) == 0:
        return None
    last_block = max(o - 1000 for o in order if o >= 1000)
    drop = {k for k, o in enumerate(order) if 0 <= o < 1000 and o > last_block}
    if not drop or len(drop) >= len(group) or not _commutes_past_the_rest(gates, group, drop):
        return None if not drop else group
    return [g for k, g in enumerate(group) if k not in drop]


def _trim_tail_sweep(gates, group, n, dtype, max_tail):
    """A register sweep costs about as much as 15 register ops (the tile makes a round trip through shared
    memory behind a barrier).  When the engine's planner needs a LAST sweep for only a few leftover ops of a
    pass, those ops are handed back to the scheduler: they are last in the pass's execution order, so they may
    run at the start of a later pass instead, where they usually share a sweep with other gates.
    qcs_plan_stats (no device needed) reports the ops of the last sweep."""
    lib = _lib.load()
    stats = (ctypes.c_int32 * 128)()
    ops, keep = make_ops([(gates[i].plan, gates[i].bits, False) for i in group])
    if lib.qcs_plan_stats(n, _lib.dtype_code(dtype), ops, len(ops), stats) != 0:
        return group
    nsweeps, ntail = stats[1], stats[6]
    if nsweeps < 2 or ntail < 1 or ntail > min(max_tail, 12) or ntail >= len(group):
        return group
    drop = {stats[116 + i] for i in range(ntail)}
    # Deferring an op past the rest of the pass is only legal if every kept op that follows it in program
    # order commutes with it (same rule as the scheduler's: no shared bit on which either acts
    # non-diagonally).  The planner's report is trusted for WHICH ops are last, never for legality.
    for k in sorted(drop):
        a = gates[group[k]]
        for j in range(k + 1, len(group)):
            if j in drop:
                continue
            b = gates[group[j]]
            if (a.nd_mask & (b.nd_mask | b.d_mask)) or (a.d_mask & b.nd_mask):
                return group
    return [g for k, g in enumerate(group) if k not in drop]


def _cheap_form(M):
    """True if the tile engine has a cheap register body for the 2x2 matrix M (qcs_tile.cu arm table):
    diagonal (a phase), real (H, RY, X, Z ...) or real diagonal with imaginary off-diagonal (RX, Y)."""
    if M[0, 1] == 0 and M[1, 0] == 0:
        return True
    if not M.imag.any():
        return True
    return M[0, 0].imag == 0 and M[1, 1].imag == 0 and M[0, 1].real == 0 and M[1, 0].real == 0


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)


def expand_swaps(gates):
    """Swap = CNOT(a,b) CNOT(b,a) CNOT(a,b), exactly (all three are permutations): as controlled-X ops the
    swap runs in the register sweeps of a fused pass (moves only when both bits are register bits) instead
    of as a dense two-qubit op through shared memory."""
    from . import _gates
    out = []
    cnot = None
    for g in gates:
        p = g.plan
        if (not g.wide and not p.diagonal and not p.controls and len(p.targets) == 2
                and np.array_equal(p.matrix, _SWAP)):
            if cnot is None:
                cnot = _gates.analyze(_CNOT)
            a, b = g.bits[p.targets[0]], g.bits[p.targets[1]]
            out.extend([_Gate(None, [a, b], plan=cnot), _Gate(None, [b, a], plan=cnot), _Gate(None, [a, b], plan=cnot)])
        else:
            out.append(g)
    return out


def merge_single_qubit_gates(gates):
    """Peephole fusion before scheduling: an uncontrolled one-qubit gate waits in a per-bit accumulator;
    later one-qubit gates on the same bit multiply into it, gates that act *diagonally* on the bit
    (controls, diagonal targets) are let through while the accumulator is itself diagonal (they commute),
    and anything else flushes it.  Two gates are multiplied only while the product keeps a cheap register
    body (diagonal, real, RX-like: RZ.RZ, RX.RX, H.RY ...); H.RZ would be a general complex 2x2 that costs
    more instructions than H and the phase separately.  Returns a new gate list (same unitary up to
    rounding)."""
    from . import _gates
    out = []
    acc = {}          # bit -> 2x2 matrix

    def flush(b):
        M = acc.pop(b, None)
        if M is None:
            return
        if np.array_equal(M, np.eye(2)):
            return
        out.append(_Gate(None, [b], plan=_gates.analyze(M)))

    for g in gates:
        plan = g.plan
        if not g.wide and len(plan.targets) == 1 and not plan.controls:
            b = g.bits[plan.targets[0]]
            M = np.diag(plan.matrix) if plan.diagonal else plan.matrix
            M = np.array(M, dtype=np.complex128)
            if b in acc:
                prod = M @ acc[b]
                if _cheap_form(prod):
                    acc[b] = prod
                    continue
                flush(b)        # the product would need the general complex 2x2 body: keep the two cheap ops
            acc[b] = M
            continue
        nd = g.nd_mask
        for b in list(acc):
            bit = 1 << b
            if nd & bit:
                flush(b)
            elif g.d_mask & bit:
                M = acc[b]
                if M[0, 1] != 0 or M[1, 0] != 0:
                    flush(b)
        out.append(g)
    for b in sorted(acc):
        flush(b)
    return out


def build_passes(gates, groups, n, dtype):
    """ctypes op arrays for the scheduled groups, each checked against the engine's pass planner
    (qcs_plan_stats plans without launching and needs no device).  A group the planner cannot hold in one
    pass -- too many register ops, sweeps or coefficients for its tables, which the scheduler only
    estimates -- is split in halves until it fits; a single op always does.
    Returns [(ops, keepalive, gate count)]."""
    lib = _lib.load()
    code = _lib.dtype_code(dtype)
    stats = (ctypes.c_int32 * 128)()
    out = []

    def emit(group):
        ops, keep = make_ops([(gates[i].plan, gates[i].bits, False) for i in group])
        rc = lib.qcs_plan_stats(n, code, ops, len(ops), stats)
        if rc == 0 or len(group) == 1:
            if rc != 0:
                _lib.check(rc)
            out.append((ops, keep, len(group)))
            return
        half = len(group) // 2
        emit(group[:half])
        emit(group[half:])

    for group in groups:
        emit(list(group))
    return out


class Circuit:
    """An ordered list of gates on ``n`` qubits.

    >>> c = Circuit(3)
    >>> c.append(qc.Hadamard(), [0]); c.append(qc.CNOT(), [0, 1])
    >>> out = c.run(qc.zeros(3))            # == CNOT(H(zeros, [0]), [0, 1])
    """

    def __init__(self, n):
        self.n = int(n)
        self.gates = []
        self._plans = {}

    def append(self, op, qubit_indices=None):
        from .operators import _validated_indices
        qi = _validated_indices(op.rank // 2, self.n, qubit_indices)
        self.gates.append(_Gate(op, [self.n - 1 - q for q in qi]))
        self._plans.clear()
        return self

    def __len__(self):
        return len(self.gates)

    # -- scheduling ------------------------------------------------------------------------------------
    def schedule(self, dtype=np.complex128, fuse=True, max_ops=128, lookahead=512):
        """Group the gates into passes.  Returns a list of lists of gate indices (execution order
        inside a pass = list order).  ``fuse=False`` gives one pass per gate."""
        if not fuse:
            return [[i] for i in range(len(self.gates))]
        return schedule_passes(self.gates, self.n, dtype, max_ops=max_ops, lookahead=lookahead)

    def _compiled(self, dtype, fuse, max_ops):
        key = (np.dtype(dtype).str, fuse, max_ops)
        comp = self._plans.get(key)
        if comp is None:
            comp = []
            gates = merge_single_qubit_gates(expand_swaps(self.gates)) if fuse else self.gates
            groups = (schedule_passes(gates, self.n, dtype, max_ops=max_ops) if fuse
                      else [[i] for i in range(len(gates))])
            for group in groups:
                g0 = gates[group[0]]
                if g0.wide:
                    comp.append(('dense', g0))
                else:
                    comp.extend(('tile', ops, keep, cnt) for ops, keep, cnt in build_passes(gates, [group], self.n, dtype))
            self._plans[key] = comp
        return comp

    # -- execution ---------------------------------------------------------------------------------------
    def run(self, state, fuse=True, max_ops=128):
        """Apply the circuit to ``state`` and return the new State; the argument is left untouched
        (the first pass writes a fresh buffer, later passes update it in place)."""
        if state.rank != self.n:
            raise ValueError('circuit and state have different numbers of qubits')
        src = state._dv
        if not self.gates:
            return State(src.clone())
        lib = _lib.load()
        ws = _lib.workspace().data_ptr()
        stream = _lib.stream_ptr()
        code = src.code
        cur = DeviceVector(_lib.alloc_amplitudes(src.nbits, src.dtype), None, src.nbits, src.dtype)
        norms = [_lib.alloc_scalar(), _lib.alloc_scalar()]
        norm_in = None if src.norm is None else src.norm.data_ptr()
        # a state whose norm is known without measuring (a basis state) ranges the first pass's tensor-core operands
        bound = src.bound.data_ptr() if (src.norm is None and src.bound is not None) else None
        src_ptr, which = src.buf.data_ptr(), 0
        for item in self._compiled(src.dtype, fuse, max_ops):
            norm_out = norms[which].data_ptr()
            if item[0] == 'tile':
                _lib.check(lib.qcs_apply_fused_bounded(cur.buf.data_ptr(), src_ptr, self.n, code, item[1], len(item[1]),
                                                       norm_in, norm_out, bound, ws, stream))
            else:
                g = item[1]
                if src_ptr == cur.buf.data_ptr():      # the dense path is out of place
                    tmp = cur.buf.clone()
                    src_ptr = tmp.data_ptr()
                _lib.check(lib.qcs_apply_dense(cur.buf.data_ptr(), src_ptr, self.n, code,
                                               g.op._device_matrix(src.dtype).data_ptr(), g.plan.k,
                                               _lib.int32_array(g.bits), norm_in, norm_out, ws, stream))
            src_ptr, norm_in, bound = cur.buf.data_ptr(), norm_out, None
            cur.norm = norms[which]
            which ^= 1
        return State(cur)

    def run_host(self, host_in, host_out, fuse=True, max_ops=128):
        """Host-buffer entry: ``host_in`` / ``host_out`` are (pinned) torch CPU tensors holding 2^n
        complex amplitudes (complex64/complex128 or interleaved reals).  Copies the state to the
        device, runs the circuit, copies the normalised result back.  Returns ``host_out``."""
        t = _lib.require_cuda()
        real_in = t.view_as_real(host_in).reshape(-1) if host_in.is_complex() else host_in.reshape(-1)
        real_out = t.view_as_real(host_out).reshape(-1) if host_out.is_complex() else host_out.reshape(-1)
        np_dtype = np.complex64 if real_in.dtype == t.float32 else np.complex128
        dev = _lib.alloc_amplitudes(self.n, np_dtype)
        dev.copy_(real_in, non_blocking=True)
        out = self.run(State(DeviceVector(dev, None, self.n, np_dtype)), fuse=fuse, max_ops=max_ops)
        res = out._dv.scaled(1.0)            # fold the lazy 1/sqrt(norm) before leaving the device
        real_out.copy_(res.buf, non_blocking=True)
        t.cuda.current_stream().synchronize()
        return host_out

    def stats(self, dtype=np.complex128, fuse=True, max_ops=128):
        """(number of passes, number of gates) for the schedule used by run()."""
        return len(self._compiled(dtype, fuse, max_ops)), len(self.gates)

    def plan_summary(self, dtype=np.complex128, fuse=True, max_ops=128):
        """How the engine will run the compiled passes (qcs_plan_stats, no device needed): passes, passes on the
        tensor-core kernel, sweeps, tensor-core blocks, gates inside blocks, register ops."""
        lib = _lib.load()
        stats = (ctypes.c_int32 * 128)()
        out = {'passes': 0, 'tensor_core_passes': 0, 'sweeps': 0, 'blocks': 0, 'gates_in_blocks': 0, 'register_ops': 0,
               'dense_passes': 0}
        for item in self._compiled(dtype, fuse, max_ops):
            out['passes'] += 1
            if item[0] != 'tile':
                out['dense_passes'] += 1
                continue
            _lib.check(lib.qcs_plan_stats(self.n, _lib.dtype_code(dtype), item[1], len(item[1]), stats))
            nb = stats[7] & 255
            out['tensor_core_passes'] += 1 if nb else 0
            out['sweeps'] += stats[1]
            out['blocks'] += nb
            out['gates_in_blocks'] += stats[7] >> 8
            out['register_ops'] += stats[2] - stats[1]
        return out
