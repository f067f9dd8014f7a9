"""State and the state factories -- the drop-in for reference qcircuits/state.py.

A State owns a device buffer of 2^d amplitudes (flat C order: qubit 0 is the most significant
bit, reference state.py:60,111) instead of the NumPy rank-d tensor ``_t``.  ``_t`` survives as a
property that copies the represented tensor to the host (the boundary the reference's tests read),
and assigning to it uploads.  Amplitude arithmetic -- gate passes, probabilities, measurement,
tensor product, sums -- is done by the sm_100a kernels of libqcs; nothing here falls back to NumPy
for amplitude work.

Extension over the reference (which hard-codes complex128, tensors.py:18): constructors and
factories accept ``dtype=np.complex64``.
"""
import os

import numpy as np

from . import _lib
from ._device import DeviceVector, make_ops
from .tensors import Tensor


class State(Tensor):
    """The state of a d-qubit system: amplitudes of shape (2,) * d."""

    def __init__(self, tensor, dtype=np.complex128):
        if isinstance(tensor, DeviceVector):
            self._dv = tensor
        else:
            # private complex128 copy semantics of the reference (tensors.py:17-18)
            self._dv = DeviceVector.from_host(np.array(tensor, dtype=np.complex128), dtype)

    # -- the host view -------------------------------------------------------------------------------
    @property
    def _t(self):
        return self._dv.to_host().reshape([2] * self._dv.nbits)

    @_t.setter
    def _t(self, tensor):
        self._dv = DeviceVector.from_host(np.array(tensor, dtype=np.complex128), self._dv.dtype)

    @property
    def shape(self):
        return (2,) * self._dv.nbits

    @property
    def rank(self):
        return self._dv.nbits

    @property
    def dtype(self):
        return self._dv.dtype

    def __deepcopy__(self, memo):
        return State(self._dv.clone())

    def __copy__(self):
        return State(self._dv.clone())

    @staticmethod
    def from_column_vector(v, dtype=np.complex128):
        """Column vector of size 2^d (Kronecker convention) -> State (state.py:32-60)."""
        if type(v) is list:
            v = np.array(v, dtype=np.complex128)
        d = np.log2(v.size)
        if not float(d).is_integer():
            raise ValueError('Vector size should be a power of 2.')
        return State(np.reshape(v, [2] * int(d)), dtype=dtype)

    def to_column_vector(self):
        """State -> column vector of size 2^d (state.py:99-111)."""
        return self._dv.to_host()

    def __repr__(self):
        head = 'State('
        return head + str(self._t).replace('\n', '\n' + ' ' * len(head)) + ')'

    def __str__(self):
        return '{}-qubit state. Tensor:\n{}'.format(self.rank, self._t)

    # -- arithmetic --------------------------------------------------------------------------------------
    def _same_kind(self, arg):
        if arg._dv.dtype != self._dv.dtype:
            raise TypeError('states have different amplitude dtypes')
        return arg._dv

    def dot(self, arg):
        """<self|arg> (state.py:74-89)."""
        if arg.rank != self.rank:
            raise ValueError('states have different numbers of qubits')
        return self._dv.dot(self._same_kind(arg))

    def renormalize_(self):
        """Scale to unit length, in place (state.py:91-97).  The division is deferred: the squared
        norm is attached to the buffer and folded into the next kernel that reads it."""
        if self._dv.norm is None:
            self._dv.norm = self._dv.norm2()

    def __add__(self, arg):
        if arg.rank != self.rank:
            raise ValueError('states have different numbers of qubits')
        return State(self._dv.lincomb(self._same_kind(arg), 1.0, 1.0))

    def __sub__(self, arg):
        return self + (-1) * arg

    def __neg__(self):
        return State(self._dv.scaled(-1.0))

    def __mul__(self, scalar):
        if isinstance(scalar, (float, int, complex)):
            return State(self._dv.scaled(scalar))
        if type(scalar).__name__ == 'ShardedState':        # small (x) sharded: the sharded operand builds the product
            return scalar.__rmul__(self)
        return self.tensor_product(scalar)

    def __rmul__(self, scalar):
        if isinstance(scalar, (float, int, complex)):
            return State(self._dv.scaled(scalar))

    def __truediv__(self, scalar):
        return State(self._dv.scaled(1.0 / complex(scalar)))

    def tensor_product(self, arg):
        """self (x) arg, self's qubits first (tensors.py:56-76)."""
        return State(self._dv.kron(self._same_kind(arg)))

    def tensor_power(self, n):
        acc = self
        for _ in range(n - 1):
            acc = acc.tensor_product(self)
        return State(acc._dv.clone()) if n <= 1 else acc

    # -- qubit relabelling -------------------------------------------------------------------------------
    def _permuted_tensor(self, axes, inverse=False):
        if inverse:
            axes = np.argsort(axes)
        return np.transpose(self._t, axes)

    def permute_qubits(self, axes, inverse=False):
        """Relabel the qubits: new qubit i is old qubit axes[i] (state.py:119-140).  In place."""
        axes = [int(a) for a in axes]
        d = self.rank
        if sorted(axes) != list(range(d)):
            raise ValueError("axes don't match array")
        if inverse:
            axes = [int(a) for a in np.argsort(axes)]
        src_bit = [0] * d
        for new_q, old_q in enumerate(axes):
            src_bit[d - 1 - new_q] = d - 1 - old_q
        self._dv = self._dv.permuted(src_bit)
        return self

    def swap_qubits(self, axis1, axis2):
        """Swap two qubits, in place (state.py:142-163)."""
        axes = list(range(self.rank))
        axes[axis1], axes[axis2] = axes[axis2], axes[axis1]
        return self.permute_qubits(axes)

    # -- probabilities -----------------------------------------------------------------------------------
    @property
    def probabilities(self):
        """Probability of each computational basis vector, float64 of shape (2,) * d (state.py:187-203)."""
        probs, total = self._dv.probabilities()
        assert abs(total - 1.0) < 1e-4, ('State probabilities'
                                         'do not sum to 1.')
        return probs.reshape([2] * self.rank)

    def _probability_sum(self):
        if self._dv.norm is not None:
            return 1.0
        return float(self._dv.norm2()[0].item())

    @property
    def amplitudes(self):
        """Read-only copy of the amplitudes (state.py:205-221)."""
        total = self._probability_sum()
        assert abs(total - 1.0) < 1e-4, ('State probabilities'
                                         'do not sum to 1.')
        amp = self._t
        amp.flags.writeable = False
        return amp

    def schmidt_number(self, indices):
        """Schmidt number across the cut (indices | rest) (state.py:225-260).  The reference transposes the tensor so
        that the listed qubits come first, reshapes it into a 2^a x 2^b matrix and counts its singular values above
        1e-10.  Here the relabelling is the device bit permutation (qcs_permute_bits), the matrix is a view of the
        permuted buffer and the singular values come from the device solver (torch.linalg.svdvals in float64, the
        reference's precision): nothing of the state goes through the host."""
        if len(indices) in [0, self.rank]:
            raise ValueError('At least one qubit index should be included '
                             'and at least one should be excluded')
        if min(indices) < 0 or max(indices) >= self.rank:
            raise ValueError('Indices should be between 0 and d-1 for a d-qubit state.')
        if not all([isinstance(idx, int) for idx in indices]):
            raise ValueError('Indices should be integers.')
        inside = set(indices)
        outside = set(range(self.rank)) - inside
        d = self.rank
        axes = list(inside) + list(outside)
        src_bit = [0] * d
        for new_q, old_q in enumerate(axes):
            src_bit[d - 1 - new_q] = d - 1 - old_q
        dv = self._dv.permuted(src_bit)
        t = _lib.torch()
        M = t.view_as_complex(dv.buf.view(-1, 2)).reshape(2 ** len(inside), 2 ** len(outside)).to(t.complex128)
        if dv.norm is not None:
            M = M / t.sqrt(dv.norm[0])
        return int((t.linalg.svdvals(M) > 1e-10).sum().item())

    # -- application of an operator (called by Operator._apply) -------------------------------------------
    def _apply_operator(self, op, qubit_indices):
        """op applied to the listed qubits -> new State, renormalised (operators.py:255-276).
        Operator qubit j acts on flat bit d-1-qubit_indices[j]."""
        d = self.rank
        plan = op._plan()
        bits = [d - 1 - q for q in qubit_indices]
        if len(plan.targets) <= _lib.QCS_MAX_OP_QUBITS:
            ops, keep = make_ops([(plan, bits, False)])
            return State(self._dv.apply_ops(ops, keep))
        return State(self._dv.apply_dense(op._device_matrix(self._dv.dtype), plan.k, bits))

    def _apply_uf(self, op, qubit_indices):
        """U_f on the listed qubits as an index permutation (operators.UfOperator; reference operators.py:809-853 +
        :261): the first d-1 listed qubits are f's arguments (first = most significant), the last one is flipped."""
        d = self.rank
        bits = [d - 1 - q for q in qubit_indices]
        src = self._dv
        dst = src._new_like()
        _lib.check(_lib.load().qcs_apply_uf(dst.buf.data_ptr(), src.buf.data_ptr(), d, src.code, op._device_table().data_ptr(),
                                            _lib.int32_array(bits[:-1]), len(bits) - 1, bits[-1],
                                            None if src.norm is None else src.norm.data_ptr(), _lib.stream_ptr()))
        return State(dst)

    # -- measurement -----------------------------------------------------------------------------------------
    def marginal_probabilities(self, qubit_indices):
        """float64 marginal distribution of the listed qubits (first = most significant bit of the outcome): the
        ``ps`` of the reference's State._measurement_probabilities (state.py:262-269), from the marginal kernel."""
        d = self.rank
        return self._dv.marginals([d - 1 - int(q) for q in qubit_indices])

    def _measurement_probabilities(self, qubit_indices):
        """(ps, num_outcomes, amplitudes, permute) exactly like the reference's private helper (state.py:262-269; its
        own tests unpack all four, tests.py:1051).  ``amplitudes`` -- the state transposed so that the measured qubits
        lead -- is a host copy made only when somebody reads it."""
        qubit_indices = list(qubit_indices)
        ps = self.marginal_probabilities(qubit_indices)
        unmeasured = list(set(range(self.rank)) - set(qubit_indices))
        permute = qubit_indices + unmeasured
        return ps, 2 ** len(qubit_indices), _LazyTranspose(self, permute), permute

    def measure(self, qubit_indices=None, remove=False, u=None):
        """Measure the listed qubits in the computational basis, collapsing this state in place
        (state.py:307-387).  Returns an int for an int argument, else a tuple of ints; the first
        listed qubit is the most significant bit of the sampled outcome.

        ``u`` (extension): a uniform draw in [0, 1) to use instead of NumPy's global stream.  By
        default the outcome is drawn with ``np.random.choice`` exactly as the reference does, so a
        seeded run reproduces the reference's outcomes."""
        total = self._probability_sum()
        assert abs(total - 1.0) < 1e-4, ('State probabilities'
                                         'do not sum to 1.')
        int_arg = False
        if isinstance(qubit_indices, int):
            qubit_indices = [qubit_indices]
            int_arg = True
        if qubit_indices is None:
            qubit_indices = range(self.rank)
        qubit_indices = list(qubit_indices)
        if qubit_indices == []:
            raise ValueError('Must measure at least one qubit.')
        if min(qubit_indices) < 0 or max(qubit_indices) >= self.rank:
            raise ValueError('Trying to measure qubit index i not 0<=i<d, '
                             'where d is the rank of the state vector.')
        if len(qubit_indices) != len(set(qubit_indices)):
            raise ValueError('Qubit indices list contains repeated elements.')

        d = self.rank
        bitpos = [d - 1 - int(q) for q in qubit_indices]
        m = len(qubit_indices)
        if m <= MEASURE_CHUNK_BITS:
            ps = self._dv.marginals(bitpos)
            outcome = _sample(ps, u)
        else:
            # many qubits (the default measure() of a large state): 2^m marginals would not fit the host, so
            # the same inverse-CDF draw is taken hierarchically, MEASURE_CHUNK_BITS qubits at a time
            work = [self._dv]

            def conditional(done_bits, done_outcome, chunk_bits):
                if done_bits:
                    work[0] = work[0].collapsed(done_bits, done_outcome, False)
                return work[0].marginals(chunk_bits)

            outcome = _sample_chunked(conditional, bitpos, u, MEASURE_CHUNK_BITS)
        bits = tuple((outcome >> i) & 1 for i in range(m - 1, -1, -1))
        # collapse + renormalise: the kernel keeps the matching amplitudes and reports their squared
        # norm, which becomes the lazy scale (covers both the /sqrt(p) and the trailing renormalize_)
        self._dv = self._dv.collapsed(bitpos, outcome, remove)
        return bits[0] if int_arg else bits

    # -- density operators ---------------------------------------------------------------------------------------
    def density_operator(self):
        """|psi><psi| (state.py:271-283)."""
        from .density_operator import DensityOperator
        return DensityOperator.from_ensemble([self], [1.])

    def reduced_density_operator(self, qubit_indices):
        """Reduced density operator of the listed qubits (state.py:285-305)."""
        return self.density_operator().reduced_density_operator(qubit_indices)


class _LazyTranspose:
    """np.transpose(state._t, permute), materialised on first use (a full device -> host copy)"""

    def __init__(self, state, permute):
        self._state, self._permute, self._value = state, permute, None

    def _get(self):
        if self._value is None:
            self._value = np.transpose(self._state._t, self._permute)
        return self._value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype=dtype)

    def __getitem__(self, key):
        return self._get()[key]

    @property
    def shape(self):
        return (2,) * self._state.rank


MEASURE_CHUNK_BITS = 16


def _sample_chunked(conditional, bitpos, u, chunk):
    """Inverse-CDF sampling of an outcome over ``bitpos`` (first = most significant) without ever holding the
    2^m probabilities: the outcome of the first ``chunk`` bits is drawn from their marginals, u is rescaled to
    the position inside that outcome's probability interval, and the next chunk is drawn from the marginals
    conditioned on what has been drawn -- in exact arithmetic the same index as
    ``searchsorted(cumsum(p) / sum(p), u, 'right')`` over all 2^m outcomes (state.py:364, np.random.choice).
    ``conditional(done_bits, done_outcome, chunk_bits)`` returns the marginals of ``chunk_bits`` given that the
    previous chunk ``done_bits`` came out as ``done_outcome`` (and everything before it as drawn).
    ``u`` None draws the one uniform np.random.choice would draw from NumPy's global stream."""
    if u is None:
        u = float(np.random.random_sample())
    outcome = 0
    done_bits, done_outcome = [], 0
    for start in range(0, len(bitpos), chunk):
        bits = bitpos[start:start + chunk]
        ps = np.asarray(conditional(done_bits, done_outcome, bits), dtype=np.float64)
        cdf = np.cumsum(ps)
        cdf /= cdf[-1]
        o = min(int(np.searchsorted(cdf, u, side='right')), len(cdf) - 1)
        lo = cdf[o - 1] if o > 0 else 0.0
        width = cdf[o] - lo
        u = (u - lo) / width if width > 0 else 0.0
        u = min(max(u, 0.0), float(np.nextafter(1.0, 0.0)))
        outcome = (outcome << len(bits)) | o
        done_bits, done_outcome = bits, o
    return outcome


def _sample(ps, u=None):
    """Outcome index from the marginals: ``np.random.choice`` (one uniform from the global legacy
    stream, like state.py:364), or the same inverse-CDF rule on a caller-supplied uniform."""
    if u is None:
        # complex64 amplitudes sum to 1 only to ~1e-7, np.random.choice wants 1.5e-8: hand it the
        # normalised vector (its own cdf /= cdf[-1] makes this a no-op for the outcome)
        return int(np.random.choice(len(ps), p=ps / ps.sum()))
    cdf = np.cumsum(ps)
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side='right'))


# -----------------------------------------------------------------------------------------------------
# factories (reference state.py:394-618)
# -----------------------------------------------------------------------------------------------------
def qubit(*, alpha=None, beta=None, theta=None, phi=None, global_phase=0.0, dtype=np.complex128):
    """One qubit from amplitudes (alpha, beta) or Bloch angles (theta, phi)."""
    have_ab = alpha is not None and beta is not None
    have_tp = theta is not None and phi is not None
    if have_ab and theta is None and phi is None:
        if abs(np.real(np.conj(alpha) * alpha + np.conj(beta) * beta)) - 1 > 1e-4:
            raise ValueError('Qubit probabilities should sum to 1.')
    elif have_tp and alpha is None and beta is None:
        alpha = np.cos(theta / 2)
        beta = np.sin(theta / 2) * np.exp(1j * phi)
    else:
        raise ValueError('Incorrect combination of arguments for qubit. '
                         'Supply alpha and beta, or theta and phi.')
    return State(np.array([alpha, beta], dtype=np.complex128) * np.exp(1j * global_phase), dtype=dtype)


def _basis(d, index, dtype):
    return State(DeviceVector.basis(d, index, dtype))


SHARD_MIN_QUBITS = int(os.environ.get('QCS_SHARD_MIN_QUBITS', '32'))


def _wants_sharding(d, sharded):
    """north_star (4): states of 32 qubits or more are sharded over the ranks of the process group (one process per
    GPU, torchrun); ``sharded=True / False`` overrides, QCS_SHARD_MIN_QUBITS moves the threshold."""
    if sharded is not None:
        return bool(sharded)
    if d < SHARD_MIN_QUBITS:
        return False
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def zeros(d=1, dtype=np.complex128, sharded=None):
    """|0...0> on d qubits: a device memset plus one store.  In a multi-GPU program (torch.distributed initialised,
    one process per GPU) a state of >= 32 qubits comes back as a dist.ShardedState."""
    if d < 1:
        raise ValueError('Rank must be at least 1.')
    if _wants_sharding(d, sharded):
        from .dist import ShardedState
        return ShardedState.zeros(d, dtype)
    return _basis(d, 0, dtype)


def ones(d=1, dtype=np.complex128):
    """|1...1> on d qubits."""
    if d < 1:
        raise ValueError('Rank must be at least 1.')
    return _basis(d, (1 << d) - 1, dtype)


def bitstring(*bits, dtype=np.complex128):
    """Computational basis state with the given bits (first bit = qubit 0)."""
    d = len(bits)
    if d == 0:
        raise ValueError('Rank must be at least 1.')
    index = 0
    for b in bits:
        if b not in (0, 1):
            raise IndexError('bits must be 0 or 1')
        index = (index << 1) | int(b)
    return _basis(d, index, dtype)


def positive_superposition(d=1, dtype=np.complex128):
    """|+>^d, i.e. H on every qubit of |0...0> (d one-qubit passes instead of a 2^d x 2^d operator)."""
    if d < 1:
        raise ValueError('Rank must be at least 1.')
    from .operators import Hadamard
    H = Hadamard()
    x = zeros(d, dtype=dtype)
    for q in range(d):
        x = H(x, qubit_indices=[q])
    return x


def bell_state(a=0, b=0, dtype=np.complex128):
    """One of the four Bell states."""
    if a not in [0, 1] or b not in [0, 1]:
        raise ValueError('Bell state arguments are bits, and must be 0 or 1.')
    from .operators import CNOT, Hadamard
    phi = Hadamard()(bitstring(a, b, dtype=dtype), qubit_indices=[0])
    return CNOT()(phi)
