"""DensityOperator -- the drop-in for reference qcircuits/density_operator.py.

The rank-2d tensor rho[r0, c0, r1, c1, ...] (interleaved, operators.py:69-73) lives on the device
as a flat vector of 2^(2d) amplitudes: the row bit of qubit q is flat bit 2(d-1-q)+1, its column
bit is 2(d-1-q).  U rho U^dagger (operators.py:317-321) is then the ordinary gate pass applied
twice inside ONE fused kernel launch -- U on the row bits, conj(U) on the column bits -- with no
conjugate-transpose copies; probabilities are a strided gather of the diagonal; measurement
reuses the State marginal / collapse kernels on the doubled bits.
"""
import numpy as np

from . import _lib
from ._device import DeviceVector, make_ops
from .operators import OperatorBase


class DensityOperator(OperatorBase):
    """A mixed state of d qubits."""

    def __init__(self, tensor, dtype=np.complex128):
        if isinstance(tensor, DeviceVector):
            self._dv = tensor
        else:
            arr = np.array(tensor, dtype=np.complex128)
            if arr.ndim % 2:
                raise ValueError('a density operator tensor has even rank')
            self._dv = DeviceVector.from_host(arr, dtype)

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
        return DensityOperator(self._dv.clone())

    def __copy__(self):
        return DensityOperator(self._dv.clone())

    def __repr__(self):
        head = 'DensityOperator('
        return head + str(self._t).replace('\n', '\n' + ' ' * len(head)) + ')'

    def __str__(self):
        return 'Density operator for {}-qubit state space. Tensor:\n{}'.format(self.rank // 2, self._t)

    # -- qubit relabelling on the device (both wires of each qubit move together) ----------------------
    def permute_qubits(self, axes, inverse=False):
        axes = [int(a) for a in axes]
        d = self.rank // 2
        if sorted(axes) != list(range(d)):
            raise ValueError("axes don't match array")
        if inverse:
            axes = [int(a) for a in np.argsort(axes)]
        src_bit = [0] * (2 * d)
        for new_q, old_q in enumerate(axes):
            for wire in (0, 1):
                src_bit[2 * (d - 1 - new_q) + wire] = 2 * (d - 1 - old_q) + wire
        self._dv = self._dv.permuted(src_bit)
        return self

    def swap_qubits(self, axis1, axis2):
        axes = list(range(self.rank // 2))
        axes[axis1], axes[axis2] = axes[axis2], axes[axis1]
        return self.permute_qubits(axes)

    # -- probabilities -------------------------------------------------------------------------------------
    def _diagonal(self):
        """device float64 vector of Re rho[i, i] and its sum"""
        t = _lib.torch()
        d = self.rank // 2
        probs = t.empty(1 << d, dtype=t.float64, device='cuda')
        total = _lib.alloc_scalar()
        _lib.check(_lib.load().qcs_diag_probs(probs.data_ptr(), self._dv.buf.data_ptr(), d, self._dv.code,
                                              total.data_ptr(), _lib.workspace().data_ptr(), _lib.stream_ptr()))
        return probs, total

    @property
    def probabilities(self):
        """Probability of each computational basis vector (density_operator.py:47-63)."""
        probs, total = self._diagonal()
        assert abs(float(total[0].item()) - 1.0) < 1e-4, ('State probabilities'
                                                          'do not sum to 1.')
        return probs.cpu().numpy().reshape([2] * (self.rank // 2))

    # -- construction -----------------------------------------------------------------------------------------
    @staticmethod
    def from_ensemble(states, ps=None):
        """sum_i p_i |psi_i><psi_i| (density_operator.py:65-117)."""
        if ps is None:
            ps = np.ones(len(states)) / len(states)
        if len(ps) != len(states):
            raise ValueError('Probabilities list ps should be '
                             'same length as states list.')
        if len(states) == 0:
            raise ValueError('Provide at least one state.')
        if not np.isclose(np.sum(ps), 1):
            raise ValueError('Probabilities should sum to 1.')
        d = states[0].rank
        dtype = states[0].dtype
        if d == 0:
            # rank-0 states (what is left after measuring and removing the last qubit; the reference's own tests
            # build density operators of them, tests.py:1146-1190): the 1 x 1 "matrix" sum p |a|^2 / <a|a>, on the host
            total = 0.0
            for p, s in zip(ps, states):
                if s.rank != 0:
                    raise ValueError('Pure state dimensionalities do not match.')
                a = complex(np.asarray(s._t).reshape(-1)[0])
                total += float(p) * (a * np.conj(a)).real
            return DensityOperator(np.array(total, dtype=np.complex128), dtype=dtype)
        rho = DeviceVector(_lib.alloc_amplitudes(2 * d, dtype), None, 2 * d, dtype)
        lib = _lib.load()
        for i, (p, s) in enumerate(zip(ps, states)):
            if s.rank != d:
                raise ValueError('Pure state dimensionalities do not match.')
            if s.dtype != dtype:
                # (the dtype= extension: a float2 buffer must never be read as double2)
                raise TypeError('states of an ensemble must share one amplitude dtype (%s vs %s)' % (s.dtype, dtype))
            norm = s._dv.norm
            _lib.check(lib.qcs_outer_accumulate(rho.buf.data_ptr(), s._dv.buf.data_ptr(), d, rho.code, float(p),
                                                None if norm is None else norm.data_ptr(), 1 if i else 0,
                                                _lib.stream_ptr()))
        return DensityOperator(rho)

    @staticmethod
    def _tensor_from_state_outer_product(state):
        return DensityOperator.from_ensemble([state], [1.])._t

    # -- U rho U^dagger (called by Operator._apply) ----------------------------------------------------------------
    def _apply_operator(self, op, qubit_indices):
        d = self.rank // 2
        plan = op._plan()
        row = [2 * (d - 1 - q) + 1 for q in qubit_indices]
        col = [2 * (d - 1 - q) for q in qubit_indices]
        if len(plan.targets) <= _lib.QCS_MAX_OP_QUBITS:
            ops, keep = make_ops([(plan, row, False), (plan, col, True)])
            try:
                return DensityOperator(self._dv.apply_ops(ops, keep, track_norm=False))
            except _lib.QcsError:
                # the two halves together touch too many high bits for one tile: two passes
                first, keep1 = make_ops([(plan, row, False)])
                second, keep2 = make_ops([(plan, col, True)])
                half = self._dv.apply_ops(first, keep1, track_norm=False)
                return DensityOperator(half.apply_ops(second, keep2, track_norm=False))
        half = self._dv.apply_dense(op._device_matrix(self._dv.dtype), plan.k, row, track_norm=False)
        return DensityOperator(half.apply_dense(op._device_matrix(self._dv.dtype, conj=True), plan.k, col,
                                                track_norm=False))

    # -- partial trace / purification (host: SURVEY 8f-4, LAPACK-class work outside the hot path) ---------------
    def _reduced_device(self, retain_indices):
        """Trace out every qubit not listed (density_operator.py:127-138) on the device: qcs_partial_trace
        replaces the two transposes and the np.einsum('ii...->...') of the reference."""
        d = self.rank // 2
        retain = [int(q) for q in retain_indices]
        out = DeviceVector(_lib.alloc_amplitudes(2 * len(retain), self._dv.dtype), None, 2 * len(retain), self._dv.dtype)
        _lib.check(_lib.load().qcs_partial_trace(out.buf.data_ptr(), self._dv.buf.data_ptr(), d, self._dv.code,
                                                 _lib.int32_array(retain), len(retain), _lib.stream_ptr()))
        return out

    def _reduced_tensor(self, retain_indices):
        """The reduced tensor as a host array (the reference's private helper, density_operator.py:127-138)."""
        retain = list(retain_indices)
        return self._reduced_device(retain).to_host().reshape([2] * (2 * len(retain)))

    def reduced_density_operator(self, qubit_indices):
        """Reduced density operator of the listed qubits (density_operator.py:169-200)."""
        if isinstance(qubit_indices, int):
            qubit_indices = [qubit_indices]
        qubit_indices = list(qubit_indices)
        if qubit_indices == []:
            raise ValueError('Must retain at least one qubit.')
        if min(qubit_indices) < 0 or max(qubit_indices) >= self.rank // 2:
            raise ValueError('Trying to measure qubit index i not 0<=i<d, '
                             'where d is the rank of the state vector.')
        if len(qubit_indices) != len(set(qubit_indices)):
            raise ValueError('Qubit indices list contains repeated elements.')
        return DensityOperator(self._reduced_device(qubit_indices))

    def purify(self):
        """A 2d-qubit pure state whose reduced density operator on the first d qubits is this one
        (density_operator.py:149-167); host eigendecomposition."""
        from .state import State
        eigvals, eigvecs = np.linalg.eig(self.to_matrix())
        column = (np.sqrt(np.expand_dims(eigvals, 0)) * eigvecs).flatten()
        return State.from_column_vector(column, dtype=self._dv.dtype)

    # -- measurement ------------------------------------------------------------------------------------------------
    def _measurement_probabilities(self, qubit_indices):
        """Marginals of the diagonal = diagonal of the partial trace (density_operator.py:140-147)."""
        d = self.rank // 2
        diag, _ = self._diagonal()
        ps = self._dv.marginals([d - 1 - int(q) for q in qubit_indices], real_input=diag)
        unmeasured = list(set(range(d)) - set(qubit_indices))
        return ps, unmeasured

    def measure(self, qubit_indices=None, remove=False, u=None):
        """Measure the listed qubits in the computational basis, in place (density_operator.py:202-286).
        ``u``: optional caller-supplied uniform draw (see State.measure)."""
        from .state import _sample
        int_arg = False
        if isinstance(qubit_indices, int):
            qubit_indices = [qubit_indices]
            int_arg = True
        if qubit_indices is None:
            qubit_indices = range(self.rank // 2)
        qubit_indices = list(qubit_indices)
        if qubit_indices == []:
            raise ValueError('Must measure at least one qubit.')
        if min(qubit_indices) < 0 or max(qubit_indices) >= self.rank // 2:
            raise ValueError('Trying to measure qubit index i not 0<=i<d, '
                             'where d is the rank of the state vector.')
        if len(qubit_indices) != len(set(qubit_indices)):
            raise ValueError('Qubit indices list contains repeated elements.')

        d = self.rank // 2
        ps, _ = self._measurement_probabilities(qubit_indices)
        outcome = _sample(ps, u)
        m = len(qubit_indices)
        bits = tuple((outcome >> i) & 1 for i in range(m - 1, -1, -1))
        # P_m rho P_m / p(m): fix the row AND column bit of every measured qubit to its outcome bit
        bitpos, doubled = [], 0
        for q, b in zip(qubit_indices, bits):
            bitpos += [2 * (d - 1 - int(q)) + 1, 2 * (d - 1 - int(q))]
            doubled = (doubled << 2) | (3 if b else 0)
        self._dv = self._dv.collapsed(bitpos, doubled, remove, scale=1.0 / ps[outcome], track_norm=False)
        return bits[0] if int_arg else bits
