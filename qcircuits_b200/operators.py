"""Operator, OperatorBase and the gate factories -- the drop-in for reference
qcircuits/operators.py.

Gate tensors are parameters (at most 2^22 entries in any configuration), so they stay host
NumPy arrays; Operator-on-Operator composition (operators.py:225-278 with an OperatorBase
argument) is host arithmetic for small operators and a device gate pass over the out-wires from
8 qubits up (``Operator._compose_on_device``).  What changes is Operator-on-State and Operator-on-DensityOperator:
those no longer run ``np.tensordot`` over the rank-n tensor but call the sm_100a kernels of
libqcs (qcs_apply_fused / qcs_apply_dense) on the device buffer, with the reference's
renormalise-after-every-gate folded into the same pass (see include/qcs.h, "lazy normalisation").
"""
import os
from itertools import product

import numpy as np

from . import _gates
from .tensors import Tensor


def _pair_axes(qubit_order):
    """axes of an interleaved operator tensor when its qubits are listed in qubit_order"""
    return [a for q in qubit_order for a in (2 * q, 2 * q + 1)]


class OperatorBase(Tensor):
    """Common functionality of Operator and DensityOperator (reference operators.py:24-170)."""

    def __init__(self, tensor):
        super().__init__(tensor)

    @staticmethod
    def from_matrix(M):
        """Matrix (Kronecker-product convention) -> Operator (operators.py:37-73)."""
        if type(M) is list:
            M = np.array(M, dtype=np.complex128)
        if M.ndim != 2 or M.shape[0] != M.shape[1]:
            raise ValueError('The matrix should be square.')
        d = np.log2(M.shape[0])
        if not float(d).is_integer():
            raise ValueError('The matrix dimension should be a power of 2.')
        d = int(d)
        # matrix axes are (out_0..out_{d-1}, in_0..in_{d-1}); the tensor interleaves them
        interleave = np.arange(2 * d).reshape(2, d).T.reshape(-1)
        return Operator(np.reshape(M, [2] * (2 * d)).transpose(interleave))

    @staticmethod
    def _tensor_to_matrix(t):
        d = t.ndim // 2
        outs_then_ins = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
        return t.transpose(outs_then_ins).reshape(2 ** d, 2 ** d)

    def to_matrix(self):
        """Operator -> matrix (operators.py:81-93)."""
        return self._tensor_to_matrix(self._t)

    @property
    def adj(self):
        """Conjugate transpose: conjugate and swap every (out, in) axis pair (operators.py:95-114)."""
        r = self.rank
        if isinstance(self, Operator) and r >= DEVICE_COMPOSE_MIN_BITS and _cuda_ready():
            # large operators: one conjugating bit permutation on the device, flat bit b <- b ^ 1;
            # the result stays on the device until somebody reads its tensor
            return Operator._from_device(self._device_tensor().permuted([b ^ 1 for b in range(r)], conj=True))
        swap_pairs = np.arange(r).reshape(r // 2, 2)[:, ::-1].reshape(-1)
        return self.__class__(np.conj(self._t).transpose(swap_pairs))

    def _permuted_tensor(self, axes, inverse=False):
        if inverse:
            axes = np.argsort(axes)
        return np.transpose(self._t, _pair_axes(axes))

    def permute_qubits(self, axes, inverse=False):
        """Relabel the qubits (both wires of each); in place, returns self (operators.py:125-146)."""
        self._t = self._permuted_tensor(axes, inverse=inverse)
        return self

    def swap_qubits(self, axis1, axis2):
        """Swap two qubits (both wires); in place, returns self (operators.py:148-170)."""
        order = list(range(self.rank // 2))
        order[axis1], order[axis2] = order[axis2], order[axis1]
        self._t = np.transpose(self._t, _pair_axes(order))
        return self


from .density_operator import DensityOperator  # noqa: E402  (circular by design, like the reference)
from .state import State  # noqa: E402


def _validated_indices(op_qubits, d, qubit_indices):
    """The reference's argument checks, same order and exception type (operators.py:234-253)."""
    if op_qubits > d:
        raise ValueError('An operator for a d-rank state space can only be applied to '
                         'a system whose rank is >= d.')
    if op_qubits < d and qubit_indices is None:
        raise ValueError('Applying operator to too-large system without supplying '
                         'qubit indices.')
    if qubit_indices is None:
        return list(range(d))
    qubit_indices = [int(q) for q in qubit_indices]
    if len(set(qubit_indices)) != len(qubit_indices):
        raise ValueError('Qubit indices list contains repeated elements.')
    if min(qubit_indices) < 0:
        raise ValueError('Supplied qubit index < 0.')
    if max(qubit_indices) >= d:
        raise ValueError('Supplied qubit index larger than system size.')
    if len(qubit_indices) != op_qubits:
        raise ValueError('Length of qubit_indices does not match operator.')
    return qubit_indices


DEVICE_COMPOSE_MIN_BITS = 16      # operator-on-operator composition runs on the device from 8 qubits up


def _cuda_ready():
    from . import _lib
    try:
        return bool(_lib.torch().cuda.is_available()) and os.path.exists(_lib.LIB_PATH)
    except Exception:
        return False


class Operator(OperatorBase):
    """An operator on the state space of d qubits: a rank-2d tensor, axes (out0, in0, out1, in1, ...)."""

    def __init__(self, tensor):
        super().__init__(tensor)

    # ``_t`` is a property so that everything derived from the tensor -- the structure analysis of the matrix and
    # its device copies -- is dropped whenever the tensor changes: permute_qubits / swap_qubits rewrite it in place
    # (reference operators.py:125-170) and it may be assigned directly; the reference recomputes from _t on every call.
    # Large operators (from 8 qubits up) that were produced on the device -- compositions, adjoints -- keep their tensor
    # there (``_dev_tensor``, complex128) and download it only when somebody reads ``_t``.
    @property
    def _t(self):
        if self._tensor is None:
            dv = self._dev_tensor
            self._tensor = dv.to_host().reshape([2] * dv.nbits)
        return self._tensor

    @_t.setter
    def _t(self, value):
        self._tensor = value
        self._plan_cache = None       # structure analysis of the matrix (diagonal / controlled)
        self._dev_matrix = {}         # dtype -> device copy, for operators wider than 4 qubits
        self._dev_tensor = None       # device copy of the tensor itself (large operators)

    @classmethod
    def _from_device(cls, dv):
        """An Operator whose tensor lives on the device (``dv``: DeviceVector of 2d bits, complex128, no pending norm)."""
        op = Operator.__new__(Operator)
        op._tensor = None
        op._plan_cache = None
        op._dev_matrix = {}
        op._dev_tensor = dv
        return op

    def _device_tensor(self):
        """The tensor as a device vector of 2d flat bits (complex128), uploaded once and kept."""
        if self._dev_tensor is None:
            from ._device import DeviceVector
            self._dev_tensor = DeviceVector.from_host(self._t, np.complex128)
        return self._dev_tensor

    @property
    def _device_only(self):
        """True while the tensor exists only on the device (nobody has read ``_t`` yet)."""
        return self._tensor is None and getattr(self, '_dev_tensor', None) is not None

    @property
    def shape(self):
        if self._device_only:
            return (2,) * self._dev_tensor.nbits
        return self._t.shape

    def __repr__(self):
        head = 'Operator('
        return head + str(self._t).replace('\n', '\n' + ' ' * len(head)) + ')'

    def __str__(self):
        return 'Operator for {}-qubit state space. Tensor:\n{}'.format(self.rank // 2, self._t)

    @classmethod
    def _wrap(cls, array):
        """An Operator around a fresh complex128 array that nobody else holds (results of the arithmetic below): the
        private copy of Tensor.__init__ (reference tensors.py:17-18) would be a second 64 MB array for 11 qubits."""
        if array.dtype != np.complex128:
            return Operator(array)
        op = Operator.__new__(Operator)
        op._t = array
        return op

    def __add__(self, arg):
        return Operator._wrap(self._t + arg._t)

    def __sub__(self, arg):
        return self + (-1) * arg

    def __neg__(self):
        return Operator._wrap(-self._t)

    def __mul__(self, scalar):
        if isinstance(scalar, (float, int, complex)):
            return Operator._wrap(scalar * self._t)
        return super().__mul__(scalar)

    def __rmul__(self, scalar):
        if isinstance(scalar, (float, int, complex)):
            return Operator._wrap(scalar * self._t)

    def __truediv__(self, scalar):
        return Operator._wrap(self._t / scalar)

    def tensor_product(self, arg):
        """A (x) B (tensors.py:56-76).  From 8 qubits up the product is made on the device (qcs_kron on the flat
        tensors: the left factor's wires are the high bits) and stays there until its tensor is read."""
        if (isinstance(arg, Operator) and self.rank + arg.rank >= DEVICE_COMPOSE_MIN_BITS and self.rank > 0 and arg.rank > 0
                and _cuda_ready()):
            return Operator._from_device(self._device_tensor().kron(arg._device_tensor()))
        if isinstance(arg, Operator):
            return Operator._wrap(np.multiply.outer(self._t, arg._t))
        return super().tensor_product(arg)

    def tensor_power(self, n):
        acc = self._t
        for _ in range(n - 1):
            acc = np.multiply.outer(acc, self._t)
        return Operator._wrap(acc) if n > 1 else Operator(acc)

    # -- application ---------------------------------------------------------------------------
    def _compose(self, arg, qubit_indices):
        """A(B) for operator-shaped B on the host: contract this operator's in-wires with B's
        out-wires on the chosen qubits (operators.py:226-274)."""
        d = arg.rank // 2
        k = self.rank // 2
        qi = _validated_indices(k, d, qubit_indices)
        if 2 * d >= DEVICE_COMPOSE_MIN_BITS and _cuda_ready():
            return self._compose_on_device(arg, qi)
        order = qi + sorted(set(range(d)) - set(qi))
        moved = np.transpose(arg._t, _pair_axes(order))
        out = np.tensordot(self._t, moved, (list(range(1, 2 * k, 2)), list(range(0, 2 * k, 2))))
        # tensordot yields (out_0..out_{k-1}, in_0..in_{k-1}, rest); restore the interleaving
        back = np.concatenate([np.arange(2 * k).reshape(2, k).T.reshape(-1), np.arange(2 * k, 2 * d)])
        result = arg.__class__(np.transpose(out, back))
        return result.permute_qubits(order, inverse=True)

    def _compose_on_device(self, arg, qi):
        """A(B) for a large operator-shaped B: the rank-2d tensor of B is a flat vector of 2d bits
        in which the out-wire of qubit q is flat bit 2(d-1-q)+1 (the row bits of the density-operator layout),
        so the composition is the ordinary gate pass of this operator on those bits -- the tile engine for up to
        4 target qubits, the dense path above (e.g. the 10-qubit Hadamard of examples/grover_algorithm.py:41).
        complex128, no renormalisation (operators.py:275 renormalises states only)."""
        from ._device import DeviceVector, make_ops
        d = arg.rank // 2
        dv = arg._device_tensor() if isinstance(arg, Operator) else DeviceVector.from_host(arg._t, np.complex128)
        bits = [2 * (d - 1 - q) + 1 for q in qi]
        plan = self._plan()
        if len(plan.targets) <= 4:
            ops, keep = make_ops([(plan, bits, False)])
            out = dv.apply_ops(ops, keep, track_norm=False)
        else:
            out = dv.apply_dense(self._device_matrix(np.complex128), plan.k, bits, track_norm=False)
        if isinstance(arg, Operator):
            return Operator._from_device(out)         # stays on the device; ``_t`` downloads on demand
        return arg.__class__(out.to_host().reshape([2] * (2 * d)))

    def _apply(self, arg, qubit_indices=None):
        if isinstance(arg, State):
            qi = _validated_indices(self.rank // 2, arg.rank, qubit_indices)
            return arg._apply_operator(self, qi)
        if isinstance(arg, DensityOperator):
            qi = _validated_indices(self.rank // 2, arg.rank // 2, qubit_indices)
            return arg._apply_operator(self, qi)
        if type(arg).__name__ == 'ShardedState':          # (dist.py is only imported by multi-GPU programs)
            qi = _validated_indices(self.rank // 2, arg.rank, qubit_indices)
            return arg._apply_operator(self, qi)
        return self._compose(arg, qubit_indices)

    def __call__(self, arg, qubit_indices=None):
        """A(v) for a State, A rho A^dagger for a DensityOperator, A(B) composition for an
        Operator; ``qubit_indices`` selects (and orders) the qubits of a larger system
        (operators.py:280-323).  The argument is never modified."""
        return self._apply(arg, qubit_indices)

    def _device_matrix(self, dtype, conj=False):
        """Device copy of the full matrix in the amplitude dtype (operators wider than 4 qubits)."""
        from . import _lib
        key = (np.dtype(dtype).str, conj, _lib.torch().cuda.current_device())
        dev = self._dev_matrix.get(key)
        if dev is None:
            # tensor (interleaved wires) -> matrix (out wires, then in wires) on the device: flat matrix bit b <- tensor
            # bit 2b (in wire), k + b <- 2b + 1 (out wire); the conjugated copy serves U rho U^dagger
            from ._device import DeviceVector
            k = self.rank // 2
            src_bit = [2 * b for b in range(k)] + [2 * b + 1 for b in range(k)]
            if np.dtype(dtype) == np.complex128:
                dvt = self._device_tensor()
            else:
                dvt = DeviceVector.from_host(self._t, dtype)
            dev = dvt.permuted(src_bit, conj=conj).buf
            self._dev_matrix[key] = dev
        return dev

    def _plan(self):
        """Structure of the matrix, cached: see _gates.analyze.  Operators wider than ANALYZE_MAX_QUBITS go to the
        dense kernel, which has no use for control / target structure: they skip the analysis (a scan per qubit of
        a 2^k x 2^k matrix -- it was 2/3 of the run time of examples/grover_algorithm.py at d = 10) and never build
        the host matrix at all."""
        if self._plan_cache is None:
            k = self.rank // 2
            if k > _gates.ANALYZE_MAX_QUBITS:
                # dense-kernel operand: its matrix is made on the device from the tensor (_device_matrix)
                self._plan_cache = _gates.GatePlan(k, [], list(range(k)), False, None, None)
            else:
                self._plan_cache = _gates.analyze(np.ascontiguousarray(self.to_matrix()))
        return self._plan_cache


# ---------------------------------------------------------------------------------------------------
# gate factories (reference operators.py:328-853); matrices are the standard textbook ones
# ---------------------------------------------------------------------------------------------------
def _single(matrix, d):
    return Operator(np.array(matrix, dtype=np.complex128)).tensor_power(d)


def Identity(d=1):
    """d-qubit identity."""
    return _single([[1, 0], [0, 1]], d)


def PauliX(d=1):
    """X on each of d qubits (NOT)."""
    return _single([[0, 1], [1, 0]], d)


def PauliY(d=1):
    """Y on each of d qubits: |0> -> i|1>, |1> -> -i|0>."""
    return _single([[0, -1j], [1j, 0]], d)


def PauliZ(d=1):
    """Z on each of d qubits."""
    return _single([[1, 0], [0, -1]], d)


def Hadamard(d=1):
    """H on each of d qubits."""
    return _single(1 / np.sqrt(2) * np.array([[1.0 + 0.0j, 1.0 + 0.0j], [1.0 + 0.0j, -1.0 + 0.0j]]), d)


def Phase(d=1):
    """S = diag(1, i)."""
    return _single([[1, 0], [0, 1j]], d)


def PiBy8(d=1):
    """T = diag(1, exp(i pi/4))."""
    return _single([[1, 0], [0, np.exp(1j * np.pi / 4)]], d)


def Rotation(v, theta):
    """Rotation of the Bloch vector by theta about the real unit 3-vector v."""
    v = np.array(v)
    if v.shape != (3,) or abs(v.dot(v) - 1.0) > 1e-8 or not np.all(np.isreal(v)):
        raise ValueError('Rotation vector v should be a 3D real unit vector.')
    return np.cos(theta / 2) * Identity() - 1j * np.sin(theta / 2) * (
        v[0] * PauliX() + v[1] * PauliY() + v[2] * PauliZ())


def RotationX(theta):
    return Rotation([1., 0., 0.], theta)


def RotationY(theta):
    return Rotation([0., 1., 0.], theta)


def RotationZ(theta):
    return Rotation([0., 0., 1.], theta)


def SqrtNot(d=1):
    """Square root of X."""
    return _single(0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]), d)


def _from_basis_map(d, image_of):
    """Permutation operator sending basis bit-tuple b to image_of(b)."""
    t = np.zeros([2] * (2 * d), dtype=np.complex128)
    for bits in product([0, 1], repeat=d):
        out = image_of(bits)
        t[tuple(x for pair in zip(out, bits) for x in pair)] = 1.0
    return Operator(t)


def CNOT():
    """Flip the second qubit if the first is set."""
    return _from_basis_map(2, lambda b: (b[0], b[1] ^ b[0]))


def Toffoli():
    """Flip the third qubit if the first two are set."""
    return _from_basis_map(3, lambda b: (b[0], b[1], b[2] ^ (b[0] & b[1])))


def Swap():
    """Exchange two qubits."""
    return _from_basis_map(2, lambda b: (b[1], b[0]))


def SqrtSwap():
    """Square root of Swap."""
    p, m = 0.5 * (1 + 1j), 0.5 * (1 - 1j)
    return OperatorBase.from_matrix(np.array([[1, 0, 0, 0], [0, p, m, 0], [0, m, p, 0], [0, 0, 0, 1]]))


def ControlledU(U):
    """(d+1)-qubit operator applying the d-qubit U when the first qubit is set."""
    d = U.rank // 2 + 1
    t = np.zeros([2] * (2 * d), dtype=np.complex128)
    t[0, 0] = Identity(d - 1)._t
    t[1, 1] = U._t
    return Operator(t)


class UfOperator(Operator):
    """U_f tagged at construction: the Boolean function is kept as a truth table and applied to a State as
    an index permutation (qcs_apply_uf) -- the 2^(2d)-entry tensor of reference operators.py:809-853 is only built if
    somebody reads ``_t`` (composition with another operator, to_matrix, a DensityOperator argument), and any in-place
    change of the tensor (permute_qubits, swap_qubits, assignment) turns it back into an ordinary Operator."""

    def __init__(self, table, d=None):
        self._dev_table = {}
        if d is None:
            # ``op.__class__(tensor)`` -- how the base classes build results (.adj, composition): an ordinary operator
            Operator.__init__(self, table)
            self._uf_d = self._tensor.ndim // 2
            self._uf_table = None
            return
        self._uf_table = np.ascontiguousarray(table, dtype=np.uint8)       # f over the 2^(d-1) inputs, x MSB-first
        self._uf_d = d
        self._tensor = None
        self._plan_cache = None
        self._dev_matrix = {}
        self._dev_tensor = None

    @property
    def _t(self):
        if self._tensor is None:
            d = self._uf_d
            cols = np.arange(2 ** d)
            rows = (cols & ~1) | ((cols & 1) ^ self._uf_table[cols >> 1])
            M = np.zeros((2 ** d, 2 ** d), dtype=np.complex128)
            M[rows, cols] = 1.0
            interleave = np.arange(2 * d).reshape(2, d).T.reshape(-1)
            self._tensor = np.reshape(M, [2] * (2 * d)).transpose(interleave)
        return self._tensor

    @_t.setter
    def _t(self, value):
        Operator._t.fset(self, value)
        self._uf_table = None            # no longer known to be U_f

    @property
    def shape(self):
        return (2,) * (2 * self._uf_d) if self._tensor is None else self._tensor.shape

    def _device_table(self):
        from . import _lib
        t = _lib.require_cuda()
        key = t.cuda.current_device()
        dev = self._dev_table.get(key)
        if dev is None:
            bits = np.zeros(((self._uf_table.size + 31) // 32) * 32, dtype=np.uint8)
            bits[:self._uf_table.size] = self._uf_table
            words = np.packbits(bits.reshape(-1, 32)[:, ::-1], axis=1).view('>u4').astype(np.uint32).reshape(-1)
            dev = self._dev_table[key] = t.from_numpy(words.view(np.int32)).to('cuda')
        return dev

    def _apply(self, arg, qubit_indices=None):
        if isinstance(arg, State) and self._uf_table is not None:
            qi = _validated_indices(self._uf_d, arg.rank, qubit_indices)
            return arg._apply_uf(self, qi)
        return super()._apply(arg, qubit_indices)


def U_f(f, d):
    """d-qubit operator flipping the last qubit when the Boolean f of the first d-1 bits is 1."""
    if d < 2:
        raise ValueError('U_f operator requires rank >= 2.')
    table = np.zeros(2 ** (d - 1), dtype=np.uint8)
    for x, bits in enumerate(product([0, 1], repeat=d - 1)):
        r = f(*bits)
        if r not in [0, 1]:
            raise RuntimeError('Function f for U_f operator should be Boolean,'
                               'i.e., return 0 or 1.')
        table[x] = r
    return UfOperator(table, d)
