"""Tensor: the container base of State, Operator and DensityOperator.

Mirrors reference qcircuits/tensors.py:4-107 (``_t``, ``shape``, ``rank``, indexing, ``*`` and
``**``).  Operators keep ``_t`` as a host complex128 ndarray (they are parameters: a few
entries).  States and density operators override ``_t`` with a property backed by a device
buffer (see _device.py); everything below is written against ``self._t`` so it serves both.
"""
import numpy as np


class Tensor:
    def __init__(self, tensor):
        # always a private complex128 copy, like the reference (tensors.py:17-18)
        self._t = np.array(tensor, dtype=np.complex128)

    def __str__(self):
        return str(self._t)

    def __getitem__(self, key):
        return self._t[key]

    @property
    def shape(self):
        """(2,) * d for a d-qubit state, (2,) * 2d for a d-qubit operator."""
        return self._t.shape

    @property
    def rank(self):
        return len(self.shape)

    def tensor_product(self, arg):
        """A (x) B; the left factor's qubits come first (tensors.py:56-76)."""
        return self.__class__(np.multiply.outer(self._t, arg._t))

    def tensor_power(self, n):
        """A (x) A (x) ... n times (tensors.py:78-101)."""
        acc = self._t
        for _ in range(n - 1):
            acc = np.multiply.outer(acc, self._t)
        return self.__class__(acc)

    def __mul__(self, arg):
        return self.tensor_product(arg)

    def __pow__(self, n):
        return self.tensor_power(n)
