"""qcircuits_b200: the qcircuits Python API on a B200-native (sm_100a) state-vector engine.

Same public names as reference qcircuits/__init__.py:1-13.  ``import qcircuits_b200 as qc`` is a
drop-in for ``import qcircuits as qc``; ``install_as_qcircuits()`` additionally registers this
package under the name ``qcircuits`` so unmodified scripts and the reference's own tests run on it.
"""
from .operators import Identity, PauliX, PauliY, PauliZ
from .operators import Hadamard, Phase, PiBy8, SqrtNot
from .operators import Rotation, RotationX, RotationY, RotationZ
from .operators import CNOT, Toffoli, Swap, SqrtSwap
from .operators import ControlledU, U_f
from .operators import Operator
from .state import qubit, zeros, ones, bitstring
from .state import positive_superposition, bell_state
from .state import State
from .density_operator import DensityOperator
from . import operators, state, density_operator, tensors

__version__ = '0.6.0+b200.1'


def install_as_qcircuits():
    """Make ``import qcircuits`` (and its submodules) resolve to this engine."""
    import sys
    me = sys.modules[__name__]
    sys.modules['qcircuits'] = me
    for sub in ('operators', 'state', 'density_operator', 'tensors'):
        sys.modules['qcircuits.' + sub] = getattr(me, sub)
    return me
