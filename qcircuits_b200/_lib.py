"""ctypes binding of libqcs.so (include/qcs.h) and the device-buffer plumbing around it.

PyTorch is used only to own device memory and to name the current CUDA stream; every
amplitude-touching operation is a call into the hand-written sm_100a kernels of libqcs.
There is no CPU fallback: if the library or a CUDA device is missing, operations raise.
"""
import ctypes
import os

import numpy as np

QCS_C64, QCS_C128, QCS_F64 = 0, 1, 2
QCS_OP_DENSE, QCS_OP_DIAG = 0, 1
QCS_MAX_OP_QUBITS = 4
QCS_MAX_OPS = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('QCS_LIB') or os.path.join(_HERE, 'libqcs.so')     # (QCS_LIB: planner experiments)

c_void_p, c_int, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_dblp = ctypes.POINTER(ctypes.c_double)


class QcsOp(ctypes.Structure):
    """struct qcs_op of include/qcs.h."""
    _fields_ = [('kind', ctypes.c_int32), ('k', ctypes.c_int32),
                ('bitpos', ctypes.c_int32 * QCS_MAX_OP_QUBITS),
                ('ctrl_mask', ctypes.c_uint64), ('matrix', c_void_p)]


# name -> (restype, argtypes); the order and meaning are those of include/qcs.h
PROTOTYPES = {
    'qcs_version': (c_int, []),
    'qcs_last_error': (ctypes.c_char_p, []),
    'qcs_workspace_bytes': (ctypes.c_int64, []),
    'qcs_apply_uf': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'qcs_plan_order': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    'qcs_plan_blocks': (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'qcs_tile_limits': (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    'qcs_set_tile_mode': (c_int, [c_int]),
    'qcs_apply_gate': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    'qcs_apply_dense': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    'qcs_plan_stats': (c_int, [c_int, c_int, c_void_p, c_int, c_void_p]),
    'qcs_apply_fused': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    'qcs_apply_fused_bounded': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    'qcs_abs2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'qcs_dot': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'qcs_scale': (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                          c_void_p]),
    'qcs_lincomb': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p]),
    'qcs_kron': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p]),
    'qcs_permute_bits': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'qcs_conj_permute_bits': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'qcs_set_basis': (c_int, [c_void_p, c_int, c_int, ctypes.c_uint64, c_void_p]),
    'qcs_marginal_probs': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                   ctypes.c_int64, c_void_p]),
    'qcs_collapse': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_uint64, c_int, c_double,
                             c_void_p, c_void_p, c_void_p]),
    'qcs_diag_probs': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'qcs_partial_trace': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    'qcs_dist_init': (c_int, [c_int, c_void_p, c_void_p]),
    'qcs_dist_swap': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'qcs_dist_destroy': (c_int, [c_void_p]),
    'qcs_dist_swap_rank': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'qcs_dist_alloc': (c_int, [c_void_p, ctypes.c_int64]),
    'qcs_dist_free': (c_int, [c_void_p]),
    'qcs_ipc_get_handle': (c_int, [c_void_p, c_void_p]),
    'qcs_ipc_open': (c_int, [c_void_p, c_void_p]),
    'qcs_ipc_close': (c_int, [c_void_p]),
    'qcs_outer_accumulate': (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_int, c_void_p]),
}

_lib = None


def load():
    """Load libqcs.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libqcs.so not found at %s -- build it with `python -c "import __graft_entry__ as g; '
                               'g.build()"` (there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class QcsError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().qcs_last_error()
        raise QcsError('libqcs error %d: %s' % (rc, msg.decode() if msg else '?'))


# --------------------------------------------------------------------------------------------------
# device plumbing (torch owns the memory)
# --------------------------------------------------------------------------------------------------
_torch = None
_workspaces = {}
MARGINAL_WS_BYTES = 8 << 20


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError('qcircuits_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return t


def stream_ptr():
    return torch().cuda.current_stream().cuda_stream


def workspace():
    """Zero-initialised scratch of the CURRENT stream of the current device: reduction tickets/partials, marginal
    histograms, the block-matrix images of the pass in flight.  Calls that share a workspace must be stream-ordered
    (include/qcs.h), so every (device, stream) pair gets its own."""
    t = require_cuda()
    key = (t.cuda.current_device(), stream_ptr())
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(load().qcs_workspace_bytes()) + MARGINAL_WS_BYTES
        ws = t.zeros(nbytes, dtype=t.uint8, device='cuda')
        _workspaces[key] = ws
    return ws


def dtype_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.complex64:
        return QCS_C64
    if np_dtype == np.complex128:
        return QCS_C128
    raise TypeError('amplitude dtype must be complex64 or complex128')


def real_torch_dtype(np_dtype):
    t = torch()
    return t.float32 if np.dtype(np_dtype) == np.complex64 else t.float64


def alloc_amplitudes(nbits, np_dtype):
    """Uninitialised device buffer of 2^nbits complex amplitudes (as interleaved reals)."""
    t = require_cuda()
    return t.empty(2 << nbits, dtype=real_torch_dtype(np_dtype), device='cuda')


def alloc_scalar():
    t = require_cuda()
    return t.empty(2, dtype=t.float64, device='cuda')


_units = {}


def unit_scalar():
    """device float64 holding 1.0 (per device, never written): the squared norm of a basis state"""
    t = require_cuda()
    dev = t.cuda.current_device()
    u = _units.get(dev)
    if u is None:
        u = _units[dev] = t.ones(2, dtype=t.float64, device='cuda')
    return u


def upload(array, np_dtype):
    """Host complex array (any shape) -> flat device buffer of interleaved reals."""
    t = require_cuda()
    host = np.ascontiguousarray(np.asarray(array, dtype=np_dtype).reshape(-1))
    return t.from_numpy(host.view(np.float32 if np.dtype(np_dtype) == np.complex64 else np.float64)).to('cuda')


def download(buf, np_dtype):
    """Flat device buffer of interleaved reals -> host complex vector."""
    return buf.cpu().numpy().view(np_dtype)


def int32_array(values):
    arr = (ctypes.c_int32 * max(1, len(values)))(*[int(v) for v in values])
    return arr


def host_matrix(M):
    """complex matrix/vector -> contiguous complex128 buffer + pointer (keep the array alive)."""
    arr = np.ascontiguousarray(np.asarray(M, dtype=np.complex128))
    return arr, arr.ctypes.data
