"""DeviceVector: a flat device array of 2^nbits complex amplitudes plus its lazy squared norm.

This is the Python side of the conventions in include/qcs.h: ``buf`` is a torch tensor of
interleaved reals (torch only owns the memory), ``norm`` is ``None`` (contents are taken as they
are) or a device float64 holding the squared norm of ``buf`` (the represented vector is
``buf / sqrt(norm)``).  States (nbits = d) and density operators (nbits = 2d) both sit on it.
Every method is a thin wrapper over one libqcs call on the current CUDA stream.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import QcsOp, check


def _ptr(t):
    return None if t is None else t.data_ptr()


class DeviceVector:
    __slots__ = ('buf', 'norm', 'nbits', 'dtype', 'bound')

    def __init__(self, buf, norm, nbits, dtype, bound=None):
        self.buf, self.norm, self.nbits, self.dtype = buf, norm, nbits, np.dtype(dtype)
        # device scalar >= the squared norm of ``buf`` when that is known without measuring (a basis state: 1), else
        # None.  It only ranges the fp16 operands of the tensor-core sweeps (qcs_apply_fused_bounded) when ``norm`` is None.
        self.bound = bound

    # -- construction / export ---------------------------------------------------------------------
    @classmethod
    def from_host(cls, array, dtype):
        array = np.asarray(array)
        nbits = int(round(np.log2(array.size))) if array.size else 0
        if array.size != 2 ** nbits or any(s != 2 for s in array.shape):
            raise ValueError('tensor must have shape (2,) * d')
        return cls(_lib.upload(array, dtype), None, nbits, dtype)

    @classmethod
    def basis(cls, nbits, index, dtype):
        buf = _lib.alloc_amplitudes(nbits, dtype)
        check(_lib.load().qcs_set_basis(buf.data_ptr(), nbits, _lib.dtype_code(dtype), index, _lib.stream_ptr()))
        return cls(buf, None, nbits, dtype, bound=_lib.unit_scalar())

    @property
    def code(self):
        return _lib.dtype_code(self.dtype)

    def to_host(self):
        """complex128 host vector of the represented amplitudes (pending scale applied)."""
        v = _lib.download(self.buf, self.dtype).astype(np.complex128)
        if self.norm is not None:
            v /= np.sqrt(float(self.norm[0].item()))
        return v

    def clone(self):
        return DeviceVector(self.buf.clone(), None if self.norm is None else self.norm.clone(), self.nbits, self.dtype,
                            bound=self.bound)

    def _new_like(self, nbits=None):
        nbits = self.nbits if nbits is None else nbits
        return DeviceVector(_lib.alloc_amplitudes(nbits, self.dtype), None, nbits, self.dtype)

    # -- gate passes ---------------------------------------------------------------------------------
    def apply_ops(self, ops, keep, track_norm=True, out=None):
        """One fused pass (qcs_apply_fused).  ``ops`` is a ctypes array of QcsOp, ``keep`` the host
        arrays its matrix pointers refer to.  Returns a new DeviceVector (or writes ``out``)."""
        lib = _lib.load()
        dst = self._new_like() if out is None else out
        norm_out = _lib.alloc_scalar() if track_norm else None
        ws = _lib.workspace()
        bound = self.bound if self.norm is None else None
        check(lib.qcs_apply_fused_bounded(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, self.code, ops, len(ops),
                                          _ptr(self.norm), _ptr(norm_out), _ptr(bound), ws.data_ptr(), _lib.stream_ptr()))
        dst.norm = norm_out
        del keep
        return dst

    def apply_dense(self, dev_matrix, k, bitpos, track_norm=True):
        lib = _lib.load()
        dst = self._new_like()
        norm_out = _lib.alloc_scalar() if track_norm else None
        ws = _lib.workspace()
        check(lib.qcs_apply_dense(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, self.code,
                                  dev_matrix.data_ptr(), k, _lib.int32_array(bitpos), _ptr(self.norm),
                                  _ptr(norm_out), ws.data_ptr(), _lib.stream_ptr()))
        dst.norm = norm_out
        return dst

    # -- reductions -------------------------------------------------------------------------------------
    def norm2(self):
        """Squared norm of the raw buffer (device scalar)."""
        out = _lib.alloc_scalar()
        check(_lib.load().qcs_abs2(None, self.buf.data_ptr(), self.nbits, self.code, None, out.data_ptr(),
                                   _lib.workspace().data_ptr(), _lib.stream_ptr()))
        return out

    def probabilities(self):
        """(host float64 vector of |amp|^2 of the represented vector, their sum)"""
        t = _lib.torch()
        probs = t.empty(1 << self.nbits, dtype=t.float64, device='cuda')
        total = _lib.alloc_scalar()
        check(_lib.load().qcs_abs2(probs.data_ptr(), self.buf.data_ptr(), self.nbits, self.code, _ptr(self.norm),
                                   total.data_ptr(), _lib.workspace().data_ptr(), _lib.stream_ptr()))
        return probs.cpu().numpy(), float(total[0].item())

    def dot(self, other):
        out = _lib.alloc_scalar()
        check(_lib.load().qcs_dot(self.buf.data_ptr(), other.buf.data_ptr(), self.nbits, self.code, out.data_ptr(),
                                  _lib.workspace().data_ptr(), _lib.stream_ptr()))
        re, im = out.cpu().numpy()
        s = 1.0
        for v in (self, other):
            if v.norm is not None:
                s /= np.sqrt(float(v.norm[0].item()))
        return np.complex128(complex(re, im) * s)

    def marginals(self, bitpos, real_input=None):
        """float64 marginal distribution over the listed bits (bitpos[0] = MSB of the outcome)."""
        t = _lib.torch()
        m = len(bitpos)
        out = t.empty(1 << m, dtype=t.float64, device='cuda')
        ws = _lib.workspace()
        if real_input is None:
            src, code, nbits, norm = self.buf, self.code, self.nbits, self.norm
        else:
            src, code, nbits, norm = real_input, _lib.QCS_F64, int(round(np.log2(real_input.numel()))), None
        check(_lib.load().qcs_marginal_probs(out.data_ptr(), src.data_ptr(), nbits, code, _lib.int32_array(bitpos), m,
                                             _ptr(norm), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out.cpu().numpy()

    # -- elementwise --------------------------------------------------------------------------------------
    def scaled(self, scalar):
        scalar = complex(scalar)
        dst = self._new_like()
        check(_lib.load().qcs_scale(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, self.code, scalar.real,
                                    scalar.imag, _ptr(self.norm), None, None, _lib.stream_ptr()))
        return dst

    def lincomb(self, other, alpha, beta):
        a = (ctypes.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        b = (ctypes.c_double * 2)(complex(beta).real, complex(beta).imag)
        dst = self._new_like()
        check(_lib.load().qcs_lincomb(dst.buf.data_ptr(), self.buf.data_ptr(), other.buf.data_ptr(), self.nbits,
                                      self.code, a, b, _ptr(self.norm), _ptr(other.norm), None, None,
                                      _lib.stream_ptr()))
        return dst

    def kron(self, other):
        dst = self._new_like(self.nbits + other.nbits)
        check(_lib.load().qcs_kron(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, other.buf.data_ptr(),
                                   other.nbits, self.code, _ptr(self.norm), _ptr(other.norm), None, None,
                                   _lib.stream_ptr()))
        return dst

    def permuted(self, src_bit, conj=False):
        """dst[i] = src[j] (conj: its complex conjugate) with bit src_bit[b] of j = bit b of i."""
        dst = self._new_like()
        fn = _lib.load().qcs_conj_permute_bits if conj else _lib.load().qcs_permute_bits
        check(fn(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, self.code, _lib.int32_array(src_bit),
                 _lib.stream_ptr()))
        dst.norm = self.norm
        return dst

    def collapsed(self, bitpos, outcome, remove, scale=1.0, track_norm=True):
        m = len(bitpos)
        dst = self._new_like(self.nbits - m if remove else self.nbits)
        norm_out = _lib.alloc_scalar() if track_norm else None
        check(_lib.load().qcs_collapse(dst.buf.data_ptr(), self.buf.data_ptr(), self.nbits, self.code,
                                       _lib.int32_array(bitpos), m, outcome, 1 if remove else 0, scale,
                                       _ptr(norm_out), _lib.workspace().data_ptr(), _lib.stream_ptr()))
        dst.norm = norm_out
        return dst


def make_ops(plans_and_bits):
    """Build a ctypes QcsOp array from [(GatePlan, bit of each operator qubit, conj?)].
    Returns (ops, keepalive)."""
    ops = (QcsOp * len(plans_and_bits))()
    keep = []
    for o, (plan, bits, conj) in zip(ops, plans_and_bits):
        mat = np.conj(plan.matrix) if conj else plan.matrix
        arr, ptr = _lib.host_matrix(mat)
        keep.append(arr)
        o.kind = _lib.QCS_OP_DIAG if plan.diagonal else _lib.QCS_OP_DENSE
        o.k = len(plan.targets)
        for j, q in enumerate(plan.targets):
            o.bitpos[j] = bits[q]
        mask = 0
        for q in plan.controls:
            mask |= 1 << bits[q]
        o.ctrl_mask = mask
        o.matrix = ptr
    return ops, keep
